"""Long-running differential fuzz on the CPU: random aggregate and join queries (generators of tests/test_fuzz_sql.py) through the
planner and the oracle against SQLite, for a range of seeds — the tests run a few seeds, this runs as many as asked:

    python tests/fuzz_long.py 200 208        # seeds 200..207, ~1900 comparisons in ~2 min

Prints every mismatch with the query.  fp32 columns mixed with double arithmetic can differ from SQLite in the 5th digit
(float * float is a float in the reference, a double in SQLite)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util
from oracle import oracle
from hdk_b200 import planner
from tests.test_fuzz_sql import queries, join_queries, join_tables
from tests.test_sqlite_oracle import reference_test_table, decode_with_dictionaries
tables=reference_test_table(); st=util.make_storage(tables, fragment_size=3)
tot=bad=0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    for text in queries(seed, 150):
        try: pq=util.plan_sql(st,text)
        except planner.UnsupportedPlan: continue
        except Exception as e: print("PLANERR",type(e).__name__,str(e)[:100],"|",text[:300]); bad+=1; continue
        buf,err=util.run_oracle(oracle,st,pq,kind="port")
        if err!=0: continue
        got=sorted(decode_with_dictionaries(st,pq,buf),key=repr); exp=sorted(util.sqlite_rows(tables,text,0),key=repr); tot+=1
        try: util.assert_rows_equal(got,exp,rel=1e-5)
        except AssertionError as e: bad+=1; print("DIFF",str(e)[:80],"|",text[:400],"|",got[:2],exp[:2])
    T=join_tables(seed); sj=util.make_storage(T, fragment_size=90)
    for text in join_queries(seed, 100):
        try: pq=util.plan_sql(sj,text)
        except planner.UnsupportedPlan: continue
        buf,err=util.run_oracle(oracle,sj,pq,kind="port")
        if err!=0: continue
        got=sorted(decode_with_dictionaries(sj,pq,buf),key=repr); exp=sorted(util.sqlite_rows(T,text,0),key=repr); tot+=1
        try: util.assert_rows_equal(got,exp,rel=1e-6)
        except AssertionError as e: bad+=1; print("JDIFF",str(e)[:80],"|",text[:400],"|",got[:2],exp[:2])
print("ran",tot,"bad",bad)
