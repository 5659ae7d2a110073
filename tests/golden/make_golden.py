#!/usr/bin/env python
"""Regenerates tests/golden/r1_reference_rows.json.gz: the result rows of the golden queries computed by the
REFERENCE's own runtime (QueryEngine/RuntimeFunctions.cpp & co. compiled from /root/reference into
oracle/_ref/liboracle_ref.so by oracle/Makefile) driven by the oracle's restated JIT call sequence.
Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Integers are stored as such, doubles as C99 hex strings (bit exact), NULL as null; rows sorted by the key columns."""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from tests import util  # noqa: E402
from tests.golden import tables as G  # noqa: E402


def main():
    oracle.build()
    assert oracle.ref_available(), "oracle/_ref is missing: the reference sources are needed to regenerate the fixture"
    st = util.make_storage(G.golden_tables(), fragment_size=G.FRAGMENT_SIZE)
    out = {"seed": G.SEED, "rows": G.ROWS, "runtime": "reference RuntimeFunctions.cpp (oracle/_ref)", "queries": {}}
    for name, text, nk, kw in G.QUERIES:
        pq = util.plan_sql(st, text, **kw)
        buf, err = util.run_oracle(oracle, st, pq, kind="reference", n_threads=1)
        assert err == 0, (name, err)
        rows = util.sort_rows(util.result_columns(oracle, pq, buf), nk)
        enc = [[None if x is None else (x.hex() if isinstance(x, float) else x) for x in r] for r in rows]
        out["queries"][name] = {"sql": text, "n_keys": nk, "hash_type": int(pq.qmd.hash_type), "entry_count": int(pq.qmd.entry_count),
                                "rows": enc}
        print(f"{name}: {len(rows)} rows, hash_type {pq.qmd.hash_type}")
    path = os.path.join(ROOT, "tests", "golden", "r1_reference_rows.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
