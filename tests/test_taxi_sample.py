"""ArrowStorageTaxiTest.TaxiQuery1-4 (omniscidb/Tests/ArrowStorageSqlTest.cpp:193-243): the reference's own known answers
for the four benchmark queries on its 20-row taxi sample.  Columns: tests/golden/taxi_sample_columns.json (extracted from
the reference's CSV by tests/golden/make_taxi_sample.py).  The oracle on CPU, the CUDA path on the GPU."""
import datetime
import json
import os

import numpy as np
import pyarrow as pa
import pytest

from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))

Q1 = "SELECT cab_type, count(*) FROM trips GROUP BY cab_type"
Q2 = "SELECT passenger_count, AVG(total_amount) FROM trips GROUP BY passenger_count ORDER BY passenger_count"
Q3 = ("SELECT passenger_count, extract(year from pickup_datetime) AS pickup_year, count(*) FROM trips "
      "GROUP BY passenger_count, pickup_year ORDER BY passenger_count")
Q4 = ("SELECT passenger_count, extract(year from pickup_datetime) AS pickup_year, cast(trip_distance as int) AS distance, "
      "count(*) AS the_count FROM trips GROUP BY passenger_count, pickup_year, distance ORDER BY pickup_year, the_count desc")
f32 = np.float32
EXPECTED = {      # compare_res_data(...) of the four tests, verbatim
    Q1: [("green", 20)],
    Q2: [(1, float(f32(98.19) / f32(16))), (2, 75.0), (5, float(f32(13.58) / f32(3)))],
    Q3: [(1, 2013, 16), (2, 2013, 1), (5, 2013, 3)],
    Q4: [(1, 2013, 0, 16), (5, 2013, 0, 3), (2, 2013, 0, 1)],
}


def trips_table():
    c = json.load(open(os.path.join(HERE, "golden", "taxi_sample_columns.json")))
    ts = [datetime.datetime.strptime(s, "%Y-%m-%d %H:%M:%S") for s in c["pickup_datetime"]]
    return pa.table({"pickup_datetime": pa.array(ts, pa.timestamp("s")),
                     "passenger_count": pa.array([int(v) for v in c["passenger_count"]], pa.int16()),
                     "trip_distance": pa.array([float(v) for v in c["trip_distance"]], pa.float64()),
                     "total_amount": pa.array([float(v) for v in c["total_amount"]], pa.float64()),
                     "cab_type": pa.array(c["cab_type"])})


def check(text, got):
    exp = EXPECTED[text]
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        for a, b in zip(g, e):
            if isinstance(b, float):
                assert abs(a - b) <= 1e-6 * abs(b)      # the reference's expectations are single-precision quotients
            else:
                assert a == b


@pytest.mark.parametrize("text", [Q1, Q2, Q3, Q4])
def test_taxi_sample_known_answers_oracle(oracle_mod, text):
    from tests.test_sqlite_oracle import decode_with_dictionaries
    st = util.make_storage({"trips": trips_table()}, fragment_size=7)
    pq = util.plan_sql(st, text)
    for kind in ("port", "reference"):
        buf, err = util.run_oracle(oracle_mod, st, pq, kind=kind)
        assert err == 0
        check(text, decode_with_dictionaries(st, pq, buf))


@pytest.mark.gpu
@pytest.mark.parametrize("text", [Q1, Q2, Q3, Q4])
def test_taxi_sample_known_answers_gpu(text):
    import hdk_b200.hdk as hdk_mod
    h = hdk_mod.init()
    h.import_arrow(trips_table(), "trips", fragment_size=7)
    check(text, util.arrow_rows(h.sql(text).to_arrow()))
