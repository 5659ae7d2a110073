"""Golden fixtures: result rows computed by the REFERENCE's own runtime functions (compiled from the reference's
sources into oracle/_ref by the build container, tests/golden/make_golden.py) and committed, so that parity stays
pinned on machines where /root/reference does not exist.  CPU: the oracle's restatement must reproduce them exactly.
GPU: the CUDA path must reproduce them — bit-exact keys / COUNT / MIN / MAX / integer SUM, 1e-9 relative for fp64
SUM / AVG, 1e-5 for fp32 aggregates (accumulated in double on the GPU, in float by the reference)."""
import gzip
import json
import os

import pytest

from tests import util
from tests.golden import tables as G

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    with gzip.open(os.path.join(_HERE, "golden", "r1_reference_rows.json.gz"), "rb") as f:
        fx = json.loads(f.read().decode())
    assert fx["seed"] == G.SEED and fx["rows"] == G.ROWS
    dec = lambda x: float.fromhex(x) if isinstance(x, str) else x  # noqa: E731
    return {name: [tuple(dec(x) for x in r) for r in q["rows"]] for name, q in fx["queries"].items()}, fx["queries"]


@pytest.fixture(scope="module")
def golden():
    rows, meta = _load()
    st = util.make_storage(G.golden_tables(), fragment_size=G.FRAGMENT_SIZE)
    return rows, meta, st


def test_fixture_covers_every_golden_query(golden):
    rows, meta, _ = golden
    assert sorted(rows) == sorted(q[0] for q in G.QUERIES)
    assert all(len(r) > 0 for r in rows.values())


@pytest.mark.parametrize("name,text,nk,kw", G.QUERIES, ids=[q[0] for q in G.QUERIES])
def test_oracle_port_reproduces_reference_rows(oracle_mod, golden, name, text, nk, kw):
    rows, meta, st = golden
    pq = util.plan_sql(st, text, **kw)
    assert int(pq.qmd.hash_type) == meta[name]["hash_type"] and int(pq.qmd.entry_count) == meta[name]["entry_count"]
    buf, err = util.run_oracle(oracle_mod, st, pq, kind="port", n_threads=1)
    assert err == 0
    got = util.sort_rows(util.result_columns(oracle_mod, pq, buf), nk)
    util.assert_rows_equal(got, rows[name], rel=0.0)   # same call sequence, same summation order: bit for bit


@pytest.mark.gpu
@pytest.mark.parametrize("name,text,nk,kw", G.QUERIES, ids=[q[0] for q in G.QUERIES])
def test_cuda_path_reproduces_reference_rows(oracle_mod, golden, name, text, nk, kw):
    import torch
    from hdk_b200 import sql
    from hdk_b200.executor import Executor
    rows, meta, st = golden
    ex = Executor(st, kw.get("cfg"))
    pq = ex.plan(sql.parse(text, st.tables), kw.get("max_groups_buffer_entry_count"), kw.get("output_columnar"))
    prep = ex.prepare(pq)
    ex.launch(pq, prep)
    torch.cuda.synchronize()
    assert int(prep["err"].item()) == 0
    got = util.sort_rows(util.result_columns(oracle_mod, pq, prep["out"].cpu().numpy()), nk)
    tol = 1e-5 if any(ti.float_argument_input for ti in pq.infos) else 1e-9
    util.assert_rows_equal(got, rows[name], rel=tol)
