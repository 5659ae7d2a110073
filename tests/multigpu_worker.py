"""Run under torchrun (one process per GPU): the multi-GPU paths against the CPU oracle on the union of
all ranks' rows.  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_worker.py"""
import os
import sys

import numpy as np
import pyarrow as pa
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hdk_b200 import abi, distributed as D, sql, storage  # noqa: E402
from hdk_b200.executor import Executor, ResultSet  # noqa: E402
from oracle import oracle  # noqa: E402
from tests import util  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if os.environ.get("HDK_B200_TEST_JIT"):
        from hdk_b200 import _lib
        _lib.debug_set("jit", int(os.environ["HDK_B200_TEST_JIT"]))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(7)
    n = 120_000
    t = pa.table({"s": rng.integers(0, 10, n).astype(np.int16), "k": rng.integers(0, 500, n).astype(np.int32),
                  "v": pa.array(rng.integers(-2**40, 2**40, n), mask=rng.random(n) < 0.02), "f": rng.normal(0, 100, n),
                  "big": rng.integers(0, 20000, n) * 2**33, "fk": rng.integers(0, 120, n).astype(np.int32)})
    dim = pa.table({"pk": rng.permutation(100).astype(np.int32), "attr": rng.integers(0, 7, 100).astype(np.int32)})
    full = storage.ArrowStorage()                      # every rank knows the full table only to build the expected answer
    full.import_arrow_table(t, "t", fragment_size=10_000)
    full.import_arrow_table(dim, "dim")
    st = storage.ArrowStorage()                        # this rank's shard: fragment i → rank i mod world
    st.import_arrow_table(t, "t", fragment_size=10_000, shard=(rank, world))
    st.import_arrow_table(dim, "dim")
    # all ranks must plan with the same statistics: use the full table's
    for cname in st.get_table("t").columns:
        lo, hi, hn = full.get_table("t").col_stats(cname)
        for f in st.get_table("t").fragments:
            f.stats[cname].min, f.stats[cname].max, f.stats[cname].has_nulls = lo, hi, hn
    ex = Executor(st, device=local)
    ok = True
    # (1) perfect hash: partial scan → all-reduce → finalize; every rank ends with the full result
    for text, nk in [("SELECT s, COUNT(*), SUM(v), MIN(v), MAX(v), AVG(f) FROM t GROUP BY s", 1),
                     ("SELECT k, s, COUNT(v), SUM(f), MIN(f) FROM t WHERE f > -50 GROUP BY k, s", 2),
                     ("SELECT dim.attr, SUM(t.f), COUNT(*) FROM t JOIN dim ON t.fk = dim.pk GROUP BY dim.attr", 1)]:
        pq = ex.plan(sql.parse(text, st.tables))
        prep = ex.prepare(pq)
        ex.execute_sharded(pq, prep)
        torch.cuda.synchronize()
        assert int(prep["err"].item()) == 0
        got = util.sort_rows(util.result_columns(oracle, pq, prep["out"].cpu().numpy()), nk)
        pq_full = util.plan_sql(full, text)
        obuf, oerr = util.run_oracle(oracle, full, pq_full)
        exp = util.sort_rows(util.result_columns(oracle, pq_full, obuf), nk)
        util.assert_rows_equal(got, exp)
    # (1b) the same plans with the merge over peer memory (hdk_b200_launch_exchange) instead of NCCL, three epochs each
    for text, nk in [("SELECT s, COUNT(*), SUM(v), MIN(v), MAX(v), AVG(f) FROM t GROUP BY s", 1),
                     ("SELECT k, s, COUNT(v), SUM(f), MIN(f) FROM t WHERE f > -50 GROUP BY k, s", 2),
                     ("SELECT dim.attr, SUM(t.f), COUNT(*) FROM t JOIN dim ON t.fk = dim.pk GROUP BY dim.attr", 1)]:
        pq = ex.plan(sql.parse(text, st.tables))
        prep = ex.prepare(pq)
        xchg = D.PeerExchange(ex.lib, pq.plan, pq.qmd, torch.device("cuda", local))
        pq_full = util.plan_sql(full, text)
        obuf, oerr = util.run_oracle(oracle, full, pq_full)
        exp = util.sort_rows(util.result_columns(oracle, pq_full, obuf), nk)
        for _ in range(3):
            prep["out"].zero_()
            ex.launch_exchange(pq, prep, xchg)
            torch.cuda.synchronize()
            assert int(prep["err"].item()) == 0, f"in-band error {int(prep['err'].item())}"
            got = util.sort_rows(util.result_columns(oracle, pq, prep["out"].cpu().numpy()), nk)
            util.assert_rows_equal(got, exp)
        dist.barrier()
        xchg.close()
    # (2) baseline hash: shuffle by key hash → all-to-all → local aggregate; union of ranks = full result
    text = "SELECT big, s, SUM(v), COUNT(*), SUM(f) FROM t GROUP BY big, s"
    rs, n_recv = ex.execute_partitioned(sql.parse(text, st.tables), 262144)
    mine = util.sort_rows(util.result_columns(oracle, rs.planned, rs.buffer), 2)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    keysets = [set(r[:2] for r in g) for g in gathered]
    for i in range(world):
        for j in range(i + 1, world):
            assert not (keysets[i] & keysets[j]), "a key landed on two ranks"
    union = sorted([r for g in gathered for r in g], key=lambda r: r[:2])
    pq_full = util.plan_sql(full, text, max_groups_buffer_entry_count=262144)
    obuf, oerr = util.run_oracle(oracle, full, pq_full)
    exp = util.sort_rows(util.result_columns(oracle, pq_full, obuf), 2)
    util.assert_rows_equal(union, exp)
    tot = torch.tensor([n_recv], device="cuda")
    dist.all_reduce(tot)
    assert int(tot.item()) == n
    dist.barrier()
    # (3) the public API: hdk.sql() over a sharded table picks the multi-GPU strategy itself (perfect hash → merge over peer
    #     memory, baseline hash → key-hash shuffle, join → build once + broadcast) and returns the merged result on every
    #     rank — the shapes of the five named configs against SQLite on the union of all shards
    import hdk_b200.hdk as hdk_mod
    t2 = t.append_column("ts", pa.array((rng.integers(1230768000, 1467331200, n) * 1000).astype("datetime64[ms]")))
    h = hdk_mod.init(device=local)
    h.import_arrow(t2, "t", fragment_size=10_000, shard=(rank, world))
    h.import_arrow(dim, "dim")
    tables = {"t": t2, "dim": dim}
    api_queries = [
        ("SELECT k, COUNT(*), SUM(v), MIN(v), MAX(v) FROM t GROUP BY k", 1),                                                  # config 1
        ("SELECT s, AVG(f) FROM t GROUP BY s", 1),                                                                            # taxi Q2
        ("SELECT s, EXTRACT(YEAR FROM ts) AS y, COUNT(*) FROM t GROUP BY s, y", 2),                                           # taxi Q3
        ("SELECT s, SUM(f), SUM(f * (1 - f / 1000)), AVG(f), COUNT(*) FROM t WHERE k <= 400 GROUP BY s", 1),                  # TPC-H Q1 shape
        ("SELECT big, s, SUM(v), COUNT(*) FROM t GROUP BY big, s", 2),                                                        # config 4
        ("SELECT dim.attr, SUM(t.f), COUNT(*) FROM t JOIN dim ON t.fk = dim.pk GROUP BY dim.attr", 1),                        # config 5
        ("SELECT big, COUNT(*) AS c, SUM(v) AS sv FROM t GROUP BY big ORDER BY c DESC, big LIMIT 7", 0),                      # baseline + top-k
    ]
    for text, nk in api_queries:
        order = (" ORDER BY " + ", ".join(str(i + 1) for i in range(nk))) if nk else ""
        got = util.arrow_rows(h.sql(text + order).to_arrow())
        text_sqlite = text.replace("EXTRACT(YEAR FROM ts)", "CAST(strftime('%Y', ts) AS INT)")
        exp = util.sqlite_rows(tables, text_sqlite + order, nk)
        if nk == 0:
            exp = [tuple(r) for r in exp]
            got_s, exp_s = got, sorted(exp, key=lambda r: (-r[1], r[0]))
            util.assert_rows_equal(got_s, exp_s, rel=1e-9)
        else:
            util.assert_rows_equal(got, exp, rel=1e-9)
    strategies = h.executor._peer_ok
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU OK world={world} (hdk.sql across ranks: {len(api_queries)} queries, peer memory {'on' if strategies else 'off'})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
