"""pyhdk-shaped entry point: `hdk = hdk_b200.init(); ht = hdk.import_arrow(tbl, "t"); hdk.sql("SELECT …")`.

Mirrors python/pyhdk/hdk.py for the hot-path plan shapes: `init` (:2956-2960), `HDK.import_arrow`
(:2361-2392), `HDK.sql` (:2456-2519), `HDK.drop_table`, and the builder nodes' `agg` / `join` /
`filter` / `run` (:1606, :1747, :1832, :1992).  Queries outside the supported shapes raise
`UnsupportedPlan`; nothing ever falls back to a CPU path.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional

import pyarrow as pa

from . import ir, planner, sql
from .executor import ExecutionResult, Executor, RelAlgExecutor
from .storage import DEFAULT_FRAGMENT_SIZE, ArrowStorage


class QueryNode:
    """A scan (optionally filtered / joined) that can be aggregated — the builder API subset."""

    def __init__(self, hdk: "HDK", table: str, quals=None, joins=None):
        self._hdk = hdk
        self.table_name = table
        self._quals = list(quals or [])
        self._joins = list(joins or [])

    # column references -------------------------------------------------------------------
    def _tables(self):
        return [self.table_name] + [j.inner_table for j in self._joins]

    def ref(self, name: str) -> ir.ColumnRef:
        for tidx, tname in enumerate(self._tables()):
            t = self._hdk.storage.get_table(tname)
            if name in t.columns:
                ci = t.columns[name]
                return ir.ColumnRef(tidx, name, ci.type, ci.phys_width)
        raise KeyError(name)

    def __getitem__(self, name):
        return self.ref(name)

    # relational operators ---------------------------------------------------------------
    def filter(self, *conds) -> "QueryNode":
        qs = []
        for c in conds:
            qs.append(self._parse_expr(c) if isinstance(c, str) else c)
        return QueryNode(self._hdk, self.table_name, self._quals + qs, self._joins)

    def join(self, rhs: "QueryNode", lhs_cols, rhs_cols=None, how: str = "inner") -> "QueryNode":
        if how != "inner":
            raise planner.UnsupportedPlan("only inner equi-joins are on the hot path")
        lhs_cols = [lhs_cols] if isinstance(lhs_cols, str) else list(lhs_cols)
        rhs_cols = lhs_cols if rhs_cols is None else ([rhs_cols] if isinstance(rhs_cols, str) else list(rhs_cols))
        if len(lhs_cols) != len(rhs_cols) or not lhs_cols:
            raise ValueError("join: left and right key lists differ in length")
        inner_t = self._hdk.storage.get_table(rhs.table_name)
        try:
            keys = [(ir.join_key_for(self.ref(lc), inner_t.columns[rc].type), rc) for lc, rc in zip(lhs_cols, rhs_cols)]
        except NotImplementedError as ex:
            raise planner.UnsupportedPlan(str(ex))
        spec = ir.JoinSpec(rhs.table_name, keys[0][0], rhs_cols[0], keys[1:],   # > 1 column: baseline join table
                           [inner_t.columns[rc].type for rc in rhs_cols])
        return QueryNode(self._hdk, self.table_name, self._quals, self._joins + [spec])

    def _parse_expr(self, text: str) -> ir.Expr:
        tables = {t: self._hdk.storage.get_table(t) for t in self._tables()}
        p = sql._Parser("select " + text + " from " + self.table_name, tables)
        p.bigint_count = self._hdk.config.bigint_count
        p.scopes = [(t, t, i) for i, t in enumerate(self._tables())]
        p.i = 1
        return p.expr()

    def agg(self, group_keys, aggs=None, **kw_aggs) -> "AggNode":
        group_keys = [group_keys] if isinstance(group_keys, str) else list(group_keys)
        aggs = dict(aggs or {})
        aggs.update(kw_aggs)
        gb, targets, names = [], [], []
        for k in group_keys:
            e = self._parse_expr(k) if isinstance(k, str) else k
            gb.append(e)
            targets.append(e)
            names.append(k if isinstance(k, str) and re.fullmatch(r"\w+", k) else f"key{len(gb)}")
        for name, spec in aggs.items():
            if isinstance(spec, ir.AggExpr):
                e = spec
            else:
                s = spec.strip()
                if s.lower() == "count":
                    s = "count(*)"
                elif re.fullmatch(r"\w+", s):
                    s = f"{s}({name})"
                e = self._parse_expr(s)
            targets.append(e)
            names.append(name)
        unit = ir.ExecutionUnit(self.table_name, gb, targets, names, self._quals, self._joins)
        return AggNode(self._hdk, unit)


class AggNode:
    def __init__(self, hdk: "HDK", unit: ir.ExecutionUnit):
        self._hdk = hdk
        self.unit = unit

    def sort(self, *fields) -> "AggNode":
        order = []
        for f in fields:   # "x" | ("x", "asc" | "desc"[, "first" | "last"]); NULLs last by default (python/pyhdk/hdk.py:1687-1689)
            f = (f,) if isinstance(f, str) else tuple(f)
            desc = len(f) > 1 and str(f[1]).lower().startswith("desc")
            nulls_first = len(f) > 2 and str(f[2]).lower() == "first"
            order.append((self.unit.target_names.index(f[0]), desc, nulls_first))
        u = self.unit
        return AggNode(self._hdk, ir.ExecutionUnit(u.table, u.groupby_exprs, u.target_exprs, u.target_names, u.quals,
                                                   u.joins, order, u.limit))

    def run(self, device_type: str = "GPU") -> ExecutionResult:
        return RelAlgExecutor(self._hdk.executor, self._hdk.storage, self.unit).execute(device_type=device_type)


class HDK:
    """python/pyhdk/hdk.py:2113-2128"""

    def __init__(self, device: int = 0, n_devices: int = 1, **config):
        self.config = planner.Config(**config)
        self.storage = ArrowStorage(n_devices=n_devices)
        self._executor: Optional[Executor] = None
        self._device = device

    @property
    def executor(self) -> Executor:
        if self._executor is None:
            self._executor = Executor(self.storage, self.config, self._device)
        return self._executor

    def import_arrow(self, at: pa.Table, table_name: Optional[str] = None, fragment_size: Optional[int] = None,
                     shard=None, on_device: bool = False) -> QueryNode:
        """on_device: copy the raw Arrow buffers to the GPU and convert them there (NULL sentinels + chunk
        statistics, hdk_b200_materialize_nulls_on_device) instead of materialising on the host first."""
        name = table_name or f"tab_{len(self.storage.tables) + 1}"
        if on_device:
            import torch
            self.storage.import_arrow_table_to_device(at, name, torch.device("cuda", self._device), fragment_size or DEFAULT_FRAGMENT_SIZE,
                                                      shard=shard)
        else:
            self.storage.import_arrow_table(at, name, fragment_size or DEFAULT_FRAGMENT_SIZE, shard=shard)
        return QueryNode(self, name)

    def import_pydict(self, values: Dict[str, list], table_name: Optional[str] = None, **kw) -> QueryNode:
        return self.import_arrow(pa.table(values), table_name, **kw)

    def drop_table(self, table):
        name = table.table_name if isinstance(table, QueryNode) else table
        self.storage.drop_table(name)
        if self._executor is not None:
            self._executor.join_tables = {k: v for k, v in self._executor.join_tables.items() if k[0] != name}

    def scan(self, table_name: str) -> QueryNode:
        return QueryNode(self, table_name)

    def sql(self, sql_query: str, **kwargs) -> ExecutionResult:
        return RelAlgExecutor(self.executor, self.storage, sql_query).execute(device_type="GPU", **kwargs)


def init(**kwargs) -> HDK:
    """pyhdk.init() (python/pyhdk/hdk.py:2956-2960)"""
    return HDK(**kwargs)
