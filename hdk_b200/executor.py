"""Executor / RelAlgExecutor / ResultSet façade over the C ABI.

Mirrors the call structure of the reference for the hot path only:
  Executor.execute_work_unit      Executor::executeWorkUnit  (QE/Execute.cpp:1736-1798): build the memory
                                  descriptor, fetch chunks to the device, launch, collect, reduce
  QueryExecutionContext           QE/QueryExecutionContext.cpp:238-550 (launchGpuCode): kernel params,
                                  group-by buffer, launch, error code, copy back
  RelAlgExecutor.execute          QE/RelAlgExecutor.cpp:158-213 + the out-of-slots → bigger table retry
                                  ladder (:1544-1566); python/pyhdk/_sql.pyx:169-213
  ResultSet / ExecutionResult     omniscidb/ResultSet (iteration, Arrow conversion), python/pyhdk/_sql.pyx:80-83
PyTorch is used only as the device allocator / stream provider and for torch.distributed.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import pyarrow as pa

from . import _lib, abi, ir, planner, sql
from .storage import ArrowStorage, Table


# in-band error codes of a launch (ERROR_CODE; > 0 persistent): the reference's (QE/ErrorHandling.h ERR_DIV_BY_ZERO,
# ERR_OVERFLOW_OR_UNDERFLOW) and the library's own (include/hdk_b200.h)
ERROR_TEXT = {1: "division by zero", 7: "overflow or underflow", 1003: "group key outside the range of the perfect-hash layout",
              1004: "a peer's partial table never arrived", 1005: "baseline hash: a claimed entry was never published"}


class QueryError(RuntimeError):
    """In-band query error (QE/Execute.h:1019-1031 error codes)."""

    def __init__(self, code, msg):
        super().__init__(f"query failed with error code {code}: {msg}")
        self.code = code


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.HdkB200Error("no CUDA device visible: hdk_b200 has no CPU fallback")
    return torch


class DeviceContext:
    """Per-GPU state: chunk cache (DataMgr GPU level stand-in), scratch, stream."""

    def __init__(self, device: int = 0, hot_data: bool = True):
        torch = _torch()
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.index = device
        self.scratch = None
        self.hot_data = hot_data        # keep chunks resident between queries (USE_HOT_DATA in the reference's taxi bench)
        self.h2d_bytes = 0              # bytes copied host → device by fetches (e2e accounting)
        self._staging = {}
        self._batch_fresh = None

    def stream_ptr(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def upload(self, arr: np.ndarray):
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1))
        return t.to(self.device, non_blocking=False)

    def chunk(self, frag, col: str):
        """Executor::fetchChunks (QE/ExecutionKernel.cpp:205-228): chunk → device.  hot_data keeps the device
        copy; otherwise every query copies the chunk again from (pinned) host memory into a reusable staging
        buffer, like a cold DataMgr GPU level."""
        d = frag.device_chunks.get(col)
        if d is not None:
            return d
        if self.hot_data:
            d = self.upload(frag.chunks[col])
            frag.device_chunks[col] = d
            return d
        src = frag.pinned.get(col) if getattr(frag, "pinned", None) else None
        if src is None:
            src = self.torch.from_numpy(np.ascontiguousarray(frag.chunks[col]).view(np.uint8).reshape(-1).copy())
        key = (frag.frag_id, col)
        dst = self._staging.get(key)
        if dst is not None and self._batch_fresh is not None and key in self._batch_fresh:
            return dst   # already copied by an earlier query of this batch
        if dst is None or dst.numel() != src.numel():
            dst = self.torch.empty(src.numel(), dtype=self.torch.uint8, device=self.device)
            self._staging[key] = dst
        dst.copy_(src, non_blocking=True)
        self.h2d_bytes += src.numel()
        if self._batch_fresh is not None:
            self._batch_fresh.add(key)
        return dst

    @contextlib.contextmanager
    def batch(self):
        """Queries run inside one batch share chunk copies: with hot_data off a chunk is copied host → device once
        per batch instead of once per query (the reference's DataMgr GPU pool keeps chunks across queries for good,
        BufferMgr::getBuffer; a batch bounds that residency so that every batch still pays for its own input)."""
        prev, self._batch_fresh = self._batch_fresh, set()
        try:
            yield self
        finally:
            self._batch_fresh = prev

    def get_scratch(self, nbytes: int):
        if self.scratch is None or self.scratch.numel() < nbytes:
            self.scratch = self.torch.empty(max(nbytes, 1 << 20), dtype=self.torch.uint8, device=self.device)
        return self.scratch


@dataclass
class JoinTable:
    """PerfectJoinHashTable (JHT/PerfectJoinHashTable.h) over a device buffer."""
    buffer: object           # torch uint8 tensor holding int32 entries
    hash_type: str           # "OneToOne" | "OneToMany"
    min_key: int
    max_key: int
    entry_count: int
    inner_table: Table
    inner_columns_dev: Dict[str, object]
    bitmap: object = None    # presence bitmap (uint32 words) once a slot-ordered column has been made
    by_slot: Dict[str, object] = None   # inner column name -> copy ordered by hash slot
    dense: bool = False      # OneToOne and every slot occupied
    key_width: int = 0       # baseline join tables ("Baseline*"): bytes per key component / payload cell
    n_keys: int = 1


class ResultSet:
    """A group-by buffer + its descriptor, iterable like omniscidb/ResultSet/ResultSet.h."""

    def __init__(self, planned: planner.PlannedQuery, buffer: np.ndarray, dictionaries=None):
        self.planned = planned
        self.buffer = buffer
        self.dictionaries = dictionaries or {}
        self._cols = None
        self.sorted_on_device = False     # ORDER BY / LIMIT already applied by hdk_b200_sort_permutation + gather_rows

    def _decode(self):
        """ResultSet iteration on the host (ResultSetIteration.cpp:1264-1360): numpy restatement used for
        small results; large results go through hdk_b200_compact_result on the device."""
        if self._cols is not None:
            return self._cols
        if getattr(self, "_dev_cells", None) is not None:
            self._cols = ResultSet.from_compact(self.planned, self._dev_cells.cpu().numpy(), self.dictionaries)._cols
            return self._cols
        pq, q, p = self.planned, self.planned.qmd, self.planned.plan
        E = q.entry_count
        buf = self.buffer
        a8 = lambda x: (x + 7) & ~7  # noqa: E731
        # layout
        if q.output_columnar:
            off = 0 if q.keyless else q.key_count * a8(8 * E)
            slot_off = []
            for s in range(q.slot_count):
                slot_off.append(off)
                off += a8(q.slot_padded[s] * E)
            row_bytes = 0
            key_bytes = 0
        else:
            key_bytes = 0 if q.keyless else a8(q.key_count * q.key_width)
            off, slot_off = 0, []
            for s in range(q.slot_count):
                if q.slot_padded[s] == 8:
                    off = a8(off)
                slot_off.append(off)
                off += q.slot_padded[s]
            row_bytes = a8(key_bytes + off)

        def read_slot(s, width=None):
            w = width or q.slot_padded[s]
            dt = np.int64 if w == 8 else np.int32
            if q.output_columnar:
                return np.frombuffer(buf, dtype=dt, count=E, offset=slot_off[s]).astype(np.int64) if w == q.slot_padded[s] \
                    else np.frombuffer(buf, dtype=np.int64, count=E, offset=slot_off[s]).astype(np.int32).astype(np.int64)
            rows = np.frombuffer(buf, dtype=np.uint8).reshape(E, row_bytes)
            raw = rows[:, key_bytes + slot_off[s]: key_bytes + slot_off[s] + w]
            return np.ascontiguousarray(raw).view(dt).reshape(E).astype(np.int64)

        def read_key(k):
            if q.output_columnar:
                return np.frombuffer(buf, dtype=np.int64, count=E, offset=k * a8(8 * E)).copy()
            rows = np.frombuffer(buf, dtype=np.uint8).reshape(E, row_bytes)
            w = q.key_width
            raw = np.ascontiguousarray(rows[:, k * w:(k + 1) * w])
            return raw.view(np.int64 if w == 8 else np.int32).reshape(E).astype(np.int64)

        # non-empty entries (ResultSetStorage::isEmptyEntry)
        if q.key_count == 0:
            valid = np.ones(E, dtype=bool)     # non-grouped aggregate: the single row is always returned (COUNT 0 / NULL sums)
        elif q.keyless:
            s = q.target_idx_for_key
            init = q.init_vals[s]
            if q.slot_padded[s] == 4:
                init = int(np.int32(init & 0xFFFFFFFF if init >= 0 else init))
            valid = read_slot(s) != init
        else:
            first = read_key(0)
            valid = first != (abi.EMPTY_KEY_32 if (q.key_width == 4 and not q.output_columnar) else abi.EMPTY_KEY_64)
        cols = []
        for t, ti in enumerate(pq.infos):
            tg = p.targets[t]
            chosen = ti.compact_type
            if ti.agg == abi.AGG_AVG:
                ssum = read_slot(tg.slot, 4 if ti.float_argument_input else None)
                cnt = read_slot(tg.slot + 1)
                if ti.float_argument_input:
                    dividend = ssum.astype(np.int32).view(np.float32).astype(np.float64)
                elif ti.type.is_fp:
                    dividend = ssum.view(np.float64)
                else:
                    dividend = ssum.astype(np.float64)
                with np.errstate(divide="ignore", invalid="ignore"):
                    v = dividend / cnt.astype(np.float64)
                col = np.ma.array(v, mask=(cnt == 0))
            else:
                if tg.slot >= 0 and q.slot_padded[tg.slot]:
                    f4 = ti.float_argument_input
                    raw = read_slot(tg.slot, 4 if f4 else None)
                else:
                    f4 = False
                    raw = read_key(ti.key_index)
                is_fp = chosen.is_fp and ti.agg != abi.AGG_COUNT
                if is_fp:
                    v = raw.astype(np.int32).view(np.float32).astype(np.float64) if f4 else raw.view(np.float64)
                    null = np.float64(np.float32(abi.FLT_MIN)) if chosen.width == 4 else abi.DBL_MIN
                    col = np.ma.array(v, mask=(v == null) & chosen.nullable)     # ResultSet::isNull: a NOT NULL type never is
                else:
                    w = chosen.width
                    trunc = raw if w == 8 else raw.astype({1: np.int8, 2: np.int16, 4: np.int32}[w]).astype(np.int64)
                    col = np.ma.array(raw, mask=(trunc == abi.int_null(w)) & chosen.nullable)
            cols.append(col[valid])
        self._cols = cols
        return cols

    @classmethod
    def from_compact(cls, planned: planner.PlannedQuery, cells: np.ndarray, dictionaries=None) -> "ResultSet":
        """Result set over the output of hdk_b200_compact_result: cells[t] holds one 8-byte cell per non-empty
        entry (integers with NULL normalised to the target type's sentinel, doubles as bits, AVG finalised)."""
        rs = cls(planned, np.zeros(0, dtype=np.uint8), dictionaries)
        cols = []
        for t, ti in enumerate(planned.infos):
            c = np.ascontiguousarray(cells[t])
            chosen = ti.compact_type
            if ti.agg == abi.AGG_AVG:
                v = c.view(np.float64)
                cols.append(np.ma.array(v, mask=(v == abi.DBL_MIN)))
            elif chosen.is_fp and ti.agg != abi.AGG_COUNT:
                v = c.view(np.float64)
                null = np.float64(np.float32(abi.FLT_MIN)) if (ti.float_argument_input or chosen.width == 4) else abi.DBL_MIN
                cols.append(np.ma.array(v, mask=(v == null) & chosen.nullable))
            else:
                cols.append(np.ma.array(c, mask=(c == abi.int_null(ti.type.width)) & chosen.nullable))
        rs._cols = cols
        return rs

    @classmethod
    def from_compact_device(cls, planned: planner.PlannedQuery, cells_dev, dictionaries, executor) -> "ResultSet":
        """Result set over compacted cells that STAY on the device (torch int64 [n_targets, rows]): to_arrow() builds the
        typed Arrow value buffers and validity bitmaps there (hdk_b200_arrow_column_on_device) and copies only those;
        Python-side access (row_count, _decode) copies the cells lazily."""
        rs = cls(planned, np.zeros(0, dtype=np.uint8), dictionaries)
        rs._dev_cells, rs._executor = cells_dev, executor
        return rs

    def _arrow_from_device(self) -> pa.Table:
        """ArrowResultSetConverter::convertToArrowTable (omniscidb/ResultSet/ArrowResultSetConverter.cpp) without the host loop
        over rows: per target one kernel writes the Arrow value buffer + validity bitmap, pa.Array.from_buffers wraps them."""
        ex, cells = self._executor, self._dev_cells
        torch, lib, st = ex.ctx.torch, ex.lib, ex.ctx.stream_ptr()
        n = int(cells.shape[1])
        unit = self.planned.unit
        arrays, names = [], []
        words = (n + 31) // 32
        for t, ti in enumerate(self.planned.infos):
            chosen, typ = ti.compact_type, ti.type
            post = None
            if ti.agg == abi.AGG_AVG:
                spec = (1, 1, 8, 1, 0, abi.DBL_MIN, pa.float64())
            elif chosen.is_fp and ti.agg != abi.AGG_COUNT:
                f4 = ti.float_argument_input or chosen.width == 4
                null_fp = float(np.float64(np.float32(abi.FLT_MIN))) if f4 else abi.DBL_MIN
                narrow = ti.agg in (abi.AGG_SUM, abi.AGG_MIN, abi.AGG_MAX) and typ.is_fp and typ.width == 4
                spec = (1, 1, 4 if narrow else 8, int(chosen.nullable), 0, null_fp, pa.float32() if narrow else pa.float64())
            else:
                null_int = abi.int_null(ti.type.width)
                if not ti.is_agg and typ.kind == "dict" and self.dictionaries.get(t) is not None:
                    w, pt, post = 4, pa.int32(), "dict"
                elif not ti.is_agg and typ.kind == "dict":
                    w, pt = 4, pa.int32()
                elif not ti.is_agg and typ.kind == "timestamp":
                    w, pt, post = 8, pa.int64(), "timestamp"
                elif ti.agg == abi.AGG_COUNT:
                    w = 4 if typ.width == 4 else 8
                    pt = pa.int32() if w == 4 else pa.int64()
                else:
                    w = 8 if (ti.is_agg and ti.agg == abi.AGG_SUM) else typ.width
                    pt = {1: pa.int8(), 2: pa.int16(), 4: pa.int32(), 8: pa.int64()}[w]
                spec = (0, 0, w, int(chosen.nullable), null_int, 0.0, pt)
            cells_fp, out_fp, w, nullable, null_int, null_fp, pt = spec
            values = torch.empty(max(n, 1) * w, dtype=torch.uint8, device=ex.ctx.device)
            validity = torch.empty(max(words, 1), dtype=torch.int32, device=ex.ctx.device) if nullable else None
            nulls = torch.zeros(1, dtype=torch.int64, device=ex.ctx.device)
            _lib.check(lib.hdk_b200_arrow_column_on_device(cells[t].data_ptr(), n, cells_fp, out_fp, w, nullable, null_int, null_fp,
                                                           values.data_ptr(), validity.data_ptr() if nullable else None, nulls.data_ptr(), st),
                       "arrow_column_on_device")
            null_count = int(nulls.item()) if nullable else 0
            vbuf = pa.py_buffer(values[: n * w].cpu().numpy())
            bbuf = pa.py_buffer(validity[:words].cpu().numpy()) if (nullable and null_count) else None
            arr = pa.Array.from_buffers(pt, n, [bbuf, vbuf], null_count)
            if post == "dict":
                arr = pa.DictionaryArray.from_arrays(arr, pa.array(self.dictionaries[t], type=pa.string())).cast(pa.string())
            elif post == "timestamp":
                arr = arr.cast(pa.timestamp({1: "s", 1000: "ms", 1000000: "us", 1000000000: "ns"}[typ.unit]))
            arrays.append(arr)
            names.append(unit.target_names[t])
        return pa.table(arrays, names=names)

    def order_entries(self) -> list:
        """One dict per ORDER BY item, typed the way the compacted 8-byte cells are (and the way `from_compact` reads them)."""
        out = []
        for entry in self.planned.unit.order_by:
            t, desc, nulls_first = (tuple(entry) + (False,))[:3]
            ti = self.planned.infos[t]
            chosen = ti.compact_type
            is_fp = ti.agg == abi.AGG_AVG or (chosen.is_fp and ti.agg != abi.AGG_COUNT)
            if ti.agg == abi.AGG_AVG:
                width = 8
            elif is_fp:
                width = 4 if (ti.float_argument_input or chosen.width == 4) else 8
            else:
                width = ti.type.width
            d = self.dictionaries.get(t) if (not ti.is_agg and ti.type.kind == "dict") else None
            nullable = 1 if ti.agg == abi.AGG_AVG else int(chosen.nullable)
            out.append(dict(column=t, is_fp=int(is_fp), type_width=width, nullable=nullable, is_desc=int(bool(desc)),
                            nulls_first=int(bool(nulls_first)), dictionary=d))
        return out

    def row_count(self):
        cols = self._decode()
        return len(cols[0]) if cols else 0

    def to_arrow(self) -> pa.Table:
        """ArrowResultSetConverter::convertToArrowTable (omniscidb/ResultSet/ArrowResultSetConverter.cpp)."""
        unit = self.planned.unit
        if getattr(self, "_dev_cells", None) is not None and (self.sorted_on_device or not unit.order_by):
            tbl = self._arrow_from_device()
            if unit.limit is not None:
                tbl = tbl.slice(0, unit.limit)
            if unit.n_hidden:
                tbl = tbl.select(list(range(tbl.num_columns - unit.n_hidden)))
            return tbl
        cols = self._decode()
        arrays, names = [], []
        for t, (ti, col) in enumerate(zip(self.planned.infos, cols)):
            mask = np.ma.getmaskarray(col)
            data = np.ma.getdata(col)
            name = unit.target_names[t]
            typ = ti.type
            if not ti.is_agg and typ.kind == "dict":
                d = self.dictionaries.get(t)
                arr = pa.array([None if m else d[int(v)] for v, m in zip(data, mask)], type=pa.string()) if d is not None \
                    else pa.array(data.astype(np.int32), mask=mask)
            elif data.dtype == np.float64:
                if ti.agg in (abi.AGG_SUM, abi.AGG_MIN, abi.AGG_MAX) and typ.is_fp and typ.width == 4:
                    arr = pa.array(data.astype(np.float32), mask=mask)
                else:
                    arr = pa.array(data, mask=mask)
            else:
                w = 8 if ti.is_agg and ti.agg in (abi.AGG_SUM,) else typ.width
                if not ti.is_agg and typ.kind == "timestamp":
                    u = {1: "s", 1000: "ms", 1000000: "us", 1000000000: "ns"}[typ.unit]
                    arr = pa.array(data, mask=mask, type=pa.int64()).cast(pa.timestamp(u))
                elif ti.agg == abi.AGG_COUNT:
                    arr = pa.array(data.astype(np.int32 if typ.width == 4 else np.int64), mask=mask)
                else:
                    arr = pa.array(data.astype({1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[w]), mask=mask)
            arrays.append(arr)
            names.append(name)
        tbl = pa.table(arrays, names=names)
        if unit.order_by and not self.sorted_on_device:
            # only result sets that did not come through Executor.execute_work_unit (tests decoding an oracle buffer):
            # stable single-key sorts, last key first, each with its own NULL placement
            import pyarrow.compute as pc
            for entry in reversed(unit.order_by):
                i, d, nf = (tuple(entry) + (False,))[:3]
                idx = pc.sort_indices(tbl, sort_keys=[(names[i], "descending" if d else "ascending")],
                                      null_placement="at_start" if nf else "at_end")
                tbl = tbl.take(idx)
        if unit.limit is not None:
            tbl = tbl.slice(0, unit.limit)
        if unit.n_hidden:
            tbl = tbl.select(list(range(len(names) - unit.n_hidden)))
        return tbl


class ExecutionResult:
    """python/pyhdk/_sql.pyx:62-100"""

    def __init__(self, result_set: ResultSet, launch_info=None):
        self.result_set = result_set
        self.launch_info = launch_info

    def to_arrow(self):
        return self.result_set.to_arrow()

    def df(self):
        return self.to_arrow().to_pandas()

    def row_count(self):
        return self.result_set.row_count()


class Executor:
    """One executor per process / GPU (Executor::getExecutor, QE/Execute.cpp:403)."""

    def __init__(self, storage: ArrowStorage, config: Optional[planner.Config] = None, device: int = 0,
                 hot_data: bool = True):
        self.storage = storage
        self.config = config or planner.Config()
        self.ctx = DeviceContext(device, hot_data)
        self.compact_threshold_bytes = 64 << 20   # larger group-by buffers are iterated on the device (hdk_b200_compact_result)
        # Baseline-hash tables beyond this many bytes get their rows regrouped by table region first (0 = never).
        # Off by default: measured on config 4 (250 M rows, 25 M groups, 134 regions) the two extra passes cost more
        # than the L2-resident aggregate saves with the current direct scatter (39 ms vs 27 ms) — see DESIGN.md §3.
        self.region_pass_threshold_bytes = 0
        self.region_bytes = 12 << 20                   # target size of one region (a fraction of L2)
        self.partition_over_peer_memory = True    # execute_partitioned: scatter straight into the owners' buffers (else NCCL all-to-all)
        self._peer_rows = {}
        # Under torch.distributed (one process per GPU) a query over a SHARDED table (Table.shard) runs across all ranks and
        # returns the merged result on every rank: perfect hash = partials merged inside the kernels over peer memory (or
        # NCCL all-reduce), baseline hash = rows re-partitioned by key hash, join build sides built once and broadcast.
        self.distributed = True
        self.broadcast_join_build = True          # False: every rank builds its own copy (what the reference does per device)
        self._exchanges = {}                      # PeerExchange per work-table shape
        self._peer_ok = None
        self._global_stats = {}
        self.lib = _lib.lib()
        self.join_tables: Dict[tuple, JoinTable] = {}
        self.last_launch_info = None

    # -- join hash tables ----------------------------------------------------------------
    def _join_columns(self, inner: Table, key_cols, key_width=None):
        """JoinColumn / JoinColumnTypeInfo PODs over the inner table's (device) key chunks."""
        n = len(key_cols)
        jcs, tis, keep = (abi.JoinColumn * n)(), (abi.JoinColumnTypeInfo * n)(), []
        rows = 0
        for k, key_col in enumerate(key_cols):
            ci = inner.columns[key_col]
            lo, hi, _ = inner.join_key_range(key_col)
            if lo is None or ci.type.is_fp:
                raise planner.UnsupportedPlan("join key without an integer range")
            chunks = (abi.JoinChunk * len(inner.fragments))()
            row = 0
            for i, f in enumerate(inner.fragments):
                d = self.ctx.chunk(f, key_col)
                keep.append(d)
                chunks[i].col_buff = d.data_ptr()
                chunks[i].num_elems = f.num_rows
                chunks[i].row_id = row
                row += f.num_rows
            rows = row
            host_chunks = np.frombuffer(bytes(chunks), dtype=np.uint8)
            dchunks = self.ctx.upload(host_chunks)
            keep.append(dchunks)
            jcs[k] = abi.JoinColumn(dchunks.data_ptr(), len(host_chunks), len(inner.fragments), row, ci.phys_width)
            # (a days-encoded date key joins on its stored days on both sides, ir.join_key_for: plain SIGNED elements)
            tis[k] = abi.JoinColumnTypeInfo(ci.phys_width, lo, hi, abi.int_null(ci.phys_width), 0, 0, abi.SIGNED)
        return jcs, tis, rows, keep

    def build_baseline_join_table(self, inner: Table, key_cols, key_width: int) -> JoinTable:
        """BaselineJoinHashTable::reify (JHT/BaselineJoinHashTable.cpp:256-259: 2 x tuples entries): the one-to-one
        layout E x (components ‖ row id) first; a duplicate composite key (err = -1, NeedsOneToManyHash) rebuilds as the
        one-to-many layout — composite-key dictionary E x components, then offsets | counts | payload."""
        cache_key = (inner.name, id(inner), tuple(key_cols), key_width)   # id(): a table re-imported under the same name is another table
        if cache_key in self.join_tables:
            return self.join_tables[cache_key]
        if self._broadcasts_build(inner):
            from . import distributed as D
            jt = self._build_baseline_join_table_local(inner, key_cols, key_width) if D.rank() == 0 else None
            jt = self._broadcast_join_table(jt, inner)
            self.join_tables[cache_key] = jt
            return jt
        jt = self._build_baseline_join_table_local(inner, key_cols, key_width)
        self.join_tables[cache_key] = jt
        return jt

    def _build_baseline_join_table_local(self, inner: Table, key_cols, key_width: int) -> JoinTable:
        torch = self.ctx.torch
        jcs, tis, rows, keep = self._join_columns(inner, key_cols)
        entries = 2 * max(rows, 1)
        n = len(key_cols)
        buf = torch.empty(entries * (n + 1) * key_width, dtype=torch.uint8, device=self.ctx.device)
        err = torch.zeros(1, dtype=torch.int32, device=self.ctx.device)
        st = self.ctx.stream_ptr()
        _lib.check(self.lib.hdk_b200_init_baseline_hash_join_buff_on_device(buf.data_ptr(), entries, n, 1, -1, key_width, st),
                   "init_baseline_hash_join_buff")
        _lib.check(self.lib.hdk_b200_fill_baseline_hash_join_buff_on_device(buf.data_ptr(), entries, -1, 0, n, 1, err.data_ptr(), jcs, tis,
                                                                            key_width, st), "fill_baseline_hash_join_buff")
        code = int(err.item())
        hash_type = "BaselineOneToOne"
        if code == -1:
            dict_bytes = entries * n * key_width
            buf = torch.empty(dict_bytes + (2 * entries + rows) * 4, dtype=torch.uint8, device=self.ctx.device)
            err.zero_()
            _lib.check(self.lib.hdk_b200_init_baseline_hash_join_buff_on_device(buf.data_ptr(), entries, n, 0, -1, key_width, st),
                       "init_baseline_hash_join_buff")
            _lib.check(self.lib.hdk_b200_fill_baseline_hash_join_buff_on_device(buf.data_ptr(), entries, -1, 0, n, 0, err.data_ptr(),
                                                                                jcs, tis, key_width, st), "fill_baseline_hash_join_buff")
            _lib.check(self.lib.hdk_b200_fill_one_to_many_baseline_hash_table_on_device(buf.data_ptr() + dict_bytes, buf.data_ptr(),
                                                                                        entries, -1, n, jcs, tis, key_width, st),
                       "fill_one_to_many_baseline_hash_table")
            code = int(err.item())
            hash_type = "BaselineOneToMany"
        if code != 0:
            raise QueryError(code, "baseline join table build failed")
        jt = JoinTable(buf, hash_type, 0, 0, entries, inner, {})
        jt.key_width, jt.n_keys = key_width, n
        return jt

    def _broadcasts_build(self, inner: Table) -> bool:
        """Build side built once (rank 0) and broadcast?  Only for a table every rank holds in full (not sharded): the
        reference rebuilds the table on every device (JHT/PerfectJoinHashTable.cpp:331-339)."""
        from . import distributed as D
        return bool(self.distributed and self.broadcast_join_build and inner.shard is None and D.is_dist() and D.world() > 1)

    def _broadcast_join_table(self, jt: Optional[JoinTable], inner: Table) -> JoinTable:
        """rank 0's table → every rank (NCCL broadcast of the buffer; the few scalars travel as an object)."""
        import torch.distributed as dist

        from . import distributed as D
        meta = [None]
        if D.rank() == 0:
            meta = [(jt.hash_type, jt.min_key, jt.max_key, jt.entry_count, jt.dense, jt.key_width, jt.n_keys, int(jt.buffer.numel()))]
        dist.broadcast_object_list(meta, src=0)
        hash_type, lo, hi, entries, dense, kw, nk, nbytes = meta[0]
        buf = D.broadcast_tensor(jt.buffer if D.rank() == 0 else None, nbytes, self.ctx.device)
        if D.rank() != 0:
            jt = JoinTable(buf, hash_type, lo, hi, entries, inner, {})
            jt.dense, jt.key_width, jt.n_keys = dense, kw, nk
        return jt

    def build_join_table(self, inner: Table, key_col: str) -> JoinTable:
        """HashJoin::getInstance → PerfectJoinHashTable::reify (JHT/PerfectJoinHashTable.cpp:90-383):
        one-to-one first; a duplicate key (err = -1) rebuilds as one-to-many (NeedsOneToManyHash); a key range too wide
        for a perfect table (TooManyHashEntries, :139-151) falls back to the baseline table like HashJoin::getInstance."""
        cache_key = (inner.name, id(inner), key_col)
        if cache_key in self.join_tables:
            return self.join_tables[cache_key]
        if self._broadcasts_build(inner):
            from . import distributed as D
            jt = self._build_join_table_local(inner, key_col) if D.rank() == 0 else None
            jt = self._broadcast_join_table(jt, inner)
            self.join_tables[cache_key] = jt
            return jt
        jt = self._build_join_table_local(inner, key_col)
        self.join_tables[cache_key] = jt
        return jt

    def _build_join_table_local(self, inner: Table, key_col: str) -> JoinTable:
        torch = self.ctx.torch
        ci = inner.columns[key_col]
        lo, hi, has_nulls = inner.join_key_range(key_col)
        if lo is None or ci.type.is_fp:
            raise planner.UnsupportedPlan("join key without an integer range")
        entries = hi - lo + 1
        if entries > self.config.max_perfect_join_entries:
            raise planner.UnsupportedPlan("join key range too large for a perfect hash table (the planner picks the baseline table)")
        chunks = (abi.JoinChunk * len(inner.fragments))()
        keep = []
        row = 0
        for i, f in enumerate(inner.fragments):
            d = self.ctx.chunk(f, key_col)
            keep.append(d)
            chunks[i].col_buff = d.data_ptr()
            chunks[i].num_elems = f.num_rows
            chunks[i].row_id = row
            row += f.num_rows
        host_chunks = np.frombuffer(bytes(chunks), dtype=np.uint8)
        dchunks = self.ctx.upload(host_chunks)
        jc = abi.JoinColumn(dchunks.data_ptr(), len(host_chunks), len(inner.fragments), row, ci.phys_width)
        ti = abi.JoinColumnTypeInfo(ci.phys_width, lo, hi, abi.int_null(ci.phys_width), 0, 0, abi.SIGNED)
        st = self.ctx.stream_ptr()
        buf = torch.empty(entries * 4, dtype=torch.uint8, device=self.ctx.device)
        err = torch.zeros(1, dtype=torch.int32, device=self.ctx.device)
        _lib.check(self.lib.hdk_b200_init_hash_join_buff_on_device(buf.data_ptr(), entries, -1, st), "init_hash_join_buff")
        _lib.check(self.lib.hdk_b200_fill_hash_join_buff_on_device(buf.data_ptr(), -1, 0, err.data_ptr(), C.byref(jc),
                                                                   C.byref(ti), 1, st), "fill_hash_join_buff")
        hash_type = "OneToOne"
        if int(err.item()) != 0:
            buf = torch.empty((2 * entries + row) * 4, dtype=torch.uint8, device=self.ctx.device)
            _lib.check(self.lib.hdk_b200_fill_one_to_many_hash_table_on_device(buf.data_ptr(), entries, -1, C.byref(jc),
                                                                               C.byref(ti), 1, st), "fill_one_to_many")
            hash_type = "OneToMany"
        jt = JoinTable(buf, hash_type, lo, hi, entries, inner, {})
        jt.dense = hash_type == "OneToOne" and not has_nulls and row == entries
        return jt

    def _linear_inner_column(self, jt: JoinTable, cname: str):
        """Inner-table column addressed by the join's row ids: the fragments' chunks back to back, like
        ColumnFetcher::linearizeColumnFragments (QE/ColumnFetcher.cpp) — the single chunk itself when there is only one."""
        frs = jt.inner_table.fragments
        if len(frs) == 1:
            return self.ctx.chunk(frs[0], cname)
        if cname not in jt.inner_columns_dev:
            jt.inner_columns_dev[cname] = self.ctx.torch.cat([self.ctx.chunk(f, cname) for f in frs]) if frs else \
                self.ctx.torch.zeros(8, dtype=self.ctx.torch.uint8, device=self.ctx.device)   # empty table: never addressed
        return jt.inner_columns_dev[cname]

    def _slot_ordered_column(self, jt: JoinTable, cname: str):
        """hdk_b200_gather_join_payload_on_device: copy of an inner column ordered by hash slot (+ presence bitmap),
        made once per (join table, column) and cached with the table."""
        if jt.by_slot is None:
            jt.by_slot = {}
        if cname in jt.by_slot:
            return jt.by_slot[cname]
        torch = self.ctx.torch
        width = jt.inner_table.columns[cname].phys_width
        out = torch.empty(max(jt.entry_count, 1) * width, dtype=torch.uint8, device=self.ctx.device)
        if jt.bitmap is None:
            jt.bitmap = torch.empty((jt.entry_count + 31) // 32 + 1, dtype=torch.int32, device=self.ctx.device)
        bcast = self._broadcasts_build(jt.inner_table)
        if bcast:
            from . import distributed as D
        if not bcast or D.rank() == 0:
            src = self._linear_inner_column(jt, cname)
            _lib.check(self.lib.hdk_b200_gather_join_payload_on_device(jt.buffer.data_ptr(), jt.entry_count, src.data_ptr(), width,
                                                                       out.data_ptr(), jt.bitmap.data_ptr(), self.ctx.stream_ptr()),
                       "gather_join_payload")
        if bcast:
            import torch.distributed as dist
            dist.broadcast(out, src=0)
            dist.broadcast(jt.bitmap, src=0)
        jt.by_slot[cname] = out
        return out

    # -- one work unit ---------------------------------------------------------------------
    def _runs_distributed(self, table: Table) -> bool:
        from . import distributed as D
        return bool(self.distributed and table.shard is not None and D.is_dist() and D.world() > 1)

    def _global_col_stats(self, table: Table, col: str):
        """Chunk statistics over ALL ranks' shards of a table (they pick perfect vs baseline hash and the key ranges, so
        every rank must plan with the same ones).  Collective; cached per table object."""
        import torch.distributed as dist
        key = (id(table), col)
        if key not in self._global_stats:
            gathered = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, table.col_stats(col))
            los = [g[0] for g in gathered if g[0] is not None]
            his = [g[1] for g in gathered if g[1] is not None]
            self._global_stats[key] = (min(los) if los else None, max(his) if his else None, any(g[2] for g in gathered), table)
        return self._global_stats[key][:3]

    def _global_num_rows(self, table: Table) -> int:
        import torch.distributed as dist
        key = (id(table), None)
        if key not in self._global_stats:
            gathered = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, int(table.num_rows))
            self._global_stats[key] = (sum(gathered), None, None, table)
        return self._global_stats[key][0]

    def _col_stats(self, unit: ir.ExecutionUnit):
        tables = [self.storage.get_table(unit.table)] + [self.storage.get_table(j.inner_table) for j in unit.joins]
        return lambda tidx, col: (self._global_col_stats(tables[tidx], col) if self._runs_distributed(tables[tidx])
                                  else tables[tidx].col_stats(col))

    def plan(self, unit: ir.ExecutionUnit, max_groups_buffer_entry_count=None, output_columnar=None) -> planner.PlannedQuery:
        outer = self.storage.get_table(unit.table)
        n_rows = self._global_num_rows(outer) if self._runs_distributed(outer) else outer.num_rows
        pq = planner.build_query(unit, self._col_stats(unit), n_rows, self.config,
                                 max_groups_buffer_entry_count=max_groups_buffer_entry_count,
                                 output_columnar=output_columnar)
        return pq

    def _kernel_params(self, pq: planner.PlannedQuery, outer: Table, fragments, joins: List[JoinTable]):
        torch = self.ctx.torch
        dev = self.ctx.device
        ncols = len(pq.columns)
        ptrs = np.zeros(max(len(fragments) * ncols, 1), dtype=np.uint64)
        keep = []
        for fi, f in enumerate(fragments):
            for ci, cname in enumerate(pq.columns):
                d = self.ctx.chunk(f, cname)
                keep.append(d)
                ptrs[fi * ncols + ci] = d.data_ptr()
        num_rows = np.array([f.num_rows for f in fragments] or [0], dtype=np.int64)
        d_ptrs = torch.from_numpy(ptrs.view(np.int64)).to(dev)
        d_rows = torch.from_numpy(num_rows).to(dev)
        jt_addr = np.zeros(abi.MAX_JOINS, dtype=np.int64)
        inner = np.zeros(abi.MAX_JOINS * abi.MAX_COLS, dtype=np.uint64)
        for j, jt in enumerate(joins):
            by_slot = bool(pq.plan.joins[j].payload_by_slot)
            jt_addr[j] = jt.bitmap.data_ptr() if by_slot else jt.buffer.data_ptr()
            for c, cname in enumerate(pq.inner_columns[j]):
                if by_slot:
                    inner[j * abi.MAX_COLS + c] = jt.by_slot[cname].data_ptr()
                    continue
                d = self._linear_inner_column(jt, cname)
                keep.append(d)
                inner[j * abi.MAX_COLS + c] = d.data_ptr()
        d_jt = torch.from_numpy(jt_addr).to(dev)
        d_inner = torch.from_numpy(inner.view(np.int64)).to(dev)
        kp = abi.KernelParams()
        kp.col_buffers = d_ptrs.data_ptr()
        kp.num_fragments = len(fragments)
        kp.num_rows = d_rows.data_ptr()
        kp.num_tables = 1 + len(joins)
        kp.join_hash_tables = d_jt.data_ptr()
        kp.inner_col_buffers = d_inner.data_ptr()
        kp.total_rows_hint = int(num_rows.sum())
        keep += [d_ptrs, d_rows, d_jt, d_inner]
        return kp, keep

    def prepare(self, pq: planner.PlannedQuery, fragments=None):
        """Everything launchGpuCode sets up before the launch: join tables, kernel params, buffers."""
        torch = self.ctx.torch
        unit = pq.unit
        outer = self.storage.get_table(unit.table)
        frags = outer.fragments if fragments is None else fragments
        joins = []
        for j, js in enumerate(unit.joins):
            pj = pq.plan.joins[j]
            inner_t = self.storage.get_table(js.inner_table)
            if pj.n_key_exprs >= 1:     # the planner chose a baseline join table (composite or wide-range key)
                jt = self.build_baseline_join_table(inner_t, js.inner_key_columns[:pj.n_key_exprs], pj.key_width)
                pj.one_to_many, pj.payload_by_slot, pj.entry_count = int(jt.hash_type == "BaselineOneToMany"), 0, jt.entry_count
                joins.append(jt)
                continue
            jt = self.build_join_table(inner_t, js.inner_key_column)
            pj.one_to_many = int(jt.hash_type == "OneToMany")
            pj.min_key, pj.max_key, pj.entry_count = jt.min_key, jt.max_key, jt.entry_count
            pj.payload_by_slot = 0
            if self.config.join_payload_by_slot and jt.hash_type == "OneToOne" and pq.inner_columns[j]:
                # a one-to-one table has no duplicate keys: as many non-NULL keys as entries ⇒ every slot is occupied
                pj.payload_by_slot = 2 if jt.dense else 1
            if pj.payload_by_slot:
                for cname in pq.inner_columns[j]:
                    self._slot_ordered_column(jt, cname)
            joins.append(jt)
        kp, keep = self._kernel_params(pq, outer, frags, joins)
        nbytes = self.lib.hdk_b200_buffer_size_bytes(C.byref(pq.qmd))
        out = torch.empty(nbytes, dtype=torch.uint8, device=self.ctx.device)
        bufptr = torch.tensor([out.data_ptr()], dtype=torch.int64, device=self.ctx.device)
        err = torch.zeros(1, dtype=torch.int32, device=self.ctx.device)
        kp.groupby_buf = bufptr.data_ptr()
        kp.error_codes = err.data_ptr()
        scratch_bytes = C.c_size_t(0)
        # (what the launch wants for this many rows: a large baseline-hash group-by stages one packed record per row there)
        _lib.check(self.lib.hdk_b200_launch_scratch_bytes(C.byref(pq.plan), C.byref(pq.qmd), int(kp.total_rows_hint),
                                                          C.byref(scratch_bytes)), "launch_scratch_bytes")
        scratch = self.ctx.get_scratch(scratch_bytes.value)
        keep += [bufptr, scratch]
        return dict(kp=kp, keep=keep, out=out, err=err, scratch=scratch, scratch_bytes=scratch_bytes.value)

    def launch(self, pq: planner.PlannedQuery, prep, ko: Optional[abi.KernelOptions] = None):
        """init buffer + fused kernel (asynchronous on the current stream)."""
        st = self.ctx.stream_ptr()
        info = abi.LaunchInfo()
        kp = prep["kp"]
        if pq.qmd.hash_type == abi.BASELINE_HASH:
            _lib.check(self.lib.hdk_b200_init_group_by_buffer(C.byref(pq.qmd), prep["out"].data_ptr(), st), "init_group_by_buffer")
            if self._wants_region_pass(pq, prep):
                kp = self._regroup_by_table_region(pq, prep)
        prep["err"].zero_()
        _lib.check(self.lib.hdk_b200_launch(C.byref(pq.plan), C.byref(pq.qmd), C.byref(ko) if ko is not None else None,
                                            C.byref(kp), prep["scratch"].data_ptr(), prep["scratch_bytes"], st,
                                            C.byref(info)), "launch")
        self.last_launch_info = info
        return info

    # -- locality pass for large baseline-hash tables ----------------------------------------------------------
    def _table_bytes(self, pq):
        return int(pq.qmd.entry_count) * 8 * (int(pq.plan.n_targets) + int(pq.qmd.key_count) + 2)   # buffer rows + work cells, roughly

    def _wants_region_pass(self, pq, prep) -> bool:
        """Worth regrouping the rows by table region?  Only when the table is far larger than L2 (random DRAM sectors per
        row otherwise) and the plan has no joins (the passes re-read plain outer columns)."""
        if not self.region_pass_threshold_bytes or pq.unit.joins:
            return False
        return self._table_bytes(pq) >= self.region_pass_threshold_bytes and int(prep["kp"].total_rows_hint) > 0

    def _regroup_by_table_region(self, pq, prep):
        """hdk_b200_region_count → exclusive scan → hdk_b200_region_scatter_to into one buffer per column, regions in
        order; returns kernel params over that single regrouped fragment.  Everything stays on the stream."""
        torch = self.ctx.torch
        dev = self.ctx.device
        st = self.ctx.stream_ptr()
        outer = self.storage.get_table(pq.unit.table)
        widths = [outer.columns[c].phys_width for c in pq.columns]
        n_regions = int(min(1024, max(2, -(-self._table_bytes(pq) // self.region_bytes))))
        total = int(prep["kp"].total_rows_hint)
        counts = torch.zeros(n_regions, dtype=torch.int64, device=dev)
        _lib.check(self.lib.hdk_b200_region_count(C.byref(pq.plan), C.byref(pq.qmd), C.byref(prep["kp"]), n_regions, counts.data_ptr(), st),
                   "region_count")
        offsets = torch.cumsum(counts, 0) - counts
        cursors = torch.zeros(n_regions, dtype=torch.int64, device=dev)
        cols = [torch.empty(max(total, 1) * w, dtype=torch.uint8, device=dev) for w in widths]
        dest = torch.tensor([t.data_ptr() for t in cols] * n_regions, dtype=torch.int64, device=dev)
        _lib.check(self.lib.hdk_b200_region_scatter_to(C.byref(pq.plan), C.byref(pq.qmd), C.byref(prep["kp"]), n_regions, dest.data_ptr(),
                                                       offsets.data_ptr(), cursors.data_ptr(), st), "region_scatter_to")
        ptrs = torch.tensor([t.data_ptr() for t in cols], dtype=torch.int64, device=dev)
        n_rows = counts.sum().reshape(1)                    # rows that passed the filters: stays on the device
        old = prep["kp"]
        kp = abi.KernelParams()
        C.memmove(C.byref(kp), C.byref(old), C.sizeof(kp))
        kp.col_buffers = ptrs.data_ptr()
        kp.num_fragments = 1
        kp.num_rows = n_rows.data_ptr()
        prep["keep"] += [cols, dest, ptrs, n_rows, counts, offsets, cursors]
        return kp

    # -- multi-GPU split of the perfect-hash launch ------------------------------------
    def work_table_layout(self, pq: planner.PlannedQuery) -> abi.WorkTableLayout:
        wl = abi.WorkTableLayout()
        _lib.check(self.lib.hdk_b200_work_table_layout_get(C.byref(pq.plan), C.byref(pq.qmd), C.byref(wl)), "work_table_layout")
        return wl

    def launch_exchange(self, pq: planner.PlannedQuery, prep, xchg, ko: Optional[abi.KernelOptions] = None):
        """Multi-GPU perfect-hash launch with the merge over peer memory (hdk_b200_launch_exchange): init + scan +
        publish to every rank + wait / merge / finalize, no NCCL call.  `xchg`: distributed.PeerExchange of this plan."""
        st = self.ctx.stream_ptr()
        info = abi.LaunchInfo()
        prep["err"].zero_()
        need = prep["scratch_bytes"] + 80
        if prep["scratch"].numel() < need:
            prep["scratch"] = self.ctx.get_scratch(need)
        _lib.check(self.lib.hdk_b200_launch_exchange(C.byref(pq.plan), C.byref(pq.qmd), C.byref(ko) if ko is not None else None,
                                                     C.byref(prep["kp"]), prep["scratch"].data_ptr(), prep["scratch"].numel(),
                                                     xchg.ptrs(), xchg.world, xchg.rank, xchg.next_epoch(), st, C.byref(info)),
                   "launch_exchange")
        return info

    def launch_partial(self, pq: planner.PlannedQuery, prep, ko: Optional[abi.KernelOptions] = None):
        """scan this rank's fragments into the neutral work table (asynchronous)."""
        st = self.ctx.stream_ptr()
        info = abi.LaunchInfo()
        prep["err"].zero_()
        _lib.check(self.lib.hdk_b200_init_work_table(C.byref(pq.plan), C.byref(pq.qmd), prep["scratch"].data_ptr(), st), "init_work_table")
        _lib.check(self.lib.hdk_b200_launch_partial(C.byref(pq.plan), C.byref(pq.qmd), C.byref(ko) if ko is not None else None,
                                                    C.byref(prep["kp"]), prep["scratch"].data_ptr(), st, C.byref(info)), "launch_partial")
        self.last_launch_info = info
        return info

    def finalize(self, pq: planner.PlannedQuery, prep):
        _lib.check(self.lib.hdk_b200_finalize(C.byref(pq.plan), C.byref(pq.qmd), prep["scratch"].data_ptr(),
                                              prep["out"].data_ptr(), self.ctx.stream_ptr()), "finalize")

    def execute_sharded(self, pq: planner.PlannedQuery, prep, group=None):
        """perfect hash across ranks: partial scan → NCCL all-reduce per merge class → finalize."""
        from . import distributed as D
        info = self.launch_partial(pq, prep)
        wl = self.work_table_layout(pq)
        D.allreduce_work_table(prep["scratch"], wl.n_cells, wl.sum_i64_cells, wl.sum_cells, wl.min_cells, wl.max_cells, group)
        if D.is_dist() and D.world() > 1:
            import torch.distributed as dist
            dist.all_reduce(prep["err"], op=dist.ReduceOp.MAX, group=group)   # a positive (persistent) code wins
        self.finalize(pq, prep)
        return info

    def execute_partitioned(self, unit: ir.ExecutionUnit, max_groups_buffer_entry_count: int, group=None):
        """baseline hash across ranks (model: QE/RelAlgExecutor.cpp:691-838): partition this rank's
        rows by the upper half of MurmurHash64A(key), range-reduced to world, on the device (count → scatter), exchange the partitions with an
        NCCL all-to-all, aggregate the received rows locally.  Every key lives on exactly one rank afterwards,
        so there is no merge: the result is the concatenation of the ranks' ResultSets.
        Returns (ResultSet of this rank's key partition, rows received)."""
        from . import distributed as D
        torch = self.ctx.torch
        dev = self.ctx.device
        W = D.world()
        pq = self.plan(unit, max_groups_buffer_entry_count)
        if pq.qmd.hash_type != abi.BASELINE_HASH:
            raise planner.UnsupportedPlan("execute_partitioned is for baseline-hash group-by")
        if unit.joins:
            raise planner.UnsupportedPlan("partitioned aggregation of joined plans is not supported")
        prep = self.prepare(pq)
        outer = self.storage.get_table(unit.table)
        st = self.ctx.stream_ptr()
        counts = torch.zeros(W, dtype=torch.int64, device=dev)
        _lib.check(self.lib.hdk_b200_shuffle_count(C.byref(pq.plan), C.byref(prep["kp"]), W, counts.data_ptr(), st), "shuffle_count")
        widths = [outer.columns[c].phys_width for c in pq.columns]
        if self.partition_over_peer_memory and D.is_dist():
            frag, n_recv = self._exchange_rows_over_peer_memory(pq, prep, counts, widths, group)
            prep2 = self.prepare(pq, fragments=[frag])
            self.launch(pq, prep2)
            code = int(prep2["err"].item())
            if code != 0:
                raise QueryError(code, "ran out of slots in the group-by buffer" if code < 0 else "runtime error")
            return ResultSet(pq, prep2["out"].cpu().numpy()), n_recv
        n_local = int(counts.sum().item())
        offsets = torch.cumsum(counts, 0) - counts
        cursors = torch.zeros(W, dtype=torch.int64, device=dev)
        widths = [outer.columns[c].phys_width for c in pq.columns]
        send_cols = [torch.empty(max(n_local, 1) * w, dtype=torch.uint8, device=dev) for w in widths]
        ptrs = torch.tensor([t.data_ptr() for t in send_cols], dtype=torch.int64, device=dev)
        _lib.check(self.lib.hdk_b200_shuffle_scatter(C.byref(pq.plan), C.byref(prep["kp"]), W, offsets.data_ptr(), cursors.data_ptr(),
                                                     ptrs.data_ptr(), st), "shuffle_scatter")
        recv_cols, n_recv = D.all_to_all_rows(send_cols, counts, widths, group)
        # local aggregate over the received rows: one device-resident fragment, filters already applied
        from .storage import ChunkStats, Fragment
        frag = Fragment(0, n_recv, 0, 0, {}, {c: ChunkStats(None, None, False) for c in pq.columns},
                        {c: t[: n_recv * w] for c, t, w in zip(pq.columns, recv_cols, widths)})
        prep2 = self.prepare(pq, fragments=[frag])
        self.launch(pq, prep2)
        code = int(prep2["err"].item())
        if code != 0:
            raise QueryError(code, "ran out of slots in the group-by buffer" if code < 0 else "runtime error")
        return ResultSet(pq, prep2["out"].cpu().numpy()), n_recv

    def _exchange_rows_over_peer_memory(self, pq, prep, counts, widths, group=None):
        """The all-to-all of execute_partitioned fused into the scatter kernel: counts are all-gathered (W x W int64),
        every rank derives where its rows start in each owner's receive buffer, hdk_b200_shuffle_scatter_to writes them
        there through NVLink.  Returns (fragment over this rank's received rows, rows received)."""
        import torch.distributed as dist

        from . import distributed as D
        from .storage import ChunkStats, Fragment
        torch = self.ctx.torch
        dev = self.ctx.device
        W, me = D.world(), D.rank()
        gathered = [torch.empty_like(counts) for _ in range(W)]
        if W > 1:
            dist.all_gather(gathered, counts, group=group)
        else:
            gathered = [counts]
        M = torch.stack(gathered).cpu()                     # M[r][d]: rows rank r sends to rank d
        recv_total = [int(M[:, d].sum()) for d in range(W)]
        need = max(recv_total)
        pr = self._peer_rows.get(tuple(widths))
        if pr is None or pr.capacity < need:                # collective decision: every rank sees the same M
            if pr is not None:
                pr.close()
            pr = D.PeerRows(self.lib, widths, int(need * 1.25) + 1024, dev, group)
            self._peer_rows[tuple(widths)] = pr
        offsets = torch.tensor([int(M[:me, d].sum()) for d in range(W)], dtype=torch.int64, device=dev)
        cursors = torch.zeros(W, dtype=torch.int64, device=dev)
        if W > 1:
            dist.barrier(group=group)                       # every owner is done with the previous contents of its buffers
        _lib.check(self.lib.hdk_b200_shuffle_scatter_to(C.byref(pq.plan), C.byref(prep["kp"]), W, pr.dest.data_ptr(), offsets.data_ptr(),
                                                        cursors.data_ptr(), self.ctx.stream_ptr()), "shuffle_scatter_to")
        torch.cuda.synchronize(dev)
        if W > 1:
            dist.barrier(group=group)                       # all rows have landed
        n_recv = recv_total[me]
        frag = Fragment(0, n_recv, 0, 0, {}, {c: ChunkStats(None, None, False) for c in pq.columns},
                        {c: pr.local_column(i, n_recv) for i, c in enumerate(pq.columns)})
        return frag, n_recv

    def compact_on_device(self, pq: planner.PlannedQuery, out, to_host=True):
        """hdk_b200_compact_result (ResultSet iteration on the device): [n_targets, rows] int64 cells (on the host, or
        the device tensor and the row count with to_host=False)."""
        torch = self.ctx.torch
        E, T = pq.qmd.entry_count, pq.plan.n_targets
        cols = torch.empty((T, E), dtype=torch.int64, device=self.ctx.device)
        ptrs = torch.tensor([cols[t].data_ptr() for t in range(T)], dtype=torch.int64, device=self.ctx.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=self.ctx.device)
        _lib.check(self.lib.hdk_b200_compact_result(C.byref(pq.plan), C.byref(pq.qmd), out.data_ptr(), ptrs.data_ptr(),
                                                    cnt.data_ptr(), self.ctx.stream_ptr()), "compact_result")
        n = int(cnt.item())
        return cols[:, :n].cpu().numpy() if to_host else (cols, n)

    def sort_on_device(self, cols, n_rows: int, order: list, limit=None):
        """ORDER BY [LIMIT] over compacted result columns (sortResultSet, QE/ResultSetSort.cpp:752-851):
        hdk_b200_sort_permutation, then hdk_b200_gather_rows of the first `limit` rows.  cols: device int64 [T, >= n_rows];
        order: ResultSet.order_entries().  Returns a device tensor [T, min(limit, n_rows)]."""
        torch = self.ctx.torch
        T = cols.shape[0]
        n_out = n_rows if limit is None else min(int(limit), n_rows)
        out = torch.empty((T, n_out), dtype=torch.int64, device=self.ctx.device)
        if n_rows == 0 or n_out == 0:
            return out
        entries = (abi.OrderEntry * len(order))()
        keep = []
        for e, oe in zip(entries, order):
            e.column, e.is_fp, e.type_width, e.nullable = oe["column"], oe["is_fp"], oe["type_width"], oe["nullable"]
            e.is_desc, e.nulls_first = oe["is_desc"], oe["nulls_first"]
            if oe.get("dictionary") is not None:
                d = oe["dictionary"]
                rank = np.empty(len(d), dtype=np.int32)
                rank[np.array(sorted(range(len(d)), key=d.__getitem__), dtype=np.int64)] = np.arange(len(d), dtype=np.int32)
                rk = torch.from_numpy(rank).to(self.ctx.device)
                keep.append(rk)
                e.dict_rank, e.dict_size = rk.data_ptr(), len(d)
        perm = torch.empty(n_rows, dtype=torch.int32, device=self.ctx.device)
        sb = self.lib.hdk_b200_sort_scratch_bytes(n_rows)
        scratch = torch.empty(sb, dtype=torch.uint8, device=self.ctx.device)
        in_ptrs = (C.c_void_p * abi.MAX_TARGETS)(*[cols[t].data_ptr() for t in range(T)])
        _lib.check(self.lib.hdk_b200_sort_permutation(in_ptrs, entries, len(order), n_rows, 0 if limit is None else n_out,
                                                      perm.data_ptr(), None, scratch.data_ptr(), sb, self.ctx.stream_ptr()),
                   "sort_permutation")
        out_ptrs = (C.c_void_p * T)(*[out[t].data_ptr() for t in range(T)])
        _lib.check(self.lib.hdk_b200_gather_rows(in_ptrs, out_ptrs, T, perm.data_ptr(), n_out, self.ctx.stream_ptr()), "gather_rows")
        return out

    # -- one query across all ranks (one process per GPU) ------------------------------------------------------------
    def _exchange_for(self, pq):
        """PeerExchange buffers for this plan's work table, shared by every plan of the same shape (collective on first
        use).  None when peer memory is not available on this box: the caller falls back to NCCL all-reduce."""
        import torch.distributed as dist

        from . import distributed as D
        wl = self.work_table_layout(pq)
        key = (int(wl.n_cells), int(wl.sum_i64_cells), int(wl.sum_cells), int(wl.min_cells), int(wl.max_cells))
        if self._peer_ok is False:
            return None
        if key not in self._exchanges:
            ok, x = 1, None
            try:
                x = D.PeerExchange(self.lib, pq.plan, pq.qmd, self.ctx.device)
            except Exception:      # every rank must take the same path
                ok = 0
            flag = self.ctx.torch.tensor([ok], dtype=self.ctx.torch.int32, device=self.ctx.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self._peer_ok = False
                return None
            self._peer_ok = True
            self._exchanges[key] = x
        return self._exchanges[key]

    def _agree_on_error(self, err):
        """One in-band code for all ranks: a positive (persistent) code wins over a negative (out of slots) one."""
        import torch.distributed as dist
        t = self.ctx.torch
        v = int(err.item())
        both = t.tensor([max(v, 0), max(-v, 0)], dtype=t.int32, device=self.ctx.device)
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
        pos, neg = int(both[0].item()), int(both[1].item())
        return pos if pos else -neg

    def _gather_rows(self, cols, n: int):
        """Compacted result rows of every rank, concatenated on every rank: device int64 [T, total]."""
        import torch.distributed as dist
        t = self.ctx.torch
        W = dist.get_world_size()
        counts = [None] * W
        dist.all_gather_object(counts, int(n))
        m = max(max(counts), 1)
        T = cols.shape[0]
        mine = t.zeros((T, m), dtype=t.int64, device=self.ctx.device)
        mine[:, :n] = cols[:, :n]
        parts = [t.empty_like(mine) for _ in range(W)]
        dist.all_gather(parts, mine)
        return t.cat([p_[:, :c] for p_, c in zip(parts, counts)], dim=1), sum(counts)

    def _execute_distributed(self, unit: ir.ExecutionUnit, output_columnar=None) -> ResultSet:
        """Executor::executeWorkUnitImpl across devices (QE/Execute.cpp:2656-2805 creates one kernel per device and
        reduceMultiDeviceResultSets merges on the host, :1224-1336; partitioned aggregation is RelAlgExecutor.cpp:691-838).
        Here every rank scans its shard; what merges them is picked from the plan:
          perfect hash  → neutral partial tables merged inside the kernels over peer memory (NCCL all-reduce as fallback)
          baseline hash → rows re-partitioned by key hash (count → scatter straight into the owners' buffers), aggregated
                          by their owner, the disjoint results gathered
          joins         → the build side is built once on rank 0 and broadcast (build_join_table)
        Every rank returns the same merged ResultSet."""
        from . import distributed as D
        torch = self.ctx.torch
        outer = self.storage.get_table(unit.table)
        tables = [outer] + [self.storage.get_table(j.inner_table) for j in unit.joins]
        dicts = {t: tables[e.table].columns[e.column].dictionary for t, e in enumerate(unit.target_exprs)
                 if isinstance(e, ir.ColumnRef) and e.type.kind == "dict"}
        guess = None
        n_global = self._global_num_rows(outer)
        for attempt in range(8):
            pq = self.plan(unit, guess, output_columnar)
            if pq.qmd.hash_type == abi.PERFECT_HASH:
                prep = self.prepare(pq)
                xchg = self._exchange_for(pq)
                if xchg is not None:
                    info = self.launch_exchange(pq, prep, xchg)
                else:
                    info = self.execute_sharded(pq, prep)
                code = self._agree_on_error(prep["err"])
                if code != 0:
                    raise QueryError(code, ERROR_TEXT.get(code, "runtime error"))
                if unit.order_by:
                    cols, n = self.compact_on_device(pq, prep["out"], to_host=False)
                    order = ResultSet(pq, np.zeros(0, dtype=np.uint8), dicts).order_entries()
                    rs = ResultSet.from_compact(pq, self.sort_on_device(cols, n, order, unit.limit).cpu().numpy(), dicts)
                    rs.sorted_on_device = True
                else:
                    rs = ResultSet(pq, prep["out"].cpu().numpy(), dicts)
                rs.launch_info = info
                return rs
            if unit.joins:
                raise planner.UnsupportedPlan("partitioned aggregation of joined plans is not supported")
            out, n_recv, code, info = self._partitioned_on_device(pq)
            if code < 0:
                cur = pq.qmd.entry_count
                guess = min(max(cur * 4, 2 * min(n_global, cur * 8)), max(2 * n_global, 16))
                if guess <= cur:
                    raise QueryError(code, "ran out of slots in the group-by buffer")
                continue
            if code != 0:
                raise QueryError(code, ERROR_TEXT.get(code, "runtime error"))
            cols, n = self.compact_on_device(pq, out, to_host=False)
            order = ResultSet(pq, np.zeros(0, dtype=np.uint8), dicts).order_entries() if unit.order_by else None
            if order and unit.limit is not None:       # each owner's first `limit` rows are enough for the global first `limit`
                top = self.sort_on_device(cols, n, order, unit.limit)
                cols, n = top, top.shape[1]
            allc, total = self._gather_rows(cols, n)
            if order:
                allc = self.sort_on_device(allc, total, order, unit.limit)
            rs = ResultSet.from_compact(pq, allc.cpu().numpy(), dicts)
            rs.sorted_on_device = bool(order)
            rs.launch_info = info
            return rs
        raise QueryError(-abi.ERR_OUT_OF_SLOTS, "ran out of slots after retries")

    def _partitioned_on_device(self, pq):
        """The shuffle of execute_partitioned with everything left on the device: (this rank's group-by buffer over the keys it
        owns, rows received, error code agreed by all ranks, launch info)."""
        from . import distributed as D
        torch = self.ctx.torch
        W = D.world()
        prep = self.prepare(pq)
        outer = self.storage.get_table(pq.unit.table)
        widths = [outer.columns[c].phys_width for c in pq.columns]
        counts = torch.zeros(W, dtype=torch.int64, device=self.ctx.device)
        _lib.check(self.lib.hdk_b200_shuffle_count(C.byref(pq.plan), C.byref(prep["kp"]), W, counts.data_ptr(), self.ctx.stream_ptr()),
                   "shuffle_count")
        if self.partition_over_peer_memory and self._peer_ok is not False:
            frag, n_recv = self._exchange_rows_over_peer_memory(pq, prep, counts, widths)
        else:
            from .storage import ChunkStats, Fragment
            n_local = int(counts.sum().item())
            offsets = torch.cumsum(counts, 0) - counts
            cursors = torch.zeros(W, dtype=torch.int64, device=self.ctx.device)
            send_cols = [torch.empty(max(n_local, 1) * w, dtype=torch.uint8, device=self.ctx.device) for w in widths]
            ptrs = torch.tensor([t.data_ptr() for t in send_cols], dtype=torch.int64, device=self.ctx.device)
            _lib.check(self.lib.hdk_b200_shuffle_scatter(C.byref(pq.plan), C.byref(prep["kp"]), W, offsets.data_ptr(), cursors.data_ptr(),
                                                         ptrs.data_ptr(), self.ctx.stream_ptr()), "shuffle_scatter")
            recv_cols, n_recv = D.all_to_all_rows(send_cols, counts, widths)
            frag = Fragment(0, n_recv, 0, 0, {}, {c: ChunkStats(None, None, False) for c in pq.columns},
                            {c: t[: n_recv * w] for c, t, w in zip(pq.columns, recv_cols, widths)})
        prep2 = self.prepare(pq, fragments=[frag])
        info = self.launch(pq, prep2)
        code = self._agree_on_error(prep2["err"])
        return prep2["out"], n_recv, code, info

    def execute_streamed(self, units: List[ir.ExecutionUnit]) -> List[ResultSet]:
        """Several perfect-hash queries over ONE host-resident table in a single pass over its fragments, the copy of
        fragment f + 1 overlapping the kernels of fragment f (Executor::fetchChunks per fragment, QE/ExecutionKernel.cpp:
        205-228, with the reference's one-kernel-per-fragment dispatch): a copy stream uploads each referenced column of a
        fragment once into one of two staging sets, the compute stream scans it for every query into that query's neutral
        work table (hdk_b200_launch_partial accumulates), the tables are finalised at the end — merged across ranks first
        when the table is sharded.  Host chunks should be pinned (Fragment.pinned) for the copies to be asynchronous."""
        from . import distributed as D
        torch = self.ctx.torch
        dev = self.ctx.device
        outer = self.storage.get_table(units[0].table)
        dist_run = self._runs_distributed(outer)
        pqs = []
        for u in units:
            if u.table != units[0].table or u.joins:
                raise planner.UnsupportedPlan("execute_streamed: plain group-bys over one table")
            pq = self.plan(u)
            if pq.qmd.hash_type != abi.PERFECT_HASH:
                raise planner.UnsupportedPlan("execute_streamed: perfect-hash plans")
            pqs.append(pq)
        cols = sorted({c for pq in pqs for c in pq.columns})
        width = {c: outer.columns[c].phys_width for c in cols}
        frags = outer.fragments
        max_rows = max([f.num_rows for f in frags] or [1])
        st = getattr(self, "_stream_state", None)
        if st is None or st["cols"] != cols or st["max_rows"] < max_rows:
            st = dict(cols=cols, max_rows=max_rows, copy=torch.cuda.Stream(device=dev),
                      stage=[{c: torch.empty(max_rows * width[c], dtype=torch.uint8, device=dev) for c in cols} for _ in range(2)],
                      ready=[torch.cuda.Event() for _ in range(2)], done=[torch.cuda.Event() for _ in range(2)])
            self._stream_state = st
        compute = torch.cuda.current_stream(dev)
        preps = []
        for pq in pqs:
            nbytes = self.lib.hdk_b200_buffer_size_bytes(C.byref(pq.qmd))
            sb = C.c_size_t(0)
            _lib.check(self.lib.hdk_b200_plan_check(C.byref(pq.plan), C.byref(pq.qmd), C.byref(sb)), "plan_check")
            work = torch.empty(max(sb.value, 8), dtype=torch.uint8, device=dev)
            out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            err = torch.zeros(1, dtype=torch.int32, device=dev)
            kps = []
            for slot in range(2):
                ptrs = torch.tensor([st["stage"][slot][c].data_ptr() for c in pq.columns], dtype=torch.int64, device=dev)
                kps.append(ptrs)
            _lib.check(self.lib.hdk_b200_init_work_table(C.byref(pq.plan), C.byref(pq.qmd), work.data_ptr(), compute.cuda_stream), "init_work_table")
            preps.append(dict(work=work, out=out, err=err, ptrs=kps, scratch=work))
        rows_dev = torch.tensor([f.num_rows for f in frags] or [0], dtype=torch.int64, device=dev)
        info = abi.LaunchInfo()
        for fi, f in enumerate(frags):
            slot = fi & 1
            with torch.cuda.stream(st["copy"]):
                if fi >= 2:
                    st["copy"].wait_event(st["done"][slot])          # the kernels of fragment fi - 2 are done with this staging set
                for c in cols:
                    src = f.pinned[c] if getattr(f, "pinned", None) else torch.from_numpy(np.ascontiguousarray(f.chunks[c]).view(np.uint8).reshape(-1))
                    st["stage"][slot][c][: src.numel()].copy_(src, non_blocking=True)
                    self.ctx.h2d_bytes += src.numel()
                st["ready"][slot].record(st["copy"])
            compute.wait_event(st["ready"][slot])
            for pq, prep in zip(pqs, preps):
                kp = abi.KernelParams()
                kp.col_buffers = prep["ptrs"][slot].data_ptr()
                kp.num_fragments = 1
                kp.num_rows = rows_dev.data_ptr() + 8 * fi
                kp.num_tables = 1
                kp.error_codes = prep["err"].data_ptr()
                kp.total_rows_hint = f.num_rows
                _lib.check(self.lib.hdk_b200_launch_partial(C.byref(pq.plan), C.byref(pq.qmd), None, C.byref(kp), prep["work"].data_ptr(),
                                                            compute.cuda_stream, C.byref(info)), "launch_partial")
            st["done"][slot].record(compute)
        results = []
        for pq, prep in zip(pqs, preps):
            if dist_run:
                wl = self.work_table_layout(pq)
                D.allreduce_work_table(prep["work"], wl.n_cells, wl.sum_i64_cells, wl.sum_cells, wl.min_cells, wl.max_cells)
                code = self._agree_on_error(prep["err"])
            _lib.check(self.lib.hdk_b200_finalize(C.byref(pq.plan), C.byref(pq.qmd), prep["work"].data_ptr(), prep["out"].data_ptr(),
                                                  compute.cuda_stream), "finalize")
            if not dist_run:
                code = int(prep["err"].item())
            if code != 0:
                raise QueryError(code, "runtime error")
            rs = ResultSet(pq, prep["out"].cpu().numpy(), {})
            rs.launch_info = info
            results.append(rs)
        return results

    def execute_work_unit(self, unit: ir.ExecutionUnit, output_columnar=None, ko=None) -> ResultSet:
        """Executor::executeWorkUnit with the out-of-slots retry of RelAlgExecutor::executeWorkUnit
        (QE/RelAlgExecutor.cpp:1544-1566: on ERR_OUT_OF_SLOTS re-run with 2 × the cardinality estimate)."""
        guess = None
        outer = self.storage.get_table(unit.table)
        if self._runs_distributed(outer):
            return self._execute_distributed(unit, output_columnar)
        for attempt in range(8):
            pq = self.plan(unit, guess, output_columnar)
            prep = self.prepare(pq)
            info = self.launch(pq, prep, ko)
            code = int(prep["err"].item())   # blocking copy of the error code = the reference's only sync
            if code < 0 and pq.qmd.hash_type == abi.BASELINE_HASH:
                cur = pq.qmd.entry_count
                # groups <= rows the kernel aggregates: the outer rows, or with a one-to-many join up to the join's
                # output cardinality (bounded here by outer rows x inner rows, and by what a uint32 entry count holds)
                bound = outer.num_rows
                if any(pq.plan.joins[j].one_to_many for j in range(pq.plan.n_joins)):
                    inner_rows = max([self.storage.get_table(j.inner_table).num_rows for j in unit.joins] or [1])
                    bound = min(outer.num_rows * max(inner_rows, 1), (1 << 31) - 1)
                guess = min(max(cur * 4, 2 * min(bound, cur * 8)), max(2 * bound, 16))
                if guess <= cur:
                    raise QueryError(code, "ran out of slots in the group-by buffer")
                continue
            if code != 0:
                raise QueryError(code, ERROR_TEXT.get(code, "runtime error"))
            dicts = {}
            tables = [outer] + [self.storage.get_table(j.inner_table) for j in unit.joins]
            for t, e in enumerate(unit.target_exprs):
                if isinstance(e, ir.ColumnRef) and e.type.kind == "dict":
                    dicts[t] = tables[e.table].columns[e.column].dictionary
            if unit.order_by:
                # ORDER BY [LIMIT]: compact, sort and cut on the device; only the rows of the answer travel
                cols, n = self.compact_on_device(pq, prep["out"], to_host=False)
                order = ResultSet(pq, np.zeros(0, dtype=np.uint8), dicts).order_entries()
                rs = ResultSet.from_compact_device(pq, self.sort_on_device(cols, n, order, unit.limit), dicts, self)
                rs.sorted_on_device = True
            elif prep["out"].numel() > self.compact_threshold_bytes:
                # large (baseline-hash) buffers: drop the empty entries and finalise AVG on the device; the rows stay there
                # until somebody asks for them (to_arrow builds the Arrow buffers on the device)
                cols, n = self.compact_on_device(pq, prep["out"], to_host=False)
                rs = ResultSet.from_compact_device(pq, cols[:, :n], dicts, self)
            else:
                rs = ResultSet(pq, prep["out"].cpu().numpy(), dicts)
            rs.launch_info = info
            return rs
        raise QueryError(-abi.ERR_OUT_OF_SLOTS, "ran out of slots after retries")


class RelAlgExecutor:
    """python/pyhdk/_sql.pyx:169-213: RelAlgExecutor(executor, schema/storage, query).execute()."""

    def __init__(self, executor: Executor, storage: ArrowStorage, query):
        self.executor = executor
        self.storage = storage
        self.unit = query if isinstance(query, ir.ExecutionUnit) else sql.parse(query, storage.tables,
                                                                                 executor.config.bigint_count)

    def execute(self, device_type: str = "GPU", **kwargs) -> ExecutionResult:
        if device_type != "GPU":
            raise _lib.HdkB200Error("hdk_b200 executes on the GPU only (no CPU fallback)")
        rs = self.executor.execute_work_unit(self.unit, output_columnar=kwargs.get("enable_columnar_output"))
        return ExecutionResult(rs, getattr(rs, "launch_info", None))
