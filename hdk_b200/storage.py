"""ArrowStorage façade: Arrow tables → fragments of fixed-width chunks with null sentinels and
chunk statistics — the physical input format of the hot path.

Restates the parts of omniscidb/ArrowStorage that define that format:
  * fragmenting by row count, default 32M rows   ArrowStorage.h:39, ArrowStorage.cpp:860-1040
  * null bitmap → in-band sentinel               ArrowStorageUtils.cpp:100-170
  * per-chunk min / max / has_nulls              ArrowStorage.cpp:1000-1040 (ChunkStats)
  * dictionary-encoded strings → int32 ids        ArrowStorageUtils.cpp (dict conversion)
  * every fragment stamped with a device id      ArrowStorage.cpp:365-367 (upstream: always 0;
    here: fragment i → GPU i mod n)
The rest of ArrowStorage (CSV/Parquet readers, append, schema registry) is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import pyarrow as pa

from . import abi, ir

DEFAULT_FRAGMENT_SIZE = 32_000_000


@dataclass
class ChunkStats:
    min: Optional[object]
    max: Optional[object]
    has_nulls: bool


@dataclass
class ColumnInfo:
    name: str
    type: ir.SqlType
    phys_width: int
    np_dtype: np.dtype
    dictionary: Optional[list] = None


@dataclass
class Fragment:
    frag_id: int
    num_rows: int
    row_offset: int
    device_id: int
    chunks: Dict[str, np.ndarray]
    stats: Dict[str, ChunkStats]
    device_chunks: Dict[str, object] = field(default_factory=dict)   # torch tensors, filled on demand


@dataclass
class Table:
    name: str
    columns: Dict[str, ColumnInfo]
    fragments: List[Fragment]
    num_rows: int
    shard: Optional[tuple] = None     # (rank, world): this process holds fragments i with i % world == rank of a table spread over
                                      # the ranks of a process group (one process per GPU); None = the whole table is here

    def col_stats(self, column: str):
        lo = hi = None
        hn = False
        for f in self.fragments:
            s = f.stats[column]
            hn = hn or s.has_nulls
            if s.min is None:
                continue
            lo = s.min if lo is None else min(lo, s.min)
            hi = s.max if hi is None else max(hi, s.max)
        return lo, hi, hn

    def join_key_range(self, column: str):
        """col_stats for the build side of a join: a table without any non-NULL key (empty, or all NULL) builds a table of
        one empty slot — no probe can match, the inner join yields nothing (the reference reaches the same result through
        an invalid ExpressionRange → HashJoinFail → loop join over zero rows)."""
        lo, hi, hn = self.col_stats(column)
        if lo is None and not self.columns[column].type.is_fp:
            lo = hi = 0
        return lo, hi, hn


def _arrow_type_to_sql(t: pa.DataType, nullable: bool):
    """→ (SqlType, numpy dtype, phys width)"""
    if pa.types.is_int8(t):
        return ir.SqlType("int", 1, nullable), np.dtype(np.int8)
    if pa.types.is_int16(t):
        return ir.SqlType("int", 2, nullable), np.dtype(np.int16)
    if pa.types.is_int32(t):
        return ir.SqlType("int", 4, nullable), np.dtype(np.int32)
    if pa.types.is_int64(t):
        return ir.SqlType("int", 8, nullable), np.dtype(np.int64)
    if pa.types.is_float32(t):
        return ir.SqlType("fp", 4, nullable), np.dtype(np.float32)
    if pa.types.is_float64(t):
        return ir.SqlType("fp", 8, nullable), np.dtype(np.float64)
    if pa.types.is_boolean(t):
        return ir.SqlType("bool", 1, nullable), np.dtype(np.int8)
    if pa.types.is_timestamp(t):
        unit = {"s": 1, "ms": 1000, "us": 1000000, "ns": 1000000000}[t.unit]
        return ir.SqlType("timestamp", 8, nullable, unit=unit), np.dtype(np.int64)
    if pa.types.is_date32(t):
        # HDK's default date encoding: days in 32 bits, decoded to seconds (8 bytes) in expressions
        return ir.SqlType("date", 8, nullable, date_in_days=True), np.dtype(np.int32)
    if pa.types.is_date64(t):
        return ir.SqlType("timestamp", 8, nullable, unit=1000), np.dtype(np.int64)
    if pa.types.is_dictionary(t) or pa.types.is_string(t) or pa.types.is_large_string(t):
        return ir.SqlType("dict", 4, nullable, dict_id=1), np.dtype(np.int32)
    raise NotImplementedError(f"Arrow type {t} is outside the hot path")


def _materialise(col: pa.ChunkedArray, sql_t: ir.SqlType, dt: np.dtype, dictionary):
    """null bitmap → sentinel (ArrowStorageUtils.cpp:100-170); returns a contiguous numpy array."""
    arr = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    t = arr.type
    if sql_t.kind == "dict":
        if not pa.types.is_dictionary(t):
            arr = arr.dictionary_encode()
        # re-map to the table-wide dictionary
        local = arr.dictionary.to_pylist()
        remap = np.array([dictionary.index(s) for s in local], dtype=np.int32) if local else np.zeros(0, np.int32)
        idx = arr.indices
        mask = np.asarray(idx.is_null()) if idx.null_count else None
        ids = np.asarray(idx.fill_null(0)).astype(np.int64)
        out = remap[ids].astype(np.int32) if len(remap) else np.zeros(len(ids), np.int32)
        if mask is not None:
            out[mask] = abi.int_null(4)
        return out
    if pa.types.is_boolean(t):
        vals = np.asarray(arr.fill_null(False)).astype(np.int8)
    elif pa.types.is_timestamp(t) or pa.types.is_date64(t):
        vals = np.asarray(arr.cast(pa.int64()).fill_null(0))
    elif pa.types.is_date32(t):
        vals = np.asarray(arr.cast(pa.int32()).fill_null(0))
    else:
        vals = np.asarray(arr.fill_null(0))
    out = np.ascontiguousarray(vals, dtype=dt).copy()
    if arr.null_count:
        mask = np.asarray(arr.is_null())
        if sql_t.is_fp:
            out[mask] = abi.fp_null(sql_t.width)
        else:
            out[mask] = abi.int_null(dt.itemsize)
    return out


def _stats(vals: np.ndarray, sql_t: ir.SqlType) -> ChunkStats:
    if len(vals) == 0:
        return ChunkStats(None, None, False)
    if sql_t.is_fp:
        null = np.float32(abi.FLT_MIN) if sql_t.width == 4 else abi.DBL_MIN
        nn = vals[vals != null]
    else:
        nn = vals[vals != abi.int_null(vals.dtype.itemsize)]
    has_nulls = len(nn) != len(vals)
    if len(nn) == 0:
        return ChunkStats(None, None, has_nulls)
    lo, hi = nn.min(), nn.max()
    return ChunkStats(float(lo) if sql_t.is_fp else int(lo), float(hi) if sql_t.is_fp else int(hi), has_nulls)


class ArrowStorage:
    """omniscidb/ArrowStorage/ArrowStorage.h:84-91 importArrowTable + fragment metadata."""

    def __init__(self, n_devices: int = 1):
        self.tables: Dict[str, Table] = {}
        self.n_devices = max(1, n_devices)

    def import_arrow_table(self, at: pa.Table, name: str, fragment_size: int = DEFAULT_FRAGMENT_SIZE,
                           shard: Optional[tuple] = None) -> Table:
        """shard=(rank, world): keep only fragments i with i % world == rank (one process per GPU)."""
        if name in self.tables:
            raise ValueError(f"table {name} already exists")
        cols: Dict[str, ColumnInfo] = {}
        for f in at.schema:
            sql_t, dt = _arrow_type_to_sql(f.type, f.nullable)
            dictionary = None
            if sql_t.kind == "dict":
                c = at.column(f.name)
                u = pa.compute.unique(c.cast(pa.string()) if not pa.types.is_dictionary(c.type)
                                      else c.cast(pa.dictionary(pa.int32(), pa.string())).cast(pa.string()))
                dictionary = [s for s in u.to_pylist() if s is not None]
            cols[f.name] = ColumnInfo(f.name, sql_t, dt.itemsize, dt, dictionary)
        n = at.num_rows
        frags: List[Fragment] = []
        fid = 0
        for off in range(0, max(n, 1), fragment_size):
            rows = min(fragment_size, n - off)
            if rows <= 0 and n > 0:
                break
            keep = shard is None or (fid % shard[1] == shard[0])
            if keep:
                sl = at.slice(off, rows)
                chunks, stats = {}, {}
                for cname, ci in cols.items():
                    v = _materialise(sl.column(cname), ci.type, ci.np_dtype, ci.dictionary)
                    chunks[cname] = v
                    stats[cname] = _stats(v, ci.type)
                frags.append(Fragment(fid, rows, off, fid % self.n_devices, chunks, stats))
            fid += 1
        t = Table(name, cols, frags, sum(f.num_rows for f in frags), shard)
        self.tables[name] = t
        return t

    def import_arrow_table_to_device(self, at: pa.Table, name: str, device, fragment_size: int = DEFAULT_FRAGMENT_SIZE,
                                     shard: Optional[tuple] = None) -> Table:
        """importArrowTable with the format conversion on the GPU: the raw Arrow value and validity
        buffers of every fixed-width column are copied to the device as they are; hdk_b200_materialize_nulls_on_device
        writes the NULL sentinels in place and reduces the chunk statistics in the same pass.  Dictionary / string
        and boolean columns (bit-packed or variable-width in Arrow) take the host path.  Fragments keep no host copy."""
        import ctypes as C

        import torch

        from . import _lib
        L = _lib.lib()
        if name in self.tables:
            raise ValueError(f"table {name} already exists")
        host_side = [f.name for f in at.schema if pa.types.is_boolean(f.type) or pa.types.is_dictionary(f.type)
                     or pa.types.is_string(f.type) or pa.types.is_large_string(f.type)]
        host_tab = self.import_arrow_table(at.select(host_side), name + ".__host", fragment_size, shard) if host_side else None
        self.tables.pop(name + ".__host", None)
        cols: Dict[str, ColumnInfo] = {}
        for f in at.schema:
            if f.name in host_side:
                cols[f.name] = host_tab.columns[f.name]
            else:
                sql_t, dt = _arrow_type_to_sql(f.type, f.nullable)
                cols[f.name] = ColumnInfo(f.name, sql_t, dt.itemsize, dt, None)
        n = at.num_rows
        frags: List[Fragment] = []
        stats_dev = torch.zeros(6, dtype=torch.int64, device=device)
        fid = kept = 0
        for off in range(0, max(n, 1), fragment_size):
            rows = min(fragment_size, n - off)
            if rows <= 0 and n > 0:
                break
            if shard is None or (fid % shard[1] == shard[0]):
                sl = at.slice(off, rows)
                fr = Fragment(fid, rows, off, fid % self.n_devices, {}, {})
                for cname, ci in cols.items():
                    if cname in host_side:
                        hf = host_tab.fragments[kept]
                        fr.device_chunks[cname] = torch.from_numpy(hf.chunks[cname].view(np.uint8).reshape(-1)).to(device)
                        fr.stats[cname] = hf.stats[cname]
                        continue
                    w = ci.phys_width
                    dst = torch.empty(max(rows, 1) * w, dtype=torch.uint8, device=device)
                    _lib.check(L.hdk_b200_init_chunk_stats_on_device(stats_dev.data_ptr(), None), "init_chunk_stats")
                    pos = 0
                    for piece in sl.column(cname).chunks:
                        m = len(piece)
                        if m == 0:
                            continue
                        validity, data = piece.buffers()[0], piece.buffers()[1]
                        raw = np.frombuffer(data, dtype=np.uint8)[piece.offset * w:(piece.offset + m) * w]
                        dst[pos * w:(pos + m) * w].copy_(torch.from_numpy(raw.copy() if not raw.flags.writeable else raw))
                        dv, bit0 = None, 0
                        if validity is not None and piece.null_count:
                            b0, b1 = piece.offset // 8, (piece.offset + m + 7) // 8
                            dv = torch.from_numpy(np.frombuffer(validity, dtype=np.uint8)[b0:b1].copy()).to(device)
                            bit0 = piece.offset - 8 * b0
                        _lib.check(L.hdk_b200_materialize_nulls_on_device(dst.data_ptr() + pos * w, w, int(ci.type.is_fp),
                                                                          dv.data_ptr() if dv is not None else None, bit0, m,
                                                                          stats_dev.data_ptr(), None), "materialize_nulls")
                        pos += m
                    fr.device_chunks[cname] = dst
                    s = stats_dev.cpu().numpy()
                    if ci.type.is_fp:
                        dec = lambda e: float(np.int64(e ^ ((e >> 63) & 0x7FFFFFFFFFFFFFFF)).view(np.float64))  # noqa: E731
                        lo, hi = (dec(int(s[2])), dec(int(s[3]))) if int(s[2]) != abi.EMPTY_KEY_64 else (None, None)
                    else:
                        lo, hi = (int(s[0]), int(s[1])) if int(s[0]) != abi.EMPTY_KEY_64 else (None, None)
                    fr.stats[cname] = ChunkStats(lo, hi, bool(s[4]))
                frags.append(fr)
                kept += 1
            fid += 1
        t = Table(name, cols, frags, sum(f.num_rows for f in frags), shard)
        self.tables[name] = t
        return t

    def add_device_table(self, name: str, columns: Dict[str, ColumnInfo], fragments: List[Fragment], shard: Optional[tuple] = None) -> Table:
        """Register fragments whose chunks already live on the device (synthetic benchmarks)."""
        t = Table(name, columns, fragments, sum(f.num_rows for f in fragments), shard)
        self.tables[name] = t
        return t

    def drop_table(self, name: str):
        self.tables.pop(name, None)

    def get_table(self, name: str) -> Table:
        return self.tables[name]
