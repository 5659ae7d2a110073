"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of the CPU oracle (oracle/liboracle_port.so, and oracle/_ref/liboracle_ref.so
when it has been built from the reference's sources).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import this module; hdk_b200/ never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from hdk_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def build(verbose=False):
    """(Re)build the oracle libraries with oracle/Makefile (g++ only)."""
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def _bind(lib):
    P, Q = C.POINTER(abi.Plan), C.POINTER(abi.Qmd)
    vp, i64, u64, i32, u32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_uint32
    lib.oracle_kind.restype = C.c_char_p
    lib.oracle_buffer_size_bytes.restype = C.c_size_t
    lib.oracle_buffer_size_bytes.argtypes = [Q]
    lib.oracle_init_group_by_buffer.argtypes = [Q, vp]
    lib.oracle_run_fragments.restype = i32
    lib.oracle_run_fragments.argtypes = [P, Q, vp, vp, u64, vp, vp, vp]
    lib.oracle_reduce.restype = C.c_int
    lib.oracle_reduce.argtypes = [P, Q, vp, vp, u32]
    lib.oracle_query.restype = i32
    lib.oracle_query.argtypes = [P, Q, vp, vp, u64, vp, vp, vp, C.c_int]
    lib.oracle_iterate.restype = u64
    lib.oracle_iterate.argtypes = [P, Q, vp, vp, vp, u64]
    lib.oracle_init_hash_join_buff.argtypes = [vp, i64, i32]
    lib.oracle_fill_hash_join_buff.restype = C.c_int
    lib.oracle_fill_hash_join_buff.argtypes = [vp, i32, C.c_int, C.POINTER(abi.JoinColumn),
                                               C.POINTER(abi.JoinColumnTypeInfo), i64]
    lib.oracle_fill_one_to_many_hash_table.argtypes = [vp, i64, i32, C.POINTER(abi.JoinColumn),
                                                       C.POINTER(abi.JoinColumnTypeInfo), i64]
    lib.oracle_init_baseline_hash_join_buff.argtypes = [vp, i64, C.c_size_t, C.c_int, i32, C.c_int]
    lib.oracle_fill_one_to_many_baseline_hash_table.restype = C.c_int
    lib.oracle_fill_one_to_many_baseline_hash_table.argtypes = [vp, vp, i64, i32, C.c_size_t, C.POINTER(abi.JoinColumn),
                                                                C.POINTER(abi.JoinColumnTypeInfo), C.c_int]
    lib.oracle_fill_baseline_hash_join_buff.restype = C.c_int
    lib.oracle_fill_baseline_hash_join_buff.argtypes = [vp, i64, i32, C.c_int, C.c_size_t, C.c_int,
                                                        C.POINTER(abi.JoinColumn),
                                                        C.POINTER(abi.JoinColumnTypeInfo), C.c_int]
    lib.oracle_probe_hash_join.argtypes = [vp, vp, i64, i64, i64, vp]
    lib.oracle_probe_baseline_hash_join.argtypes = [vp, vp, i64, C.c_size_t, C.c_int, i64, vp]
    lib.oracle_murmur3.restype = u32
    lib.oracle_murmur3.argtypes = [vp, i64, u32]
    lib.oracle_murmur1.restype = u32
    lib.oracle_murmur1.argtypes = [vp, C.c_int, u32]
    lib.oracle_murmur64a.restype = u64
    lib.oracle_murmur64a.argtypes = [vp, C.c_int, u64]
    lib.oracle_extract_year.restype = i64
    lib.oracle_extract_year.argtypes = [i64]
    lib.oracle_get_group_value.restype = vp
    lib.oracle_get_group_value.argtypes = [vp, u32, vp, u32, u32, u32]
    lib.oracle_host_read_gbs.restype = C.c_double
    lib.oracle_host_read_gbs.argtypes = [C.c_size_t, C.c_int, C.c_int]
    lib.oracle_key_hash.restype = u32
    lib.oracle_key_hash.argtypes = [vp, u32, u32]
    return lib


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "liboracle_ref.so"))


def lib(kind="port"):
    """kind: 'port' (restatement) or 'reference' (reference runtime compiled from its sources)."""
    if kind not in _libs:
        path = os.path.join(_HERE, "liboracle_port.so" if kind == "port" else "_ref/liboracle_ref.so")
        if not os.path.exists(path):
            if kind == "port":
                build()
            else:
                raise FileNotFoundError(path)
        _libs[kind] = _bind(C.CDLL(path))
        assert _libs[kind].oracle_kind().decode() == kind
    return _libs[kind]


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Fragments:
    """Host chunks laid out like ColumnFetcher hands them to a kernel: col_buffers[frag][col]."""

    def __init__(self, frags, num_rows=None):
        # frags: list of list of contiguous numpy arrays (one per plan column); num_rows: rows per fragment (needed when
        # the query reads no column at all, e.g. SELECT COUNT(*) FROM t WHERE 1 = 1 — fragment metadata in the reference)
        self.frags = [[np.ascontiguousarray(c) for c in f] for f in frags]
        self.n_frag = len(self.frags)
        self.n_cols = len(self.frags[0]) if self.frags else 0
        self.ptrs = np.array([c.ctypes.data for f in self.frags for c in f] or [0], dtype=np.uint64)
        self.num_rows = np.array([len(f[0]) if f else 0 for f in self.frags] if num_rows is None else list(num_rows), dtype=np.int64)


def run_query(planned, frags: Fragments, join_tables=None, inner_cols=None, n_threads=1, kind="port",
              per_fragment=True):
    """Execute a PlannedQuery on the CPU oracle.  per_fragment=True follows the reference's CPU
    shape (one kernel + private buffer per fragment, then reduce); False runs one kernel over all
    fragments into a single buffer.  Returns (buffer: np.uint8[], error_code)."""
    L = lib(kind)
    plan, qmd = planned.plan, planned.qmd
    nbytes = L.oracle_buffer_size_bytes(C.byref(qmd))
    buf = np.zeros(nbytes, dtype=np.uint8)
    jt = None
    if join_tables:
        jt = np.array([t.ctypes.data for t in join_tables], dtype=np.int64)
    ic = None
    if inner_cols:
        ic = np.zeros(abi.MAX_JOINS * abi.MAX_COLS, dtype=np.uint64)
        for j, cols in enumerate(inner_cols):
            for c, a in enumerate(cols):
                ic[j * abi.MAX_COLS + c] = a.ctypes.data
    if per_fragment:
        err = L.oracle_query(C.byref(plan), C.byref(qmd), _ptr(frags.ptrs), _ptr(frags.num_rows), frags.n_frag,
                             _ptr(jt), _ptr(ic), _ptr(buf), n_threads)
    else:
        L.oracle_init_group_by_buffer(C.byref(qmd), _ptr(buf))
        err = L.oracle_run_fragments(C.byref(plan), C.byref(qmd), _ptr(frags.ptrs), _ptr(frags.num_rows),
                                     frags.n_frag, _ptr(jt), _ptr(ic), _ptr(buf))
    return buf, err


def iterate(planned, buf, kind="port"):
    """ResultSet iteration: → (vals int64[n, n_targets], nulls bool[n, n_targets])."""
    L = lib(kind)
    plan, qmd = planned.plan, planned.qmd
    E, T = qmd.entry_count, plan.n_targets
    vals = np.zeros((E, max(T, 1)), dtype=np.int64)
    nulls = np.zeros((E, max(T, 1)), dtype=np.uint8)
    b = np.ascontiguousarray(buf)
    n = L.oracle_iterate(C.byref(plan), C.byref(qmd), _ptr(b), _ptr(vals), _ptr(nulls), E)
    return vals[:n], nulls[:n].astype(bool)


def rows_to_columns(planned, vals, nulls):
    """Decode iterate() output into per-target numpy columns (fp targets → float64)."""
    out = []
    for t, ti in enumerate(planned.infos):
        col = vals[:, t].copy()
        is_fp = ti.agg == abi.AGG_AVG or (ti.compact_type.is_fp and ti.agg != abi.AGG_COUNT)
        if is_fp:
            col = col.view(np.float64)
        out.append(np.ma.array(col, mask=nulls[:, t]))
    return out


def make_join_column(chunks, elem_sz):
    """chunks: list of contiguous numpy arrays (one per inner fragment)."""
    arr = (abi.JoinChunk * len(chunks))()
    row = 0
    for i, c in enumerate(chunks):
        arr[i].col_buff = c.ctypes.data
        arr[i].num_elems = len(c)
        arr[i].row_id = row
        row += len(c)
    jc = abi.JoinColumn()
    jc.col_chunks_buff = C.addressof(arr)
    jc.col_chunks_buff_sz = C.sizeof(arr)
    jc.num_chunks = len(chunks)
    jc.num_elems = row
    jc.elem_sz = elem_sz
    jc._keep = (arr, chunks)
    return jc


def make_type_info(elem_sz, min_val, max_val, null_val, uses_bw_eq=False, translated_null_val=0,
                   column_type=abi.SIGNED):
    ti = abi.JoinColumnTypeInfo()
    ti.elem_sz, ti.min_val, ti.max_val, ti.null_val = elem_sz, min_val, max_val, null_val
    ti.uses_bw_eq, ti.translated_null_val, ti.column_type = int(uses_bw_eq), translated_null_val, column_type
    return ti


# ---------------------------------------------------------------------------------------------
# ORDER BY / LIMIT over decoded result rows (test infrastructure, like everything in this file)
# ---------------------------------------------------------------------------------------------
def result_set_less(order, cols, lhs, rhs):
    """ResultSetComparator::operator() (QE/ResultSetSort.cpp:333-470) over decoded 8-byte result cells.
    order: list of dicts {column, is_fp, type_width, nullable, is_desc, nulls_first, dictionary (list of str | None)};
    cols[c][row] = int (NULL = the type's sentinel) or float (NULL = DBL_MIN / float(FLT_MIN))."""
    import struct
    for oe in order:
        l, r = cols[oe["column"]][lhs], cols[oe["column"]][rhs]
        if oe["is_fp"]:
            null = struct.unpack("<f", struct.pack("<I", 0x00800000))[0] if oe["type_width"] == 4 else 2.2250738585072014e-308
        else:
            null = -(1 << (8 * oe["type_width"] - 1))
        ln, rn = bool(oe["nullable"]) and l == null, bool(oe["nullable"]) and r == null
        if ln and rn:                       # :410-413
            continue
        if ln:                              # :414-417
            return bool(oe["nulls_first"])
        if rn:                              # :418-421
            return not oe["nulls_first"]
        if oe.get("dictionary") is not None:  # :425-438: dictionary targets compare by string
            l, r = oe["dictionary"][int(l)], oe["dictionary"][int(r)]
        if l == r:                          # :440-442
            continue
        return (l < r) != bool(oe["is_desc"])   # :443-456
    return False


def sort_permutation(cols, n_rows, order, top_n=0):
    """sortResultSet's generic path (QE/ResultSetSort.cpp:836-849): the permutation of all rows sorted with the
    comparator, cut to top_n (topPermutation, :504-520).  Pure Python: small cases only."""
    import functools
    cmp = lambda a, b: -1 if result_set_less(order, cols, a, b) else (1 if result_set_less(order, cols, b, a) else 0)  # noqa: E731
    perm = sorted(range(n_rows), key=functools.cmp_to_key(cmp))
    return perm[:top_n] if top_n else perm
