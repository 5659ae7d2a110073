"""Loader and ctypes prototypes of hdk_b200/csrc/libhdk_b200.so (the sm_100a CUDA library behind
include/hdk_b200.h).  There is deliberately no fallback: if the library is missing, or a function
is called without a GPU, the call fails loudly."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HDK_B200_LIB") or os.path.join(_HERE, "csrc", "libhdk_b200.so")   # (override: A/B builds while tuning)
_lib = None


class HdkB200Error(RuntimeError):
    pass


EXPORTS = [
    "hdk_b200_plan_check", "hdk_b200_buffer_size_bytes", "hdk_b200_init_group_by_buffer",
    "hdk_b200_init_group_by_buffer_on_device", "hdk_b200_init_columnar_group_by_buffer_on_device",
    "hdk_b200_launch", "hdk_b200_work_table_layout_get", "hdk_b200_init_work_table", "hdk_b200_launch_partial",
    "hdk_b200_finalize", "hdk_b200_reduce", "hdk_b200_init_hash_join_buff_on_device",
    "hdk_b200_fill_hash_join_buff_on_device", "hdk_b200_fill_one_to_many_hash_table_on_device",
    "hdk_b200_init_baseline_hash_join_buff_on_device", "hdk_b200_fill_baseline_hash_join_buff_on_device",
    "hdk_b200_fill_one_to_many_baseline_hash_table_on_device", "hdk_b200_probe_hash_join_on_device",
    "hdk_b200_probe_baseline_hash_join_on_device", "hdk_b200_gather_join_payload_on_device", "hdk_b200_shuffle_count", "hdk_b200_shuffle_scatter", "hdk_b200_shuffle_scatter_to", "hdk_b200_region_count", "hdk_b200_region_scatter_to",
    "hdk_b200_compact_result", "hdk_b200_sort_scratch_bytes", "hdk_b200_sort_permutation", "hdk_b200_gather_rows",
    "hdk_b200_init_chunk_stats_on_device",
    "hdk_b200_materialize_nulls_on_device", "hdk_b200_peer_alloc", "hdk_b200_peer_open", "hdk_b200_peer_close",
    "hdk_b200_peer_free", "hdk_b200_exchange_bytes", "hdk_b200_exchange_init", "hdk_b200_launch_exchange", "hdk_b200_query_host", "hdk_b200_last_error", "hdk_b200_abi_version",
    "hdk_b200_device_count", "hdk_b200_launch_count", "hdk_b200_launch_scratch_bytes", "hdk_b200_debug_set",
    "hdk_b200_jit_get_stats", "hdk_b200_jit_wait", "hdk_b200_jit_shutdown", "hdk_b200_arrow_column_on_device",
]


def _bind(lib):
    P, Q = C.POINTER(abi.Plan), C.POINTER(abi.Qmd)
    KP, KO, LI = C.POINTER(abi.KernelParams), C.POINTER(abi.KernelOptions), C.POINTER(abi.LaunchInfo)
    JC, JT = C.POINTER(abi.JoinColumn), C.POINTER(abi.JoinColumnTypeInfo)
    vp, i64, u64, i32, u32, sz, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_uint32, C.c_size_t, C.c_int
    sig = {
        "hdk_b200_plan_check": (ci, [P, Q, C.POINTER(sz)]),
        "hdk_b200_buffer_size_bytes": (sz, [Q]),
        "hdk_b200_init_group_by_buffer": (ci, [Q, vp, vp]),
        "hdk_b200_init_group_by_buffer_on_device": (ci, [vp, vp, u32, u32, u32, u32, ci, C.c_int8, sz, sz, vp]),
        "hdk_b200_init_columnar_group_by_buffer_on_device": (ci, [vp, vp, u32, u32, u32, vp, ci, ci, C.c_int8, sz, sz, vp]),
        "hdk_b200_launch": (ci, [P, Q, KO, KP, vp, sz, vp, LI]),
        "hdk_b200_work_table_layout_get": (ci, [P, Q, C.POINTER(abi.WorkTableLayout)]),
        "hdk_b200_init_work_table": (ci, [P, Q, vp, vp]),
        "hdk_b200_launch_partial": (ci, [P, Q, KO, KP, vp, vp, LI]),
        "hdk_b200_finalize": (ci, [P, Q, vp, vp, vp]),
        "hdk_b200_reduce": (ci, [P, Q, vp, vp, u32, vp, vp]),
        "hdk_b200_init_hash_join_buff_on_device": (ci, [vp, i64, i32, vp]),
        "hdk_b200_fill_hash_join_buff_on_device": (ci, [vp, i32, ci, vp, JC, JT, i64, vp]),
        "hdk_b200_fill_one_to_many_hash_table_on_device": (ci, [vp, i64, i32, JC, JT, i64, vp]),
        "hdk_b200_init_baseline_hash_join_buff_on_device": (ci, [vp, i64, sz, ci, i32, ci, vp]),
        "hdk_b200_fill_baseline_hash_join_buff_on_device": (ci, [vp, i64, i32, ci, sz, ci, vp, JC, JT, ci, vp]),
        "hdk_b200_fill_one_to_many_baseline_hash_table_on_device": (ci, [vp, vp, i64, i32, sz, JC, JT, ci, vp]),
        "hdk_b200_probe_hash_join_on_device": (ci, [vp, vp, i64, i64, i64, vp, vp]),
        "hdk_b200_probe_baseline_hash_join_on_device": (ci, [vp, vp, i64, sz, ci, i64, ci, vp, vp]),
        "hdk_b200_gather_join_payload_on_device": (ci, [vp, i64, vp, ci, vp, vp, vp]),
        "hdk_b200_shuffle_count": (ci, [P, KP, u32, vp, vp]),
        "hdk_b200_shuffle_scatter": (ci, [P, KP, u32, vp, vp, vp, vp]),
        "hdk_b200_shuffle_scatter_to": (ci, [P, KP, u32, vp, vp, vp, vp]),
        "hdk_b200_region_count": (ci, [P, Q, KP, u32, vp, vp]),
        "hdk_b200_region_scatter_to": (ci, [P, Q, KP, u32, vp, vp, vp, vp]),
        "hdk_b200_compact_result": (ci, [P, Q, vp, vp, vp, vp]),
        "hdk_b200_sort_scratch_bytes": (sz, [u64]),
        "hdk_b200_sort_permutation": (ci, [C.POINTER(vp), C.POINTER(abi.OrderEntry), ci, u64, u64, vp, C.POINTER(u64), vp, sz, vp]),
        "hdk_b200_gather_rows": (ci, [C.POINTER(vp), C.POINTER(vp), ci, vp, u64, vp]),
        "hdk_b200_init_chunk_stats_on_device": (ci, [vp, vp]),
        "hdk_b200_materialize_nulls_on_device": (ci, [vp, ci, ci, vp, i64, i64, vp, vp]),
        "hdk_b200_peer_alloc": (ci, [sz, C.POINTER(vp), vp]),
        "hdk_b200_peer_open": (ci, [vp, C.POINTER(vp)]),
        "hdk_b200_peer_close": (ci, [vp]),
        "hdk_b200_peer_free": (ci, [vp]),
        "hdk_b200_exchange_bytes": (ci, [P, Q, ci, C.POINTER(sz)]),
        "hdk_b200_exchange_init": (ci, [vp, vp]),
        "hdk_b200_launch_exchange": (ci, [P, Q, KO, KP, vp, sz, C.POINTER(vp), ci, ci, u64, vp, LI]),
        "hdk_b200_query_host": (ci, [P, Q, vp, vp, u64, vp, vp, vp, vp, vp, ci, LI]),
        "hdk_b200_last_error": (C.c_char_p, []),
        "hdk_b200_abi_version": (ci, []),
        "hdk_b200_device_count": (ci, []),
        "hdk_b200_launch_count": (u64, []),
        "hdk_b200_launch_scratch_bytes": (ci, [P, Q, u64, C.POINTER(sz)]),
        "hdk_b200_debug_set": (ci, [C.c_char_p, ci]),
        "hdk_b200_jit_get_stats": (ci, [C.POINTER(abi.JitStats)]),
        "hdk_b200_jit_wait": (ci, []),
        "hdk_b200_jit_shutdown": (ci, []),
        "hdk_b200_arrow_column_on_device": (ci, [vp, u64, ci, ci, ci, ci, i64, C.c_double, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HdkB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C hdk_b200/csrc).  hdk_b200 has no CPU fallback.")
        _lib = _bind(C.CDLL(LIB_PATH))
        if _lib.hdk_b200_abi_version() != abi.ABI_VERSION:
            raise HdkB200Error("libhdk_b200.so ABI version mismatch; rebuild")
        # the run-time compiler's worker thread must not outlive the libraries it uses: stop it at interpreter exit, ahead of
        # every C-level exit handler (hdk_b200_jit_shutdown waits for the compile in flight, a few seconds at most)
        import atexit
        atexit.register(_lib.hdk_b200_jit_shutdown)
    return _lib


def debug_set(name: str, value: int):
    """hdk_b200_debug_set: process-wide debug / tuning knobs ("force_generic", "force_strategy", "partitioned_aggregation")."""
    check(lib().hdk_b200_debug_set(name.encode(), int(value)), f"debug_set({name})")


def jit_stats() -> dict:
    """hdk_b200_jit_get_stats as a dict (shapes compiled / failed / pending, launches on run-time compiled kernels, compile ms)."""
    st = abi.JitStats()
    check(lib().hdk_b200_jit_get_stats(C.byref(st)), "jit_get_stats")
    return {f: getattr(st, f) for f, _ in abi.JitStats._fields_}


def check(rc, what=""):
    if rc != 0:
        msg = lib().hdk_b200_last_error().decode(errors="replace")
        if rc == abi.E_UNSUPPORTED:     # a plan shape outside the hot path reads the same whether the planner or the library refuses it
            from .planner import UnsupportedPlan
            raise UnsupportedPlan(f"{what}: {msg}")
        raise HdkB200Error(f"{what} failed with code {rc}: {msg}")
