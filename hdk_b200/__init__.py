"""hdk_b200 — B200-native implementation of intel/hdk's fused scan/filter/group-by/aggregate
and hash-join hot path, behind HDK's Executor / RelAlgExecutor / ResultSet API shape.

Planning (ir, sql, planner, storage) is host-side Python; all data-path work happens in
hdk_b200/csrc/libhdk_b200.so (hand-written sm_100a CUDA behind the C ABI in include/hdk_b200.h).
There is no CPU fallback: executing a query without the CUDA library or without a GPU raises.
"""
from . import abi, ir, planner, sql, storage  # noqa: F401

__version__ = "0.1.0"
