"""The C-ABI library loads, exports every symbol include/hdk_b200.h declares, and the ctypes mirror
(hdk_b200/abi.py) has the same struct sizes as the C header.  No compute calls: runs without a GPU."""
import ctypes as C
import os
import re
import subprocess
import tempfile

from hdk_b200 import _lib, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hdk_b200.h")


def declared_functions():
    src = open(HEADER).read()
    return re.findall(r"^HDK_B200_API\s+[\w\s\*]+?\b(hdk_b200_\w+)\s*\(", src, flags=re.M)


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hdk_b200.h but not exported"
    assert sorted(names) == sorted(_lib.EXPORTS)


def test_abi_version_and_loader():
    L = _lib.lib()
    assert L.hdk_b200_abi_version() == abi.ABI_VERSION
    assert L.hdk_b200_device_count() >= 0
    assert isinstance(L.hdk_b200_last_error(), bytes)


def test_struct_sizes_match_header():
    structs = {"hdk_b200_type": abi.Type, "hdk_b200_expr": abi.Expr, "hdk_b200_target": abi.Target, "hdk_b200_qmd": abi.Qmd,
               "hdk_b200_key": abi.Key, "hdk_b200_join": abi.Join, "hdk_b200_plan": abi.Plan,
               "hdk_b200_kernel_params": abi.KernelParams, "hdk_b200_kernel_options": abi.KernelOptions,
               "hdk_b200_launch_info": abi.LaunchInfo, "hdk_b200_work_table_layout": abi.WorkTableLayout,
               "hdk_b200_join_chunk": abi.JoinChunk, "hdk_b200_join_column": abi.JoinColumn,
               "hdk_b200_join_column_type_info": abi.JoinColumnTypeInfo, "hdk_b200_chunk_stats": abi.ChunkStatsPOD}
    prog = '#include <stdio.h>\n#include "hdk_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert C.sizeof(structs[n]) == int(sz), f"{n}: ctypes {C.sizeof(structs[n])} != C {sz}"


def test_plan_check_rejects_bad_plans_without_gpu():
    L = _lib.lib()
    p, q = abi.Plan(), abi.Qmd()
    assert L.hdk_b200_plan_check(C.byref(p), C.byref(q), None) == abi.E_INVALID      # wrong ABI version / empty
    assert b"ABI" in L.hdk_b200_last_error()
