"""Times hdk_b200_sort_permutation on a synthetic compacted result (the output size of BASELINE config 4: 1e8 groups).
CUDA events on the launching stream, warm-up first; prints one JSON line per case.  Under ncu
(`ncu --metrics gpu__time_duration.sum -k "regex:sort_|select_|gather_"`) `--once` runs each case a single time."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from hdk_b200 import _lib, abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    n = a.rows
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    cases = {
        "sum_i64_41bit": torch.randint(-(1 << 40), 1 << 40, (n,), dtype=torch.int64, device=dev, generator=g),
        "count_20bit": torch.randint(0, 1 << 20, (n,), dtype=torch.int64, device=dev, generator=g),
        "avg_f64": torch.rand(n, dtype=torch.float64, device=dev, generator=g).mul_(1000).view(torch.int64),
    }
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    sb = L.hdk_b200_sort_scratch_bytes(n)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    for name, col in cases.items():
        oe = (abi.OrderEntry * 1)()
        oe[0].column, oe[0].is_fp, oe[0].type_width, oe[0].nullable, oe[0].is_desc = 0, int(name.endswith("f64")), 8, 1, 1
        ptrs = (C.c_void_p * abi.MAX_TARGETS)(col.data_ptr())
        for top_n in (0, 100):
            def run():
                _lib.check(L.hdk_b200_sort_permutation(ptrs, oe, 1, n, top_n, perm.data_ptr(), None, scratch.data_ptr(), sb, None), "sort")
            reps = 1 if a.once else 5
            if not a.once:
                run()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(reps):
                run()
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / reps
            print(json.dumps({"case": name, "rows": n, "top_n": top_n, "ms": round(ms, 3), "rows_per_s": round(n / ms * 1e3)}))


if __name__ == "__main__":
    main()
