"""INTEGRATION.md's C++ shim (B200Kernel : DeviceKernel) must at least compile against include/hdk_b200.h: the code
block is extracted from the document and built with stand-ins for the handful of HDK declarations it touches
(DeviceKernel / CompilationContext / KernelOptions as in QE/DeviceKernel.h:25-61, QE/CompilationContext.h:23-26)."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STUBS = r'''
#include <cstdint>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <vector>
// stand-ins for the reference's declarations (QE/DeviceKernel.h:25-61, QE/CompilationContext.h:23-26)
struct KernelOptions { unsigned gridDimX = 1, gridDimY = 1, gridDimZ = 1, blockDimX = 1, blockDimY = 1, blockDimZ = 1;
                       unsigned sharedMemBytes = 0; unsigned literalsOffset = 0; bool hoistLiterals = true; };
class CompilationContext { public: virtual ~CompilationContext() {} };
class DeviceClock { public: virtual void start() = 0; virtual int stop() = 0; virtual ~DeviceClock() = default; };   // QE/DeviceKernel.h:25-31
class CudaEventClock : public DeviceClock {            // QE/DeviceKernel.cpp:25-42 (cuEvent* calls elided)
 public:
  void start() override {}
  int stop() override { return 0; }
};
class DeviceKernel {                                   // QE/DeviceKernel.h:45-61
 public:
  virtual void launch(const KernelOptions&, std::vector<int8_t*>& kernelParams) = 0;
  virtual void initializeDynamicWatchdog(bool, uint64_t, size_t) {}
  virtual void initializeRuntimeInterrupter() {}
  virtual std::unique_ptr<DeviceClock> make_clock() = 0;
  virtual ~DeviceKernel() = default;
};
class DeviceAllocator { public: int8_t* alloc(size_t) { return nullptr; } };
'''


def test_b200kernel_shim_compiles():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.findall(r"```cpp\n(.*?)```", doc, flags=re.S)[0]
    assert "class B200Kernel : public DeviceKernel" in block
    block = block.replace('#include "QueryEngine/DeviceKernel.h"', "")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "shim.cpp")
        with open(src, "w") as f:
            f.write(STUBS + block + "\nint main() { B200CompilationContext c; DeviceAllocator a; B200Kernel k(&c, &a); auto clk = k.make_clock(); clk->start(); return clk->stop(); }\n")
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), src],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]


def test_header_compiles_as_plain_c_and_declares_the_round2_entry_points():
    """The drop-in boundary is a C ABI: include/hdk_b200.h must compile as C99 (no C++ in the signatures), and the entry
    points INTEGRATION.md §1b / §3b tell a maintainer to call must be declared with the argument lists the document shows."""
    src_text = r'''
#include "hdk_b200.h"
static int use(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const int64_t* cells, int8_t* values, uint32_t* validity,
               uint64_t* null_count, void* stream) {
  size_t scratch = 0;
  hdk_b200_jit_stats st;
  int rc = hdk_b200_launch_scratch_bytes(plan, qmd, (uint64_t)1000000, &scratch);
  rc |= hdk_b200_arrow_column_on_device(cells, (uint64_t)10, 0, 0, 8, 1, INT64_MIN, 0.0, values, validity, null_count, stream);
  rc |= hdk_b200_debug_set("jit", 2);
  rc |= hdk_b200_jit_get_stats(&st);
  rc |= hdk_b200_jit_wait();
  rc |= hdk_b200_jit_shutdown();
  return rc + (HDK_B200_ERR_CLAIM_TIMEOUT == 1005) + (HDK_B200_ERR_PEER_TIMEOUT == 1004);
}
int main(void) { return use(0, 0, 0, 0, 0, 0, 0) ? 0 : 1; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "abi.c")
        with open(src, "w") as f:
            f.write(src_text)
        r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
