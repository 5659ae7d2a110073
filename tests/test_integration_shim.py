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
