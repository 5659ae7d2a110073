// hdk_b200/csrc/jit.cu — run-time specialisation of the fused kernel for plan shapes without pre-compiled kernels.
//
// The reference compiles every work unit with LLVM (Executor::compileWorkUnit, QE/NativeCodegen.cpp:1403-1560) and keeps
// the native code in a cache keyed by the plan (QE/CodeCacheAccessor).  Here the kernel is a C++ template over the plan's
// STRUCTURE (scan_kernel<strategy, Shape>, scan_kernel.cuh); this file instantiates it with NVRTC for a shape seen at run
// time: the shape's constexpr DPlan initialiser is printed by the same code that generates static_shapes.inc at build
// time (dump_shape_text, lower.cu), the headers are embedded in the library (jit_sources.inc), libnvrtc is loaded on
// demand.  The cubin is loaded through the runtime's library API (cudaLibraryLoadData) and its kernels are launched like
// the pre-compiled ones.  One worker thread compiles; launches of a shape that is not ready run the interpreting kernel.
#include <dlfcn.h>
#include <nvrtc.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "partagg.cuh"
#include "scan.cuh"

namespace hb {

struct JitSource {
  const char* name;
  const char* text;
};
static const JitSource kJitSources[] = {
#include "jit_sources.inc"
};

JitStats g_jit_stats{};

namespace {

struct Nvrtc {
  void* handle = nullptr;
  bool ok = false;
  std::string include_dir;
  decltype(&nvrtcCreateProgram) create = nullptr;
  decltype(&nvrtcDestroyProgram) destroy = nullptr;
  decltype(&nvrtcCompileProgram) compile = nullptr;
  decltype(&nvrtcAddNameExpression) add_name = nullptr;
  decltype(&nvrtcGetLoweredName) lowered = nullptr;
  decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
  decltype(&nvrtcGetCUBIN) cubin = nullptr;
  decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
  decltype(&nvrtcGetProgramLog) log = nullptr;
};

static bool file_exists(const std::string& p) {
  FILE* f = fopen(p.c_str(), "r");
  if (f) fclose(f);
  return f != nullptr;
}

static Nvrtc load_nvrtc() {
  Nvrtc n;
  for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) {
    n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (n.handle) break;
  }
  if (!n.handle) return n;
#define HB_SYM(field, sym)                                               \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.handle, #sym)); \
  if (!n.field) return n;
  HB_SYM(create, nvrtcCreateProgram)
  HB_SYM(destroy, nvrtcDestroyProgram)
  HB_SYM(compile, nvrtcCompileProgram)
  HB_SYM(add_name, nvrtcAddNameExpression)
  HB_SYM(lowered, nvrtcGetLoweredName)
  HB_SYM(cubin_size, nvrtcGetCUBINSize)
  HB_SYM(cubin, nvrtcGetCUBIN)
  HB_SYM(log_size, nvrtcGetProgramLogSize)
  HB_SYM(log, nvrtcGetProgramLog)
#undef HB_SYM
  // libcu++ (<cuda/std/...>: fixed-width types and type traits without host headers) ships with the toolkit's headers
  std::vector<std::string> dirs;
  for (const char* env : {"HDK_B200_CUDA_INCLUDE", "CUDA_HOME", "CUDA_PATH"})
    if (const char* v = getenv(env)) dirs.push_back(std::string(v) + (strcmp(env, "HDK_B200_CUDA_INCLUDE") ? "/include" : ""));
  Dl_info info;
  if (dladdr(reinterpret_cast<void*>(n.create), &info) && info.dli_fname) {
    std::string lib(info.dli_fname);
    const size_t slash = lib.rfind('/');
    if (slash != std::string::npos) {
      dirs.push_back(lib.substr(0, slash) + "/../include");
      dirs.push_back(lib.substr(0, slash) + "/../../include");
    }
  }
  dirs.push_back("/usr/local/cuda/include");
  for (const std::string& d : dirs)
    if (file_exists(d + "/cuda/std/cstdint")) { n.include_dir = d; break; }
  n.ok = !n.include_dir.empty();
  return n;
}

static Nvrtc& nvrtc() {
  static Nvrtc n = load_nvrtc();
  return n;
}

// kCompiled: the cubin exists but is not loaded yet.  The worker thread only runs NVRTC (host code); everything that touches
// the CUDA runtime — cudaLibraryLoadData, cudaLibraryGetKernel — happens on the next launching thread that asks for the
// shape (load_entry), so no CUDA call of ours can race the runtime's teardown at process exit.
enum { kCompiling = 0, kReady = 1, kFailed = 2, kCompiled = 3 };
struct JitWant { int strategy, g, slot; bool reg; };
struct JitEntry {
  int state = kCompiling;
  StaticEntry e{};
  cudaLibrary_t lib = nullptr;
  DPlan plan;
  uint64_t sig = 0;
  int dev = 0;
  int iter_rows = 1;
  std::string name;
  std::vector<char> cubin;
  std::vector<JitWant> wants;
  std::vector<std::string> lowered;   // mangled kernel names, parallel to wants
};

// (never destroyed: the detached worker may still wait on them when the process exits)
std::mutex& g_mutex = *new std::mutex;
std::condition_variable& g_cv = *new std::condition_variable;
std::map<uint64_t, JitEntry*>& g_table = *new std::map<uint64_t, JitEntry*>;
std::deque<JitEntry*>& g_queue = *new std::deque<JitEntry*>;
bool g_worker_started = false;
constexpr size_t kMaxShapes = 4096;   // loaded cubins kept for the life of the process (~100 KB of device code each)
constexpr size_t kMaxQueue = 16;   // shapes waiting beyond this are simply not specialised (they keep running interpreted)

static bool wants_registers(const DPlan& p) {   // = shape_wants_registers (scan.cu)
  int wide = 0;
  for (int a = 0; a < p.n_acc; ++a) wide += p.accs[a].bytes == 8;
  return p.hash_type == HDK_B200_PERFECT_HASH && p.n_joins == 0 && wide >= 3 && wide * 2 + (p.n_acc - wide) <= 18;
}

static void compile_entry(JitEntry* je) {
  Nvrtc& n = nvrtc();
  const auto t0 = std::chrono::steady_clock::now();
  const DPlan& p = je->plan;
  std::vector<char> shape(1 << 16);
  bool ok = dump_shape_text(p, shape.data(), shape.size()) >= 0;
  const int rpi = p.n_exprs <= 6 ? 4 : p.n_exprs <= 12 ? 2 : 1;   // (as tools/gen_static_shapes.py)
  std::string src = "#define HB_JIT 1\n#include \"scan_kernel.cuh\"\nnamespace hb {\ntemplate <> struct StaticShape<1000> {\n"
                    "  static constexpr bool is_static = true;\n  static constexpr int rows_per_iter = " + std::to_string(rpi) + ";\n"
                    "  __host__ __device__ static constexpr DPlan get() { return DPlan " + std::string(shape.data()) + "; }\n};\n}\n";
  std::vector<JitWant> wants;
  if (p.hash_type == HDK_B200_BASELINE_HASH) {
    wants.push_back({HDK_B200_STRATEGY_BASELINE, 8, HDK_B200_STRATEGY_BASELINE, false});
  } else {
    for (int s : {HDK_B200_STRATEGY_THREAD_PRIVATE, HDK_B200_STRATEGY_CTA_SHARED, HDK_B200_STRATEGY_GLOBAL}) wants.push_back({s, 8, s, false});
    if (wants_registers(p))
      for (int gi = 0; gi < 4; ++gi) wants.push_back({HDK_B200_STRATEGY_REGISTER, 2 * (gi + 1), gi, true});
  }
  nvrtcProgram prog = nullptr;
  std::vector<const char*> texts, names;
  for (const JitSource& s : kJitSources) { texts.push_back(s.text); names.push_back(s.name); }
  std::string log;
  std::vector<std::string> exprs;
  if (ok) ok = n.create(&prog, src.c_str(), "hdk_b200_jit.cu", int(texts.size()), texts.data(), names.data()) == NVRTC_SUCCESS;
  if (ok) {
    for (const JitWant& w : wants) {
      exprs.push_back("hb::scan_kernel<" + std::to_string(w.strategy) + ", hb::StaticShape<1000>, " + std::to_string(w.g) + ">");
      ok = ok && n.add_name(prog, exprs.back().c_str()) == NVRTC_SUCCESS;
    }
  }
  if (ok) {
    const std::string inc = "-I" + n.include_dir;
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo", "-diag-suppress=549", inc.c_str()};
    const nvrtcResult rc = n.compile(prog, int(sizeof(opts) / sizeof(opts[0])), opts);
    if (rc != NVRTC_SUCCESS) {
      size_t ls = 0;
      if (n.log_size(prog, &ls) == NVRTC_SUCCESS && ls > 1) { log.resize(ls); n.log(prog, &log[0]); }
      ok = false;
    }
  }
  std::vector<char> cubin;
  if (ok) {
    size_t cs = 0;
    ok = n.cubin_size(prog, &cs) == NVRTC_SUCCESS && cs > 0;
    if (ok) { cubin.resize(cs); ok = n.cubin(prog, cubin.data()) == NVRTC_SUCCESS; }
  }
  std::vector<std::string> lowered;
  for (size_t i = 0; ok && i < wants.size(); ++i) {
    const char* name = nullptr;
    ok = n.lowered(prog, exprs[i].c_str(), &name) == NVRTC_SUCCESS && name;
    if (ok) lowered.push_back(name);
  }
  if (prog) n.destroy(&prog);
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (ok) {
      int mw = 1;
      for (int c = 0; c < p.n_cols; ++c) mw = p.col_width[c] > mw ? p.col_width[c] : mw;
      const int vw = 16 / mw;                                   // = shape_iter_rows<Shape>() (scan_kernel.cuh)
      je->iter_rows = rpi > vw ? rpi / vw * vw : vw;
      je->cubin.swap(cubin);
      je->wants = wants;
      je->lowered.swap(lowered);
      je->state = kCompiled;
    } else {
      je->state = kFailed;
      ++g_jit_stats.failed;
      if (getenv("HDK_B200_JIT_VERBOSE")) fprintf(stderr, "[hdk_b200 jit] shape %016llx failed to compile:\n%s\n", (unsigned long long)je->sig, log.c_str());
    }
    g_jit_stats.last_compile_ms = ms;
    g_jit_stats.total_compile_ms += ms;
    --g_jit_stats.pending;   // (pending = queued or compiling; a compiled shape is loaded by the next launch that asks for it)
  }
  g_cv.notify_all();
}

// second half, on a launching thread (g_mutex held, the thread's current device is je->dev): load the cubin, look the kernels up
static void load_entry(JitEntry* je) {
  StaticEntry e{};
  cudaLibrary_t lib = nullptr;
  bool ok = cudaLibraryLoadData(&lib, je->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess;
  for (size_t i = 0; ok && i < je->wants.size(); ++i) {
    cudaKernel_t k = nullptr;
    ok = cudaLibraryGetKernel(&k, lib, je->lowered[i].c_str()) == cudaSuccess;
    if (!ok) break;
    ScanKernelFn fn = reinterpret_cast<ScanKernelFn>(k);
    const JitWant& w = je->wants[i];
    if (w.reg) {
      e.reg_fn[w.slot] = fn;
      if (w.g == 8) e.fn[HDK_B200_STRATEGY_REGISTER] = fn;
    } else {
      e.fn[w.slot] = fn;
    }
  }
  if (!ok) cudaGetLastError();
  std::vector<char>().swap(je->cubin);
  if (ok) {
    e.iter_rows = je->iter_rows;
    e.sig = je->sig;
    je->name = "jit_" + std::to_string(je->sig);
    e.name = je->name.c_str();
    je->e = e;
    je->lib = lib;
    je->state = kReady;
    ++g_jit_stats.compiled;
  } else {
    je->state = kFailed;
    ++g_jit_stats.failed;
  }
}

// Process exit while a compile is in flight: the detached worker would still be inside NVRTC when static state it uses is
// torn down (observed: SIGSEGV at interpreter exit of short scripts; NVRTC loads libnvrtc-builtins lazily, so its
// destructors can be registered AFTER any handler of ours and run before it).  hdk_b200_jit_shutdown() drops the queue and
// waits for the compile in flight (<= a few seconds): a host calls it before it exits — the Python binding registers it with
// `atexit`, which runs ahead of every C-level exit handler; std::atexit keeps a best-effort copy for other hosts.  The worker
// never calls the CUDA runtime (see kCompiled), so it cannot race the runtime's own teardown either.
bool g_shutdown = false;
int g_busy = 0;
static void jit_shutdown() {
  std::unique_lock<std::mutex> lock(g_mutex);
  g_shutdown = true;
  g_jit_stats.pending -= g_queue.size();
  g_queue.clear();
  g_cv.notify_all();
  g_cv.wait(lock, [] { return g_busy == 0; });
}

static void worker() {
  for (;;) {
    JitEntry* je = nullptr;
    {
      std::unique_lock<std::mutex> lock(g_mutex);
      g_cv.wait(lock, [] { return g_shutdown || !g_queue.empty(); });
      if (g_shutdown) return;
      je = g_queue.front();
      g_queue.pop_front();
      ++g_busy;
    }
    compile_entry(je);
    {
      std::lock_guard<std::mutex> lock(g_mutex);
      --g_busy;
    }
    g_cv.notify_all();
  }
}

}  // namespace

const StaticEntry* jit_scan_kernels(const DPlan& p, uint64_t sig, bool wait) {
  if (!nvrtc().ok) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::unique_lock<std::mutex> lock(g_mutex);
  auto it = g_table.find(sig);
  JitEntry* je = it == g_table.end() ? nullptr : it->second;
  if (je && !same_plan_shape(je->plan, p)) return nullptr;   // signature collision: this shape stays on the interpreter
  if (!je) {
    if (!wait && (g_shutdown || g_queue.size() >= kMaxQueue)) return nullptr;
    if (g_table.size() >= kMaxShapes) return nullptr;         // the cache is full: further shapes stay on the interpreter
    je = new JitEntry();
    je->plan = p;
    je->sig = sig;
    je->dev = dev;
    g_table[sig] = je;
    ++g_jit_stats.pending;
    if (wait) {
      lock.unlock();
      compile_entry(je);
      lock.lock();
    } else {
      g_queue.push_back(je);
      if (!g_worker_started) {
        g_worker_started = true;
        std::atexit(jit_shutdown);
        std::thread(worker).detach();
      }
      g_cv.notify_all();
      return nullptr;
    }
  }
  if (je->state == kCompiling && wait) g_cv.wait(lock, [je] { return je->state != kCompiling; });
  if (je->state == kCompiled && je->dev == dev) load_entry(je);
  if (je->state != kReady) return nullptr;
  ++g_jit_stats.launches;
  return &je->e;
}

}  // namespace hb

extern "C" {

int hdk_b200_jit_get_stats(hdk_b200_jit_stats* out) {
  if (!out) return HDK_B200_E_INVALID;
  const bool ok = hb::nvrtc().ok;
  std::lock_guard<std::mutex> lock(hb::g_mutex);
  out->shapes_compiled = hb::g_jit_stats.compiled;
  out->shapes_failed = hb::g_jit_stats.failed;
  out->shapes_pending = hb::g_jit_stats.pending;
  out->launches = hb::g_jit_stats.launches;
  out->last_compile_ms = hb::g_jit_stats.last_compile_ms;
  out->total_compile_ms = hb::g_jit_stats.total_compile_ms;
  out->available = ok ? 1 : 0;
  return HDK_B200_OK;
}

int hdk_b200_jit_shutdown(void) {
  hb::jit_shutdown();
  return HDK_B200_OK;
}

int hdk_b200_jit_wait(void) {
  std::unique_lock<std::mutex> lock(hb::g_mutex);
  hb::g_cv.wait(lock, [] { return hb::g_jit_stats.pending == 0; });
  return HDK_B200_OK;
}

}  // extern "C"
