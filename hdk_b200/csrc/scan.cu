// hdk_b200/csrc/scan.cu — the fused scan → filter → join probe → group-by → aggregate kernel.
//
// Replaces the JIT'd multifrag_query_hoisted_literals / query_group_by_template / row_func of the
// reference (QE/RuntimeFunctions.cpp:1692-1726, QE/QueryTemplateGenerator.cpp:488-772) and the
// runtime it calls per row (QE/GroupByRuntime.cpp, QE/cuda_mapd_rt.cu:423-1083).
//
// Execution model (sm_100a):
//   * persistent grid, a multiple of the SM count; every CTA walks tiles t = blockIdx.x, +gridDim.x
//   * one producer warp streams each tile's column slices global → shared with 1-D TMA bulk copies
//     (cp.async.bulk … mbarrier::complete_tx, SASS UBLKCP) through a kStages-deep mbarrier ring;
//     unaligned heads/tails (< 16 B per column) are patched with byte copies, so any chunk pointer
//     and any row count work on the same path
//   * kConsumerWarps consumer warps evaluate the expression DAG per row out of shared memory and
//     accumulate into NEUTRAL accumulators (SUM→0, MIN→+max, MAX→−max, counters) — the same
//     representation merges across CTAs (atomics) and across GPUs (NCCL SUM/MIN/MAX):
//       THREAD_PRIVATE : per-thread bins in shared memory, [acc][group][thread], no atomics,
//                        bank-conflict free by construction          (few groups)
//       CTA_SHARED     : one table per CTA in shared memory, shared atomics   (up to ~100 KB)
//       GLOBAL         : straight into the global work table with RED atomics
//   * a finalize kernel converts the work table into the reference's buffer encoding
//     (finalize.cu), so the result is byte-compatible with QueryMemoryDescriptor.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "accum.cuh"
#include "baseline.cuh"
#include "device_utils.cuh"
#include "eval.cuh"
#include "partagg.cuh"
#include "scan.cuh"
#include "scan_kernel.cuh"
#include "shape.cuh"

namespace hb {

// perfect-hash shapes get the three accumulation strategies, baseline-hash shapes the in-place one
// REGISTER only where it can win: several wide accumulators (it replaces n_acc shared-memory updates per row by
// kRegGroups predicated register updates per accumulator) and a register budget that fits
__host__ __device__ constexpr bool shape_wants_registers(const DPlan& p) {
  int wide = 0;
  for (int a = 0; a < p.n_acc; ++a) wide += p.accs[a].bytes == 8;
  return p.hash_type == HDK_B200_PERFECT_HASH && p.n_joins == 0 && wide >= 3 && wide * 2 + (p.n_acc - wide) <= 18;
}
template <int ID, int S, int G = kRegGroups>
constexpr ScanKernelFn pick_kernel() {
  if constexpr (S == HDK_B200_STRATEGY_REGISTER) {
    if constexpr (shape_wants_registers(StaticShape<ID>::get())) return scan_kernel<S, StaticShape<ID>, G>;
    else return nullptr;
  } else if constexpr ((StaticShape<ID>::get().hash_type == HDK_B200_BASELINE_HASH) == (S == HDK_B200_STRATEGY_BASELINE)) {
    return scan_kernel<S, StaticShape<ID>>;
  } else {
    return nullptr;
  }
}
#define HB_STATIC_SHAPE(ID, SIG, NAME, RPI, ...)                                                              \
  {SIG, NAME, shape_iter_rows<StaticShape<ID>>(), {pick_kernel<ID, HDK_B200_STRATEGY_THREAD_PRIVATE>(), pick_kernel<ID, HDK_B200_STRATEGY_CTA_SHARED>(), \
               pick_kernel<ID, HDK_B200_STRATEGY_GLOBAL>(), pick_kernel<ID, HDK_B200_STRATEGY_BASELINE>(), \
               pick_kernel<ID, HDK_B200_STRATEGY_REGISTER>()},                                                    \
   {pick_kernel<ID, HDK_B200_STRATEGY_REGISTER, 2>(), pick_kernel<ID, HDK_B200_STRATEGY_REGISTER, 4>(),                 \
    pick_kernel<ID, HDK_B200_STRATEGY_REGISTER, 6>(), pick_kernel<ID, HDK_B200_STRATEGY_REGISTER, 8>()}},
static const StaticEntry kStaticShapes[] = {
#include "static_shapes.inc"
    {0, nullptr, 0, {nullptr, nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}}};
#undef HB_STATIC_SHAPE

// ---------------------------------------------------------------------------------------------
// work table initialisation
// ---------------------------------------------------------------------------------------------
// accumulator-major [n_acc][E] for perfect hash (merge classes are contiguous for the all-reduce),
// entry-major [E][n_acc] for baseline hash (one entry's cells share a sector)
__global__ void init_work_table_kernel(int64_t* w, uint64_t E, int n_acc, bool entry_major, const __grid_constant__ AccKinds kinds, const int* run_if) {
  if (run_if && *run_if == 0) return;
  const uint64_t n = uint64_t(n_acc) * E;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
    w[i] = acc_identity(kinds.kind[entry_major ? i % uint64_t(n_acc) : i / E]);
}

int init_work_table(const Lowered& lw, int64_t* work_table, cudaStream_t stream, const int* run_if) {
  AccKinds kinds{};
  for (int a = 0; a < lw.plan.n_acc; ++a) kinds.kind[a] = lw.plan.accs[a].kind;
  const uint64_t n = uint64_t(lw.plan.n_acc) * lw.plan.entry_count;
  const int block = 256;
  const int grid = int(std::min<uint64_t>((n + block - 1) / block, uint64_t(sm_count()) * 8));
  init_work_table_kernel<<<std::max(grid, 1), block, 0, stream>>>(work_table, lw.plan.entry_count, lw.plan.n_acc,
                                                                  lw.plan.hash_type == HDK_B200_BASELINE_HASH, kinds, run_if);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// host: geometry + launch
// ---------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// launch geometry per (plan shape, entry count, fragment count, input size class, device, forced choices): per-device
// cache, nothing query-specific in it (literals, pointers and key ranges travel in the arguments of every launch)
struct GeoKey {
  uint64_t sig, entries;
  uint32_t n_frag, hint_bucket, dev, baseline, force_generic, force_strategy, grid_x;
  bool operator==(const GeoKey& o) const {
    return sig == o.sig && entries == o.entries && n_frag == o.n_frag && hint_bucket == o.hint_bucket && dev == o.dev && baseline == o.baseline &&
           force_generic == o.force_generic && force_strategy == o.force_strategy && grid_x == o.grid_x;
  }
};
struct GeoEntry {
  GeoKey key;
  ScanArgs a;            // only the geometry fields are used
  ScanKernelFn kern;
  int grid, block, variant, strategy;
  size_t smem_bytes;
};
static std::mutex g_geo_mutex;
static std::vector<GeoEntry> g_geo_cache;
static const char* geo_env() {   // tuning hook (tools/sweep_geo.py), read once unless the sweep asks for every launch
  static const char* env = getenv("HDK_B200_GEO");
  return g_debug.geo_env_refresh ? getenv("HDK_B200_GEO") : env;
}

static int launch_scan_impl(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                            int64_t* work_table, bool baseline, cudaStream_t stream, hdk_b200_launch_info* info,
                            const ExchangeTargets* xchg = nullptr, const int* run_if = nullptr) {
  const DPlan& p = lw.plan;
  if (params->num_fragments > kMaxFragments) { set_error("more than %d fragments per launch", kMaxFragments); return HDK_B200_E_UNSUPPORTED; }
  ScanArgs a{};
  a.plan = p;
  a.col_buffers = params->col_buffers;
  a.num_rows = params->num_rows;
  a.num_fragments = uint32_t(params->num_fragments);
  a.join_hash_tables = params->join_hash_tables;
  a.inner_col_buffers = params->inner_col_buffers;
  a.work_table = work_table;
  a.error_codes = params->error_codes;
  a.layout = lw.layout;
  a.groupby_buf = params->groupby_buf;
  a.run_if = run_if;
  // nesting order of the join probes: by key node (a chained join's key comes after the inner columns it reads, hence after
  // the key of the join that supplies them); stable for joins sharing a key node
  for (int j = 0; j < p.n_joins; ++j) a.join_order[j] = uint8_t(j);
  std::stable_sort(a.join_order, a.join_order + p.n_joins, [&](uint8_t x, uint8_t y) { return p.joins[x].key_expr < p.joins[y].key_expr; });
  if (xchg) {
    a.n_peers = xchg->n_peers;
    a.n_cells = uint64_t(p.n_acc) * p.entry_count;
    a.epoch = xchg->epoch;
    a.ticket = xchg->ticket;
    for (uint32_t i = 0; i < xchg->n_peers; ++i) { a.peer_slot[i] = xchg->peer_slot[i]; a.peer_flag[i] = xchg->peer_flag[i]; }
  }

  int dev = 0;
  HB_CUDA(cudaGetDevice(&dev));
  // ---- launch geometry of an earlier launch of the same shape?  (the search below queries occupancy and sets function
  //      attributes: not something to repeat per launch of a 100 µs query)
  const uint64_t total_hint = params->total_rows_hint;
  int hint_bucket = 0;
  while ((total_hint >> hint_bucket) > 1) ++hint_bucket;
  // kernels of this plan's shape: pre-compiled (static_shapes.inc), compiled at run time by an earlier launch (jit.cu), or
  // none yet — the interpreting kernel runs, and with the JIT on the specialised ones are being built meanwhile
  const uint64_t sig = plan_signature(p);
  const StaticEntry* stat = nullptr;
  // (a one-to-many join loops over its matching set: only the interpreting kernel nests those loops, and no pre-compiled
  //  shape has one)
  bool one_to_many = false;
  for (int j = 0; j < p.n_joins; ++j) one_to_many = one_to_many || p.joins[j].one_to_many;
  if (!g_debug.force_generic && !one_to_many) {  // (hdk_b200_debug_set("force_generic", 1): tests run the interpreter on the benchmark shapes)
    for (int i = 0; kStaticShapes[i].name; ++i)
      if (kStaticShapes[i].sig == sig) { stat = &kStaticShapes[i]; break; }
    if (!stat && g_debug.jit) stat = jit_scan_kernels(p, sig, g_debug.jit == 2);
  }
  GeoKey key{sig ^ (stat ? 0x9e3779b97f4a7c15ull : 0), uint64_t(p.entry_count), a.num_fragments, uint32_t(hint_bucket), uint32_t(dev), baseline ? 1u : 0u,
             uint32_t(g_debug.force_generic), uint32_t(g_debug.force_strategy + 1), ko ? ko->gridDimX : 0u};
  {
    std::lock_guard<std::mutex> lock(g_geo_mutex);
    for (const GeoEntry& ge : g_geo_cache)
      if (ge.key == key && !geo_env()) {
        a.consumer_threads = ge.a.consumer_threads; a.n_stages = ge.a.n_stages; a.tile_rows = ge.a.tile_rows; a.off_bins = ge.a.off_bins;
        a.off_stages = ge.a.off_stages; a.off_tile_prefix = ge.a.off_tile_prefix; a.stage_bytes = ge.a.stage_bytes; a.full_iters = ge.a.full_iters;
        memcpy(a.acc_bin_off, ge.a.acc_bin_off, sizeof(a.acc_bin_off));
        memcpy(a.col_region_off, ge.a.col_region_off, sizeof(a.col_region_off));
        {
          void* kargs[] = {&a};
          HB_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(ge.kern), dim3(ge.grid), dim3(ge.block), kargs, ge.smem_bytes, stream));
        }
        HB_LAUNCH_CHECK();
        if (info) {
          info->variant = ge.variant; info->strategy = ge.strategy; info->grid = ge.grid; info->block = ge.block;
          info->smem_bytes = int(ge.smem_bytes); info->n_accumulators = p.n_acc; info->tile_rows = int(ge.a.tile_rows);
        }
        return HDK_B200_OK;
      }
  }
  int max_smem = 0;
  HB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

  // ---- shared-memory map and launch geometry ---------------------------------------------------
  // fixed part: barriers (128 B) + stage headers + tile prefix
  size_t off = 128 + align_up(sizeof(StageHeader) * kStages, 16);
  a.off_tile_prefix = uint32_t(off);
  off += align_up(size_t(a.num_fragments + 1) * 4, 16);
  const size_t fixed_bytes = align_up(off, 16);
  const size_t E = p.entry_count;
  const size_t row_bytes = std::max<size_t>(lw.stage_row_bytes, 1);
  bool counters_only = true;
  for (int i = 0; i < p.n_acc; ++i) counters_only = counters_only && p.accs[i].bytes == 4;

  // pre-compiled shape for this plan?  (its iteration granularity shapes the tile size)
  const int iter_rows = stat ? stat->iter_rows : 1;

  struct Geo {
    int strategy, nct, ctas, stages;
    uint32_t tile_rows;
    size_t off_stages;
  };
  int sm_smem = 0;
  HB_CUDA(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  // Largest tile for (strategy, consumer threads, CTAs per SM, ring depth); false if nothing fits.
  auto fit = [&](int strategy, int nct, int ctas, int stages, Geo* g) -> bool {
    if (ctas * (nct + 32) > 2048) return false;
    size_t bins = 0;
    for (int i = 0; i < p.n_acc; ++i) {
      bins = align_up(bins, 16);
      if (strategy == HDK_B200_STRATEGY_THREAD_PRIVATE) bins += size_t(p.accs[i].bytes) * E * nct;
      else if (strategy == HDK_B200_STRATEGY_CTA_SHARED) bins += size_t(p.accs[i].bytes) * E;
    }
    if (strategy == HDK_B200_STRATEGY_REGISTER) bins = size_t(kRegGroups) * kMaxAcc * 4;   // NULL-row counters of COUNT(arg)
    const size_t off_stages = align_up(fixed_bytes + bins, 128);
    // each resident CTA costs its dynamic shared memory + 1 KB reserved by the driver
    const size_t budget = std::min<size_t>(size_t(sm_smem) / ctas - 1024, size_t(max_smem) - 1024);
    if (off_stages >= budget) return false;
    const size_t per_stage = std::min<size_t>((budget - off_stages) / stages, 72 * 1024);
    // a whole number of full-tile iterations (nct * iter_rows rows each); every column slice is then a multiple of 16 bytes
    const uint32_t quantum = uint32_t(nct) * uint32_t(iter_rows);
    auto fits = [&](uint32_t rows) { return align_up(size_t(rows) * row_bytes + 48 * size_t(p.n_cols), 128) <= per_stage; };
    // small inputs: keep >= 8 tiles per resident CTA when the caller told us the row count
    uint32_t cap = 32768;
    if (params->total_rows_hint) {
      const uint64_t want = params->total_rows_hint / (uint64_t(sm_count()) * ctas * 8);
      cap = uint32_t(std::min<uint64_t>(cap, std::max<uint64_t>(want, quantum)));
    }
    uint32_t tr = 0;
    for (uint32_t cand = quantum; cand <= cap && fits(cand); cand += quantum) tr = cand;
    if (tr == 0)   // tiles smaller than one iteration quantum: scalar path only
      for (uint32_t cand = 32; cand < quantum && fits(cand); cand *= 2) tr = cand;
    if (tr == 0) return false;
    *g = Geo{strategy, nct, ctas, stages, tr, off_stages};
    return true;
  };
  // Score fitted to tools/sweep_geo.py runs on the taxi shapes (profiles/r1_geometry_sweep.md): ~768 consumer
  // threads per SM is the sweet spot (three 256-thread CTAs, each with its own ring, beat two 512-thread CTAs by
  // 5-15 %), more rows per thread per tile amortise the per-tile barrier work (~2 rows' worth), and a 2-deep ring
  // only pays off when its tiles are large enough to cover the refill latency (~0.7 rows per thread).
  auto best = [&](int strategy, Geo* out) -> bool {
    static const int ncts[] = {256, 384, 192, 512, 128, 64, 32};
    double best_score = -1.0;
    for (int nct : ncts)
      for (int ctas = 1; ctas <= 6; ++ctas)
        for (int stages = 2; stages <= 3; ++stages) {
          Geo g;
          // REGISTER kernels are compiled for <= 256 consumer threads and a large register file share
          if (strategy == HDK_B200_STRATEGY_REGISTER && (nct + 32 > kRegThreads || ctas > 2)) continue;
          if (!fit(strategy, nct, ctas, stages, &g)) continue;
          if (strategy == HDK_B200_STRATEGY_REGISTER) {   // the register file decides how many of these CTAs are resident
            ScanKernelFn k = stat->reg_fn[E <= 2 ? 0 : E <= 4 ? 1 : E <= 6 ? 2 : 3];
            const size_t smem_need = g.off_stages + size_t(stages) * align_up(size_t(g.tile_rows) * row_bytes + 48 * size_t(p.n_cols), 128);
            int occ = 0;
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 1024) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, nct + 32, smem_need) != cudaSuccess || occ < ctas) {
              cudaGetLastError();
              continue;
            }
          }
          const double T = double(nct) * ctas;
          // plans that gather through a join table wait on L2 / DRAM latency per row: they want more resident warps
          // (config 5: 4.8 ms at 768 threads per SM, 2.7 ms at 1152)
          const double t_best = p.n_joins ? 1152.0 : 768.0;
          const double teff = T <= t_best ? T : t_best - 0.5 * (T - t_best);
          const double rpt = double(g.tile_rows) / nct;
          const bool full = g.tile_rows % (uint32_t(nct) * uint32_t(iter_rows)) == 0;
          // (gathering plans also want the L1 the ring would take: config 5 with 3 x 384 threads, 2.7 ms with a 2-deep
          //  ring, 4.2 ms with a 3-deep one)
          const double ring = p.n_joins ? (stages == 2 ? 1.0 : 0.7) : (stages == 2 ? rpt / (rpt + 0.7) : 1.0);
          // (REGISTER kernels: one big CTA per SM measured 10 % faster than two small ones on TPC-H Q1 — fewer flushes)
          const double one_cta = (strategy == HDK_B200_STRATEGY_REGISTER && ctas == 1) ? 1.05 : 1.0;
          const double score = teff * (rpt / (rpt + 2.0)) * (full ? 1.0 : 0.5) * ring * one_cta;
          if (score > best_score) { best_score = score; *out = g; }
        }
    return best_score >= 0.0;
  };
  Geo geo{};
  bool have = false;
  int forced = -1;
  if (!baseline && g_debug.force_strategy >= 0 && g_debug.force_strategy <= HDK_B200_STRATEGY_REGISTER &&
      g_debug.force_strategy != HDK_B200_STRATEGY_BASELINE)
    forced = g_debug.force_strategy;   // hdk_b200_debug_set("force_strategy", s): tests / tuning
  if (baseline) have = best(HDK_B200_STRATEGY_BASELINE, &geo);
  else if (forced == HDK_B200_STRATEGY_REGISTER && !(stat && stat->fn[forced] && E <= size_t(kRegGroups))) have = false;
  else if (forced >= 0) have = best(forced, &geo);
  else {
    Geo g;
    // few groups, several wide aggregates, pre-compiled shape: accumulators live in registers
    if (stat && stat->fn[HDK_B200_STRATEGY_REGISTER] && E <= size_t(kRegGroups) && best(HDK_B200_STRATEGY_REGISTER, &g)) { geo = g; have = true; }
    // counters only and enough groups to spread the native shared atomics: one table per CTA
    if (!have && counters_only && E >= 32 && best(HDK_B200_STRATEGY_CTA_SHARED, &g) && g.ctas * g.nct >= 512) { geo = g; have = true; }
    // private bins: no atomics at all.  With few groups a shared table would serialise on its hot bins, so take
    // private bins even when only a few warps fit.
    if (!have && best(HDK_B200_STRATEGY_THREAD_PRIVATE, &g) && (g.ctas * g.nct >= 256 || E < 64)) { geo = g; have = true; }
    if (!have && best(HDK_B200_STRATEGY_CTA_SHARED, &g)) { geo = g; have = true; }
    if (!have) have = best(HDK_B200_STRATEGY_GLOBAL, &geo);
  }
  // tuning hook (tools/sweep_geo.py): HDK_B200_GEO="strategy,consumer_threads,ctas_per_sm,stages,tile_rows" overrides the search
  if (const char* env = geo_env()) {
    int es = 0, en = 0, ec = 0, est = 0, et = 0;
    if (sscanf(env, "%d,%d,%d,%d,%d", &es, &en, &ec, &est, &et) == 5 && en >= 32 && en <= kConsumerWarps * 32 && en % 32 == 0 &&
        est >= 1 && est <= kStages && et >= 32 && es >= 0 && es <= HDK_B200_STRATEGY_REGISTER &&
        (es == HDK_B200_STRATEGY_BASELINE) == baseline) {
      size_t bins = 0;
      for (int i = 0; i < p.n_acc; ++i) {
        bins = align_up(bins, 16);
        if (es == HDK_B200_STRATEGY_THREAD_PRIVATE) bins += size_t(p.accs[i].bytes) * E * en;
        else if (es == HDK_B200_STRATEGY_CTA_SHARED) bins += size_t(p.accs[i].bytes) * E;
      }
      if (es == HDK_B200_STRATEGY_REGISTER) bins = size_t(kRegGroups) * kMaxAcc * 4;
      geo = Geo{es, en, ec, est, uint32_t(et), align_up(fixed_bytes + bins, 128)};
      have = true;
    }
  }
  if (!have) { set_error(forced >= 0 ? "forced strategy does not fit in shared memory" : "stage ring does not fit in shared memory"); return HDK_B200_E_UNSUPPORTED; }
  const int strategy = geo.strategy;
  a.consumer_threads = uint32_t(geo.nct);
  a.n_stages = uint32_t(geo.stages);
  a.tile_rows = geo.tile_rows;
  a.off_bins = uint32_t(fixed_bytes);
  a.off_stages = uint32_t(geo.off_stages);
  {
    size_t bins = 0;
    for (int i = 0; i < p.n_acc; ++i) {
      bins = align_up(bins, 16);
      a.acc_bin_off[i] = uint32_t(bins);
      if (strategy == HDK_B200_STRATEGY_THREAD_PRIVATE) bins += size_t(p.accs[i].bytes) * E * geo.nct;
      else if (strategy == HDK_B200_STRATEGY_CTA_SHARED) bins += size_t(p.accs[i].bytes) * E;
    }
  }
  size_t so = 0;
  for (int c = 0; c < p.n_cols; ++c) {
    a.col_region_off[c] = uint32_t(so);
    so += align_up(size_t(geo.tile_rows) * p.col_width[c] + 32, 16);  // +16 for the alignment shift, +16 slack
  }
  a.stage_bytes = uint32_t(align_up(so, 128));
  const size_t smem_bytes = a.off_stages + size_t(a.stage_bytes) * geo.stages;
  if (smem_bytes > size_t(max_smem) - 1024) { set_error("stage ring does not fit in shared memory (%zu > %d)", smem_bytes, max_smem - 1024); return HDK_B200_E_UNSUPPORTED; }
  const int block = geo.nct + 32;
  int grid = sm_count() * geo.ctas;
  if (ko && ko->gridDimX) grid = int(ko->gridDimX);

  if (strategy == HDK_B200_STRATEGY_REGISTER && !(stat && stat->fn[strategy])) { set_error("REGISTER strategy needs a pre-compiled shape"); return HDK_B200_E_UNSUPPORTED; }
  ScanKernelFn kern = strategy == HDK_B200_STRATEGY_THREAD_PRIVATE ? scan_kernel<HDK_B200_STRATEGY_THREAD_PRIVATE, GenericShape>
                      : strategy == HDK_B200_STRATEGY_CTA_SHARED   ? scan_kernel<HDK_B200_STRATEGY_CTA_SHARED, GenericShape>
                      : strategy == HDK_B200_STRATEGY_GLOBAL       ? scan_kernel<HDK_B200_STRATEGY_GLOBAL, GenericShape>
                                                                   : scan_kernel<HDK_B200_STRATEGY_BASELINE, GenericShape>;
  int variant = 0;
  if (stat && stat->fn[strategy]) {
    kern = stat->fn[strategy];
    if (strategy == HDK_B200_STRATEGY_REGISTER) kern = stat->reg_fn[E <= 2 ? 0 : E <= 4 ? 1 : E <= 6 ? 2 : 3];
    variant = (stat >= kStaticShapes && stat < kStaticShapes + sizeof(kStaticShapes) / sizeof(kStaticShapes[0])) ? int(stat - kStaticShapes) + 1
                                                                                                                    : HDK_B200_VARIANT_JIT;
    const uint32_t per_iter = uint32_t(geo.nct) * uint32_t(stat->iter_rows);
    a.full_iters = (geo.tile_rows % per_iter == 0 && geo.tile_rows % 16 == 0) ? geo.tile_rows / per_iter : 0;
  }
  HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 1024));   // (a cap, never lowered: cached geometries of the same kernel stay launchable)
  if (!(ko && ko->gridDimX)) {   // persistent grid = what is actually resident (registers can allow fewer CTAs than planned)
    int occ = 0;
    HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem_bytes));
    grid = sm_count() * std::max(1, std::min(occ, geo.ctas));
  }
  if (!geo_env()) {
    std::lock_guard<std::mutex> lock(g_geo_mutex);
    if (g_geo_cache.size() >= 256) g_geo_cache.clear();
    GeoEntry ge{};
    ge.key = key; ge.a = a; ge.kern = kern; ge.grid = grid; ge.block = block; ge.smem_bytes = smem_bytes; ge.variant = variant; ge.strategy = strategy;
    g_geo_cache.push_back(ge);
  }
  {
    // (cudaLaunchKernel rather than <<<>>>: a run-time compiled kernel is a cudaKernel_t handle, not a host stub to call)
    void* kargs[] = {&a};
    HB_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(block), kargs, smem_bytes, stream));
  }
  HB_LAUNCH_CHECK();
  if (info) {
    info->variant = variant;
    info->strategy = strategy;
    info->grid = grid;
    info->block = block;
    info->smem_bytes = int(smem_bytes);
    info->n_accumulators = p.n_acc;
    info->tile_rows = int(geo.tile_rows);
  }
  return HDK_B200_OK;
}

int launch_scan(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                int64_t* work_table, cudaStream_t stream, hdk_b200_launch_info* info) {
  return launch_scan_impl(lw, ko, params, work_table, false, stream, info);
}
int launch_scan_exchange(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                         int64_t* work_table, const ExchangeTargets& x, cudaStream_t stream, hdk_b200_launch_info* info) {
  return launch_scan_impl(lw, ko, params, work_table, false, stream, info, &x);
}
int launch_baseline_scan(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                         int64_t* work_table, cudaStream_t stream, hdk_b200_launch_info* info, const int* run_if) {
  return launch_scan_impl(lw, ko, params, work_table, true, stream, info, nullptr, run_if);
}

}  // namespace hb
