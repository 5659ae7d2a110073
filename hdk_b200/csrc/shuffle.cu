// hdk_b200/csrc/shuffle.cu — hash partitioning of rows by group key: the two passes of partitioned
// aggregation (count, scatter) that feed the NCCL all-to-all of the multi-GPU baseline-hash group-by.
//
// Model: the reference's CPU-only partitioned aggregation (QE/RelAlgExecutor.cpp:691-838):
//   partition function  MurmurHash64A over the 64-bit-widened key components (QE/RowFuncBuilder.cpp:516-577);
//                       the reference masks with (P-1), here `% n_partitions` so any GPU count works
//   pass 1 (COUNT)      per-partition histogram, reduced over kernels (QE/Execute.cpp:1365-1435)
//   pass 2 (scatter)    rows copied into pre-sized per-partition columnar buffers (QE/Execute.cpp:1933-2018)
#include <algorithm>

#include "baseline.cuh"
#include "common.cuh"
#include "eval.cuh"

namespace hb {

constexpr uint32_t kMaxPartitions = 1024;

struct ShuffleArgs {
  DPlan plan;
  const int8_t* const* col_buffers;
  const int64_t* num_rows;
  uint32_t num_fragments;
  uint32_t n_partitions;
  unsigned long long* counts;      // pass 1
  const uint64_t* offsets;         // pass 2
  unsigned long long* cursors;     // pass 2
  int8_t* const* out_cols;         // pass 2
};

// partition of one row, or -1 when the row is filtered out
__device__ __forceinline__ int row_partition(const ShuffleArgs& a, const int8_t* const* cols, uint64_t pos, V* vals) {
  const DPlan& p = a.plan;
  auto load_outer = [&](int c, int w) -> uint64_t {
    const int8_t* ptr = cols[c] + pos * w;
    return w == 8 ? *reinterpret_cast<const uint64_t*>(ptr) : w == 4 ? uint64_t(*reinterpret_cast<const uint32_t*>(ptr))
           : w == 2 ? uint64_t(*reinterpret_cast<const uint16_t*>(ptr)) : uint64_t(*reinterpret_cast<const uint8_t*>(ptr));
  };
  auto load_inner = [&](int, int, int) -> uint64_t { return 0; };
  int32_t err = 0;
  for (int n = 0; n < p.n_exprs; ++n) vals[n] = eval_node(p, p.exprs[n], vals, err, load_outer, load_inner);
  for (int f = 0; f < p.n_filters; ++f)
    if (!(vals[p.filters[f]].i > 0)) return -1;
  int64_t keys[HDK_B200_MAX_KEYS];
  for (int k = 0; k < p.n_keys; ++k) keys[k] = vals[p.keys[k].expr].i;
  return int(murmur64a_keys(keys, p.n_keys) % a.n_partitions);
}

__global__ void shuffle_count_kernel(const __grid_constant__ ShuffleArgs a) {
  __shared__ unsigned int hist[kMaxPartitions];
  for (uint32_t i = threadIdx.x; i < a.n_partitions; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  V vals[HDK_B200_MAX_EXPRS];
  const uint64_t start = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint32_t f = 0; f < a.num_fragments; ++f) {
    const int8_t* const* cols = a.col_buffers + size_t(f) * a.plan.n_cols;
    const uint64_t rows = uint64_t(a.num_rows[f]);
    for (uint64_t pos = start; pos < rows; pos += step) {
      const int part = row_partition(a, cols, pos, vals);
      if (part >= 0) atomicAdd(&hist[part], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < a.n_partitions; i += blockDim.x)
    if (hist[i]) atomicAdd(a.counts + i, (unsigned long long)hist[i]);
}

__global__ void shuffle_scatter_kernel(const __grid_constant__ ShuffleArgs a) {
  V vals[HDK_B200_MAX_EXPRS];
  const DPlan& p = a.plan;
  const uint64_t start = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (uint32_t f = 0; f < a.num_fragments; ++f) {
    const int8_t* const* cols = a.col_buffers + size_t(f) * p.n_cols;
    const uint64_t rows = uint64_t(a.num_rows[f]);
    const uint64_t rows_padded = (rows + step - 1) / step * step;  // keep warps converged for the match
    for (uint64_t pos = start; pos < rows_padded; pos += step) {
      const int part = pos < rows ? row_partition(a, cols, pos, vals) : -1;
      // warp-aggregated reservation: one atomic per distinct partition per warp
      const unsigned peers = __match_any_sync(0xffffffffu, part);
      if (part < 0) continue;
      const int leader = __ffs(peers) - 1;
      const int rank = __popc(peers & ((1u << lane) - 1));
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(a.cursors + part, (unsigned long long)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      const uint64_t dst = a.offsets[part] + base + rank;
      for (int c = 0; c < p.n_cols; ++c) {
        const int w = p.col_width[c];
        const int8_t* src = cols[c] + pos * w;
        int8_t* out = a.out_cols[c] + dst * w;
        if (w == 8) *reinterpret_cast<uint64_t*>(out) = *reinterpret_cast<const uint64_t*>(src);
        else if (w == 4) *reinterpret_cast<uint32_t*>(out) = *reinterpret_cast<const uint32_t*>(src);
        else if (w == 2) *reinterpret_cast<uint16_t*>(out) = *reinterpret_cast<const uint16_t*>(src);
        else *out = *src;
      }
    }
  }
}

// a plan without layout information: lower only what the shuffle needs
static int lower_for_shuffle(const hdk_b200_plan* plan, Lowered* lw) {
  hdk_b200_qmd q{};
  q.hash_type = HDK_B200_BASELINE_HASH;
  q.entry_count = 1;
  q.key_count = plan ? plan->n_keys : 0;
  q.key_width = 8;
  // slots: fabricate 8-byte slots so that lower_plan accepts every target
  hdk_b200_plan p2 = *plan;
  int slot = 0;
  for (int t = 0; t < p2.n_targets; ++t) {
    if (p2.targets[t].agg == HDK_B200_AGG_NONE) { p2.targets[t].slot = -1; continue; }
    p2.targets[t].slot = slot;
    const int n = p2.targets[t].agg == HDK_B200_AGG_AVG ? 2 : 1;
    for (int i = 0; i < n; ++i) q.slot_padded[slot++] = 8;
  }
  q.slot_count = slot;
  return lower_plan(&p2, &q, lw);
}

}  // namespace hb

using namespace hb;

extern "C" {

int hdk_b200_shuffle_count(const hdk_b200_plan* plan, const hdk_b200_kernel_params* params, uint32_t n_partitions,
                           uint64_t* counts, void* stream) {
  if (!plan || !params || !counts || n_partitions == 0 || n_partitions > kMaxPartitions) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  if (plan->n_joins) { set_error("shuffle of joined plans is not supported"); return HDK_B200_E_UNSUPPORTED; }
  Lowered lw;
  if (int rc = lower_for_shuffle(plan, &lw)) return rc;
  ShuffleArgs a{};
  a.plan = lw.plan;
  a.col_buffers = params->col_buffers;
  a.num_rows = params->num_rows;
  a.num_fragments = uint32_t(params->num_fragments);
  a.n_partitions = n_partitions;
  a.counts = reinterpret_cast<unsigned long long*>(counts);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * n_partitions, st));
  shuffle_count_kernel<<<sm_count() * 4, 256, 0, st>>>(a);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_shuffle_scatter(const hdk_b200_plan* plan, const hdk_b200_kernel_params* params, uint32_t n_partitions,
                             const uint64_t* offsets, uint64_t* cursors, int8_t* const* out_cols, void* stream) {
  if (!plan || !params || !offsets || !cursors || !out_cols || n_partitions == 0 || n_partitions > kMaxPartitions) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  if (plan->n_joins) { set_error("shuffle of joined plans is not supported"); return HDK_B200_E_UNSUPPORTED; }
  Lowered lw;
  if (int rc = lower_for_shuffle(plan, &lw)) return rc;
  ShuffleArgs a{};
  a.plan = lw.plan;
  a.col_buffers = params->col_buffers;
  a.num_rows = params->num_rows;
  a.num_fragments = uint32_t(params->num_fragments);
  a.n_partitions = n_partitions;
  a.offsets = offsets;
  a.cursors = reinterpret_cast<unsigned long long*>(cursors);
  a.out_cols = out_cols;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CUDA(cudaMemsetAsync(cursors, 0, sizeof(uint64_t) * n_partitions, st));
  shuffle_scatter_kernel<<<sm_count() * 4, 256, 0, st>>>(a);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

}  // extern "C"
