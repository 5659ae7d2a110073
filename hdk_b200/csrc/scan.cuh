// hdk_b200/csrc/scan.cuh — launch geometry and argument block of the fused scan kernel.
#pragma once
#include "common.cuh"

namespace hb {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;  // + one TMA producer warp
constexpr int kStages = 8;   // ring depth limit (the geometry search uses 3-4)
constexpr uint32_t kMaxFragments = 4096;

struct AccKinds {
  uint8_t kind[kMaxAcc];
};

struct ScanArgs {
  DPlan plan;
  DLayout layout;                     // baseline hash updates the reference-encoded buffer in place
  const int8_t* const* col_buffers;   // [num_fragments * n_cols]
  const int64_t* num_rows;            // [num_fragments]
  const int64_t* join_hash_tables;    // [n_joins]
  const int8_t* const* inner_col_buffers;
  int64_t* work_table;                // [n_acc * entry_count]            (perfect hash)
  int64_t* const* groupby_buf;        // GROUPBY_BUF device array, [0] used  (baseline hash)
  int32_t* error_codes;
  const int* run_if;                  // when set: the kernel does nothing unless *run_if != 0 (fallback path of the partitioned aggregation)
  uint32_t num_fragments;
  uint32_t tile_rows;
  uint32_t stage_bytes;
  uint32_t n_stages;                  // ring depth actually used (<= kStages)
  uint32_t full_iters;                // static shapes: iterations of the full-tile loop (tile_rows / (consumer_threads * iter rows)); 0 = off
  uint32_t consumer_threads;          // consumer threads per CTA (multiple of 32, <= kConsumerThreads); blockDim = that + 32
  uint32_t off_tile_prefix, off_bins, off_stages;   // dynamic shared memory map
  // multi-GPU exchange (hdk_b200_launch_exchange): the last CTA publishes the work table to every peer
  uint32_t n_peers;                   // 0 = single-GPU launch
  uint64_t n_cells;                   // n_acc * entry_count
  uint64_t epoch;
  unsigned long long* ticket;         // CTAs-done counter behind the work table, zeroed with it
  int64_t* peer_slot[HDK_B200_MAX_PEERS];              // slot `my_rank` of peer p's exchange buffer (this epoch's parity)
  unsigned long long* peer_flag[HDK_B200_MAX_PEERS];   // flag `my_rank` of peer p
  uint8_t join_order[HDK_B200_MAX_JOINS];           // joins by key node: the order in which the interpreting kernel nests their probes
  uint32_t acc_bin_off[kMaxAcc];                    // from off_bins
  uint32_t col_region_off[HDK_B200_MAX_COLS];       // from a stage's base
};

#ifndef __CUDACC_RTC__
typedef void (*ScanKernelFn)(const ScanArgs);
struct StaticEntry {   // the kernels of one plan shape: pre-compiled (static_shapes.inc) or compiled at run time (jit.cu)
  uint64_t sig;
  const char* name;
  int iter_rows;       // rows per consumer thread per iteration of the full-tile loop
  ScanKernelFn fn[5];  // THREAD_PRIVATE, CTA_SHARED, GLOBAL, BASELINE, REGISTER (8 groups)
  ScanKernelFn reg_fn[4];  // REGISTER kernels for <= 2, 4, 6, 8 groups
};
// Run-time specialisation (jit.cu): the kernels of a plan shape without a pre-compiled entry, compiled with NVRTC from the
// same scan_kernel.cuh.  `wait`: compile now if needed (else the compile runs on a worker thread and nullptr is returned
// until it is done — the caller launches the interpreting kernel meanwhile).  nullptr also when NVRTC is unavailable.
const StaticEntry* jit_scan_kernels(const DPlan& p, uint64_t sig, bool wait);
int dump_shape_text(const DPlan& p, char* out, size_t cap);
bool same_plan_shape(const DPlan& a, const DPlan& b);
struct JitStats { unsigned long long compiled, failed, launches, pending; double last_compile_ms, total_compile_ms; };
extern JitStats g_jit_stats;

int init_work_table(const Lowered& lw, int64_t* work_table, cudaStream_t stream, const int* run_if = nullptr);
int launch_scan(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                int64_t* work_table, cudaStream_t stream, hdk_b200_launch_info* info);
int launch_finalize(const Lowered& lw, const int64_t* work_table, int64_t* groups_buffer, int64_t* const* groups_buffer_indirect,
                    cudaStream_t stream, const int* run_if = nullptr);
// exchange variants (peer.cu)
struct ExchangeTargets {
  uint32_t n_peers;
  uint64_t epoch;
  unsigned long long* ticket;
  int64_t* peer_slot[HDK_B200_MAX_PEERS];
  unsigned long long* peer_flag[HDK_B200_MAX_PEERS];
};
int launch_scan_exchange(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                         int64_t* work_table, const ExchangeTargets& x, cudaStream_t stream, hdk_b200_launch_info* info);
int launch_finalize_exchange(const Lowered& lw, const int64_t* slots, const unsigned long long* flags, uint32_t n_peers, uint64_t epoch,
                             int32_t* error_codes, int64_t* const* groups_buffer_indirect, cudaStream_t stream);
int launch_baseline_scan(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                         int64_t* work_table, cudaStream_t stream, hdk_b200_launch_info* info, const int* run_if = nullptr);

#endif

}  // namespace hb
