// hdk_b200/csrc/api.cu — C-ABI entry points of the launch path (include/hdk_b200.h).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "partagg.cuh"
#include "scan.cuh"

namespace hb {
int init_group_by_buffer(const Lowered& lw, int64_t* buf, cudaStream_t stream);
}

extern "C" {

// Should this baseline-hash launch run radix-partitioned (partagg.cu)?  Eligible plan, row count known, and either forced
// or a table far beyond L2 (the per-row global probe then costs a random DRAM sector and several global atomics).
static bool wants_partitioned(const hb::Lowered& lw, uint64_t total_rows, size_t* need) {
  *need = 0;
  if (lw.plan.hash_type != HDK_B200_BASELINE_HASH || hb::g_debug.partitioned == 0 || total_rows == 0) return false;
  if (hb::partagg_scratch_bytes(lw, total_rows, need) != HDK_B200_OK) return false;
  if (hb::g_debug.partitioned == 1) return true;
  const size_t table_bytes = size_t(lw.plan.entry_count) * (lw.layout.row_bytes + size_t(lw.plan.n_acc) * 8);
  return table_bytes > (size_t(48) << 20) && total_rows >= (1u << 20);
}

int hdk_b200_launch_scratch_bytes(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, uint64_t total_rows, size_t* scratch_bytes) {
  hb::Lowered lw;
  if (int rc = hb::lower_plan(plan, qmd, &lw)) return rc;
  size_t need = 0;
  const bool pa = wants_partitioned(lw, total_rows, &need);
  if (scratch_bytes) *scratch_bytes = pa ? std::max(need, lw.work_table_bytes) : lw.work_table_bytes;
  return HDK_B200_OK;
}

int hdk_b200_init_work_table(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, int64_t* work_table, void* stream) {
  hb::Lowered lw;
  if (int rc = hb::lower_plan(plan, qmd, &lw)) return rc;
  if (qmd->hash_type != HDK_B200_PERFECT_HASH) { hb::set_error("work tables exist for perfect-hash plans only"); return HDK_B200_E_UNSUPPORTED; }
  return hb::init_work_table(lw, work_table, static_cast<cudaStream_t>(stream));
}

int hdk_b200_launch_partial(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_options* ko,
                            const hdk_b200_kernel_params* params, int64_t* work_table, void* stream,
                            hdk_b200_launch_info* info) {
  hb::Lowered lw;
  if (int rc = hb::lower_plan(plan, qmd, &lw)) return rc;
  if (qmd->hash_type != HDK_B200_PERFECT_HASH) { hb::set_error("launch_partial is for perfect-hash plans"); return HDK_B200_E_UNSUPPORTED; }
  if (!params || !work_table) { hb::set_error("null params/work table"); return HDK_B200_E_INVALID; }
  if (info) memset(info, 0, sizeof(*info));
  if (int rc = hb::launch_scan(lw, ko, params, work_table, static_cast<cudaStream_t>(stream), info)) return rc;
  if (info) info->n_launches = 1;
  return HDK_B200_OK;
}

int hdk_b200_finalize(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const int64_t* work_table,
                      int64_t* groups_buffer, void* stream) {
  hb::Lowered lw;
  if (int rc = hb::lower_plan(plan, qmd, &lw)) return rc;
  if (qmd->hash_type != HDK_B200_PERFECT_HASH) { hb::set_error("finalize is for perfect-hash plans"); return HDK_B200_E_UNSUPPORTED; }
  return hb::launch_finalize(lw, work_table, groups_buffer, nullptr, static_cast<cudaStream_t>(stream));
}

int hdk_b200_launch(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_options* ko,
                    const hdk_b200_kernel_params* params, void* scratch, size_t scratch_bytes, void* stream,
                    hdk_b200_launch_info* info) {
  hb::Lowered lw;
  if (int rc = hb::lower_plan(plan, qmd, &lw)) return rc;
  if (!params || !params->groupby_buf || !params->error_codes) { hb::set_error("null kernel params"); return HDK_B200_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (info) memset(info, 0, sizeof(*info));
  if (scratch_bytes < lw.work_table_bytes || !scratch) {
    hb::set_error("scratch too small: need %zu bytes, got %zu", lw.work_table_bytes, scratch_bytes);
    return HDK_B200_E_INVALID;
  }
  int64_t* work = static_cast<int64_t*>(scratch);
  if (qmd->hash_type == HDK_B200_BASELINE_HASH) {
    size_t need = 0;
    if (wants_partitioned(lw, params->total_rows_hint, &need) && scratch_bytes >= need) {
      const int* fallback = nullptr;
      int64_t* fwork = nullptr;
      if (int rc = hb::launch_partagg(lw, params, scratch, scratch_bytes, st, info, &fallback, &fwork)) return rc;
      // hot keys (a partition too heavy for one CTA): the partitioned kernels stood down, the global-table path below
      // does the work; otherwise these three launches return at once
      if (int rc = hb::init_work_table(lw, fwork, st, fallback)) return rc;
      if (int rc = hb::launch_baseline_scan(lw, ko, params, fwork, st, nullptr, fallback)) return rc;
      if (int rc = hb::launch_finalize(lw, fwork, nullptr, params->groupby_buf, st, fallback)) return rc;
      if (info) info->n_launches += 3;
      return HDK_B200_OK;
    }
    // keys are claimed in the caller-initialised buffer, aggregates accumulate in the entry-major work table,
    // finalize encodes the slots of the claimed entries
    if (int rc = hb::init_work_table(lw, work, st)) return rc;
    if (int rc = hb::launch_baseline_scan(lw, ko, params, work, st, info)) return rc;
    if (int rc = hb::launch_finalize(lw, work, nullptr, params->groupby_buf, st)) return rc;
    if (info) info->n_launches = 3;
    return HDK_B200_OK;
  }
  // GROUPBY_BUF is a device array of pointers: finalize dereferences it on the device, no host sync
  if (int rc = hb::init_work_table(lw, work, st)) return rc;
  if (int rc = hb::launch_scan(lw, ko, params, work, st, info)) return rc;
  if (int rc = hb::launch_finalize(lw, work, nullptr, params->groupby_buf, st)) return rc;
  if (info) info->n_launches = 3;
  return HDK_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer end-to-end call: H2D of the chunks, init, launch, D2H of buffer + error code
// ---------------------------------------------------------------------------------------------
#define HB_CUDA_HOST(expr)                                                                         \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      hb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);   \
      rc = HDK_B200_E_CUDA - 1000;                                                                 \
      goto done;                                                                                   \
    }                                                                                              \
  } while (0)

int hdk_b200_query_host(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const int8_t* const* col_buffers,
                        const int64_t* num_rows, uint64_t num_fragments, const int64_t* join_hash_tables_host,
                        const size_t* join_table_bytes, const int8_t* const* inner_col_buffers,
                        const size_t* inner_col_bytes, int8_t* out_buffer, int device, hdk_b200_launch_info* info) {
  hb::Lowered lw;
  if (int r0 = hb::lower_plan(plan, qmd, &lw)) return r0 - 1000;
  int rc = 0;
  std::vector<void*> allocs;
  cudaStream_t st = nullptr;
  const size_t out_bytes = hdk_b200_buffer_size_bytes(qmd);
  const int nc = plan->n_cols;
  std::vector<const int8_t*> dcols(size_t(num_fragments) * nc, nullptr);
  std::vector<int64_t> djoin(HDK_B200_MAX_JOINS, 0);
  std::vector<const int8_t*> dinner(size_t(HDK_B200_MAX_JOINS) * HDK_B200_MAX_COLS, nullptr);
  int8_t *d_colptrs = nullptr, *d_numrows = nullptr, *d_buf = nullptr, *d_bufptr = nullptr, *d_err = nullptr,
         *d_scratch = nullptr, *d_join = nullptr, *d_inner = nullptr;
  int32_t err_host = 0;
  auto dalloc = [&](size_t bytes, int8_t** out) -> cudaError_t {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(out), bytes ? bytes : 16);
    if (e == cudaSuccess) allocs.push_back(*out);
    return e;
  };
  hdk_b200_kernel_params kp;
  memset(&kp, 0, sizeof(kp));

  HB_CUDA_HOST(cudaSetDevice(device));
  HB_CUDA_HOST(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (uint64_t f = 0; f < num_fragments; ++f)
    for (int c = 0; c < nc; ++c) {
      int8_t* d = nullptr;
      const size_t bytes = size_t(num_rows[f]) * lw.plan.col_width[c];
      HB_CUDA_HOST(dalloc(bytes, &d));
      HB_CUDA_HOST(cudaMemcpyAsync(d, col_buffers[f * nc + c], bytes, cudaMemcpyHostToDevice, st));
      dcols[f * nc + c] = d;
    }
  for (int j = 0; j < plan->n_joins; ++j) {
    int8_t* d = nullptr;
    HB_CUDA_HOST(dalloc(join_table_bytes[j], &d));
    HB_CUDA_HOST(cudaMemcpyAsync(d, reinterpret_cast<const void*>(join_hash_tables_host[j]), join_table_bytes[j], cudaMemcpyHostToDevice, st));
    djoin[j] = reinterpret_cast<int64_t>(d);
    for (int c = 0; c < HDK_B200_MAX_COLS; ++c) {
      const size_t idx = size_t(j) * HDK_B200_MAX_COLS + c;
      if (!inner_col_buffers || !inner_col_buffers[idx]) continue;
      int8_t* dc = nullptr;
      HB_CUDA_HOST(dalloc(inner_col_bytes[idx], &dc));
      HB_CUDA_HOST(cudaMemcpyAsync(dc, inner_col_buffers[idx], inner_col_bytes[idx], cudaMemcpyHostToDevice, st));
      dinner[idx] = dc;
    }
  }
  HB_CUDA_HOST(dalloc(dcols.size() * sizeof(void*), &d_colptrs));
  HB_CUDA_HOST(cudaMemcpyAsync(d_colptrs, dcols.data(), dcols.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
  HB_CUDA_HOST(dalloc(num_fragments * sizeof(int64_t), &d_numrows));
  HB_CUDA_HOST(cudaMemcpyAsync(d_numrows, num_rows, num_fragments * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  HB_CUDA_HOST(dalloc(djoin.size() * sizeof(int64_t), &d_join));
  HB_CUDA_HOST(cudaMemcpyAsync(d_join, djoin.data(), djoin.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  HB_CUDA_HOST(dalloc(dinner.size() * sizeof(void*), &d_inner));
  HB_CUDA_HOST(cudaMemcpyAsync(d_inner, dinner.data(), dinner.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
  HB_CUDA_HOST(dalloc(out_bytes, &d_buf));
  HB_CUDA_HOST(dalloc(sizeof(void*), &d_bufptr));
  HB_CUDA_HOST(cudaMemcpyAsync(d_bufptr, &d_buf, sizeof(void*), cudaMemcpyHostToDevice, st));
  HB_CUDA_HOST(dalloc(sizeof(int32_t), &d_err));
  HB_CUDA_HOST(cudaMemsetAsync(d_err, 0, sizeof(int32_t), st));
  HB_CUDA_HOST(dalloc(lw.work_table_bytes, &d_scratch));

  kp.col_buffers = reinterpret_cast<const int8_t* const*>(d_colptrs);
  kp.num_fragments = num_fragments;
  kp.num_rows = reinterpret_cast<const int64_t*>(d_numrows);
  kp.groupby_buf = reinterpret_cast<int64_t* const*>(d_bufptr);
  kp.error_codes = reinterpret_cast<int32_t*>(d_err);
  kp.num_tables = uint32_t(1 + plan->n_joins);
  kp.join_hash_tables = reinterpret_cast<const int64_t*>(d_join);
  kp.inner_col_buffers = reinterpret_cast<const int8_t* const*>(d_inner);
  {
    int r = hdk_b200_init_group_by_buffer(qmd, reinterpret_cast<int64_t*>(d_buf), st);
    if (!r) r = hdk_b200_launch(plan, qmd, nullptr, &kp, d_scratch, lw.work_table_bytes, st, info);
    if (r) { rc = r - 1000; goto done; }
  }
  HB_CUDA_HOST(cudaMemcpyAsync(out_buffer, d_buf, out_bytes, cudaMemcpyDeviceToHost, st));
  HB_CUDA_HOST(cudaMemcpyAsync(&err_host, d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  HB_CUDA_HOST(cudaStreamSynchronize(st));
  rc = err_host;
done:
  for (void* p : allocs) cudaFree(p);
  if (st) cudaStreamDestroy(st);
  return rc;
}

}  // extern "C"
