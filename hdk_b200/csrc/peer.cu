// hdk_b200/csrc/peer.cu — multi-GPU merge of perfect-hash partials over peer memory (NVLink / NVSwitch).
//
// One process per GPU.  Instead of an NCCL all-reduce between the scan and the finalize kernels, the scan kernel's
// last CTA copies this GPU's neutral partial table into every peer's exchange buffer (plain 16-byte stores through
// NVLink, then a flag), and every GPU's finalize kernel waits for the flags, merges the n_peers tables and encodes
// the result (see include/hdk_b200.h).  The reference has no counterpart: it merges per-device ResultSets on the
// host (Executor::reduceMultiDeviceResultSets, QE/Execute.cpp:1224-1336).
#include <cstring>

#include "common.cuh"
#include "scan.cuh"

namespace hb {

constexpr size_t kFlagBytes = 2 * HDK_B200_MAX_PEERS * sizeof(unsigned long long);   // flags[parity][rank]

static size_t slot_cells_of(const Lowered& lw) { return (size_t(lw.plan.n_acc) * lw.plan.entry_count + 1) & ~size_t(1); }

}  // namespace hb

using namespace hb;

extern "C" {

int hdk_b200_peer_alloc(size_t bytes, void** dev_ptr, uint8_t handle[HDK_B200_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == HDK_B200_IPC_HANDLE_BYTES, "IPC handle size");
  if (!dev_ptr || !handle || !bytes) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  HB_CUDA(cudaMalloc(dev_ptr, bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
  if (e != cudaSuccess) {
    cudaFree(*dev_ptr);
    *dev_ptr = nullptr;
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return HDK_B200_E_CUDA;
  }
  memcpy(handle, &h, sizeof(h));
  return HDK_B200_OK;
}

int hdk_b200_peer_open(const uint8_t handle[HDK_B200_IPC_HANDLE_BYTES], void** dev_ptr) {
  if (!dev_ptr || !handle) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  HB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HDK_B200_OK;
}

int hdk_b200_peer_close(void* dev_ptr) {
  if (dev_ptr) HB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return HDK_B200_OK;
}

int hdk_b200_peer_free(void* dev_ptr) {
  if (dev_ptr) HB_CUDA(cudaFree(dev_ptr));
  return HDK_B200_OK;
}

int hdk_b200_exchange_bytes(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, int n_peers, size_t* bytes) {
  Lowered lw;
  if (int rc = lower_plan(plan, qmd, &lw)) return rc;
  if (qmd->hash_type != HDK_B200_PERFECT_HASH) { set_error("the peer exchange merges perfect-hash partials"); return HDK_B200_E_UNSUPPORTED; }
  if (n_peers < 1 || n_peers > HDK_B200_MAX_PEERS || !bytes) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  *bytes = kFlagBytes + 2 * size_t(n_peers) * slot_cells_of(lw) * 8;
  return HDK_B200_OK;
}

int hdk_b200_exchange_init(void* local_exchange, void* stream) {
  if (!local_exchange) { set_error("null buffer"); return HDK_B200_E_INVALID; }
  HB_CUDA(cudaMemsetAsync(local_exchange, 0, kFlagBytes, static_cast<cudaStream_t>(stream)));
  return HDK_B200_OK;
}

int hdk_b200_launch_exchange(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_options* ko,
                             const hdk_b200_kernel_params* params, void* scratch, size_t scratch_bytes, void* const* peer_exchange,
                             int n_peers, int my_rank, uint64_t epoch, void* stream, hdk_b200_launch_info* info) {
  Lowered lw;
  if (int rc = lower_plan(plan, qmd, &lw)) return rc;
  if (qmd->hash_type != HDK_B200_PERFECT_HASH) { set_error("the peer exchange merges perfect-hash partials"); return HDK_B200_E_UNSUPPORTED; }
  if (!params || !params->groupby_buf || !params->error_codes || !peer_exchange) { set_error("null argument"); return HDK_B200_E_INVALID; }
  if (n_peers < 1 || n_peers > HDK_B200_MAX_PEERS || my_rank < 0 || my_rank >= n_peers || epoch == 0) { set_error("bad peer arguments"); return HDK_B200_E_INVALID; }
  const size_t table = (lw.work_table_bytes + 15) & ~size_t(15);
  if (!scratch || scratch_bytes < table + 64) { set_error("scratch too small: need %zu bytes, got %zu", table + 64, scratch_bytes); return HDK_B200_E_INVALID; }
  for (int r = 0; r < n_peers; ++r)
    if (!peer_exchange[r]) { set_error("peer %d: null exchange buffer", r); return HDK_B200_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (info) memset(info, 0, sizeof(*info));
  int64_t* work = static_cast<int64_t*>(scratch);
  const size_t cells = slot_cells_of(lw);
  const uint64_t parity = epoch & 1;
  ExchangeTargets x{};
  x.n_peers = uint32_t(n_peers);
  x.epoch = epoch;
  x.ticket = reinterpret_cast<unsigned long long*>(static_cast<int8_t*>(scratch) + table);
  for (int r = 0; r < n_peers; ++r) {
    int8_t* base = static_cast<int8_t*>(peer_exchange[r]);
    x.peer_flag[r] = reinterpret_cast<unsigned long long*>(base) + parity * HDK_B200_MAX_PEERS + my_rank;
    x.peer_slot[r] = reinterpret_cast<int64_t*>(base + kFlagBytes) + (parity * size_t(n_peers) + size_t(my_rank)) * cells;
  }
  HB_CUDA(cudaMemsetAsync(x.ticket, 0, 8, st));
  if (int rc = init_work_table(lw, work, st)) return rc;
  if (int rc = launch_scan_exchange(lw, ko, params, work, x, st, info)) return rc;
  int8_t* mine = static_cast<int8_t*>(peer_exchange[my_rank]);
  const int64_t* slots = reinterpret_cast<const int64_t*>(mine + kFlagBytes) + parity * size_t(n_peers) * cells;
  const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(mine) + parity * HDK_B200_MAX_PEERS;
  if (int rc = launch_finalize_exchange(lw, slots, flags, uint32_t(n_peers), epoch, params->error_codes, params->groupby_buf, st)) return rc;
  if (info) info->n_launches = 3;
  return HDK_B200_OK;
}

}  // extern "C"
