// hdk_b200/csrc/finalize.cu — group-by buffer initialisation and the work-table → reference-encoding
// conversion of the perfect-hash path.
//
//   init      QueryMemoryInitializer::initRowGroups / initColumnarGroups (QE/QueryMemoryInitializer.cpp:502-687),
//             init_group_by_buffer_gpu / init_columnar_group_by_buffer_gpu (QE/GpuInitGroups.cu:20-167)
//   finalize  what the reference's per-row runtime leaves in the buffer after the last row:
//             keys written by get_group_value_fast / get_matching_group_value_perfect_hash
//             (QE/GroupByRuntime.cpp:198-213, QE/RuntimeFunctions.cpp:1339-1382), slots by agg_* with the
//             null-sentinel ("skip_val") protocol (QE/RuntimeFunctions.cpp:612-703, 821-880) starting
//             from get_agg_initial_val (QE/OutputBufferInitialization.cpp:112-258).
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"

namespace hb {

static size_t align8h(size_t x) { return (x + 7) & ~size_t(7); }

// ---------------------------------------------------------------------------------------------
// init from the descriptor
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRowQuads = HDK_B200_MAX_KEYS + HDK_B200_MAX_SLOTS;

struct RowTemplate {
  int64_t quad[kMaxRowQuads];
  uint32_t n_quads;
};

__global__ void init_rowwise_kernel(int64_t* buf, uint64_t total_quads, const __grid_constant__ RowTemplate t) {
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total_quads; i += step) buf[i] = t.quad[i % t.n_quads];
}

struct ColumnarTemplate {
  uint64_t col_off[kMaxRowQuads];
  int64_t fill[kMaxRowQuads];
  uint8_t width[kMaxRowQuads];
  uint32_t n_cols;
  uint32_t entry_count;
};

__global__ void init_columnar_kernel(int8_t* buf, const __grid_constant__ ColumnarTemplate t) {
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint32_t c = 0; c < t.n_cols; ++c) {
    int8_t* col = buf + t.col_off[c];
    const int w = t.width[c];
    if (w == 8) {
      for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < t.entry_count; i += step)
        reinterpret_cast<int64_t*>(col)[i] = t.fill[c];
    } else if (w == 4) {
      for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < t.entry_count; i += step)
        reinterpret_cast<int32_t*>(col)[i] = int32_t(t.fill[c]);
    }
    // zero the ≤ 7 padding bytes after the column so that whole-buffer comparisons are deterministic
    const uint64_t used = uint64_t(w) * t.entry_count;
    const uint64_t padded = (used + 7) & ~uint64_t(7);
    if (blockIdx.x == 0 && threadIdx.x < padded - used) col[used + threadIdx.x] = 0;
  }
}

static int grid_for(uint64_t n, int block) {
  return int(std::max<uint64_t>(1, std::min<uint64_t>((n + block - 1) / block, uint64_t(sm_count()) * 16)));
}

int init_group_by_buffer(const Lowered& lw, int64_t* buf, cudaStream_t stream) {
  const DLayout& L = lw.layout;
  if (L.columnar) {
    ColumnarTemplate t{};
    uint32_t n = 0;
    const size_t E = L.entry_count;
    if (!L.keyless)
      for (int k = 0; k < L.key_count; ++k) {
        t.col_off[n] = size_t(k) * align8h(8 * E);
        t.fill[n] = HDK_B200_EMPTY_KEY_64;
        t.width[n++] = 8;
      }
    for (int s = 0; s < L.slot_count; ++s) {
      if (!L.slots[s].padded) continue;
      t.col_off[n] = L.slots[s].col_off;
      t.fill[n] = L.slots[s].init_val;
      t.width[n++] = L.slots[s].padded;
    }
    t.n_cols = n;
    t.entry_count = L.entry_count;
    init_columnar_kernel<<<grid_for(E, 256), 256, 0, stream>>>(reinterpret_cast<int8_t*>(buf), t);
    HB_LAUNCH_CHECK();
    return HDK_B200_OK;
  }
  RowTemplate t{};
  t.n_quads = L.row_bytes / 8;
  int8_t* row = reinterpret_cast<int8_t*>(t.quad);
  if (!L.keyless)
    for (int k = 0; k < L.key_count; ++k) {
      if (L.key_width == 4) reinterpret_cast<int32_t*>(row)[k] = HDK_B200_EMPTY_KEY_32;
      else reinterpret_cast<int64_t*>(row)[k] = HDK_B200_EMPTY_KEY_64;
    }
  for (int s = 0; s < L.slot_count; ++s) {
    const DSlot& sl = L.slots[s];
    if (sl.padded == 8) *reinterpret_cast<int64_t*>(row + L.key_bytes + sl.off) = sl.init_val;
    else if (sl.padded == 4) *reinterpret_cast<int32_t*>(row + L.key_bytes + sl.off) = int32_t(sl.init_val);
  }
  const uint64_t total = uint64_t(t.n_quads) * L.entry_count;
  init_rowwise_kernel<<<grid_for(total, 256), 256, 0, stream>>>(buf, total, t);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// 1:1 mirrors of the reference's initialisers (init values in device memory)
// ---------------------------------------------------------------------------------------------
__global__ void ref_init_group_by_buffer_kernel(int64_t* groups_buffer, const int64_t* init_vals, uint32_t entry_count,
                                                uint32_t key_count, uint32_t key_width, uint32_t row_size_quad,
                                                bool keyless, int8_t warp_size) {
  const uint64_t start = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  if (keyless) {
    const uint64_t n = uint64_t(entry_count) * row_size_quad * uint64_t(warp_size);
    for (uint64_t i = start; i < n; i += step) groups_buffer[i] = init_vals[i % row_size_quad];
    return;
  }
  const uint32_t values_off_quad = uint32_t(((uint64_t(key_count) * key_width + 7) & ~uint64_t(7)) / 8);
  for (uint64_t e = start; e < entry_count; e += step) {
    int64_t* row = groups_buffer + e * row_size_quad;
    if (key_width == 4) for (uint32_t k = 0; k < key_count; ++k) reinterpret_cast<int32_t*>(row)[k] = HDK_B200_EMPTY_KEY_32;
    else for (uint32_t k = 0; k < key_count; ++k) row[k] = HDK_B200_EMPTY_KEY_64;
    for (uint32_t j = values_off_quad; j < row_size_quad; ++j) row[j] = init_vals[j - values_off_quad];
  }
}

__global__ void ref_init_columnar_kernel(int8_t* buf, const int64_t* init_vals, uint32_t entry_count, uint32_t key_count,
                                         uint32_t agg_col_count, const int8_t* col_sizes, bool need_padding, bool keyless,
                                         int8_t key_size) {
  const uint64_t start = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  uint64_t off = 0;
  auto fill = [&](int w, int64_t v) {
    for (uint64_t i = start; i < entry_count; i += step) {
      if (w == 1) buf[off + i] = int8_t(v);
      else if (w == 2) reinterpret_cast<int16_t*>(buf + off)[i] = int16_t(v);
      else if (w == 4) reinterpret_cast<int32_t*>(buf + off)[i] = int32_t(v);
      else reinterpret_cast<int64_t*>(buf + off)[i] = v;
    }
    off += uint64_t(w) * entry_count;
  };
  if (!keyless)
    for (uint32_t k = 0; k < key_count; ++k) {
      fill(key_size, key_size == 1 ? 127 : key_size == 2 ? 32767 : key_size == 4 ? int64_t(HDK_B200_EMPTY_KEY_32) : HDK_B200_EMPTY_KEY_64);
      off = (off + 7) & ~uint64_t(7);
    }
  uint32_t init_idx = 0;
  for (uint32_t c = 0; c < agg_col_count; ++c) {
    if (need_padding) off = (off + 7) & ~uint64_t(7);
    if (col_sizes[c] == 0) continue;
    fill(col_sizes[c], init_vals[init_idx++]);
  }
}

// ---------------------------------------------------------------------------------------------
// finalize: neutral work table → reference encoding
// ---------------------------------------------------------------------------------------------
struct FinalizeArgs {
  DLayout layout;
  DKey keys[HDK_B200_MAX_KEYS];
  int32_t n_keys;
  uint32_t entry_count;
  const int64_t* work;
  uint64_t acc_stride, entry_stride;   // cell (acc, e) = work[acc * acc_stride + e * entry_stride]
  int32_t baseline;                // baseline hash: keys were written by the scan's claim; empty entries keep the caller's init
  // multi-GPU exchange: `work` points at n_peers slots of `slot_cells` cells (one partial table per rank)
  uint32_t n_peers;
  uint64_t slot_cells;
  uint64_t epoch;
  const unsigned long long* flags; // [n_peers] flags of this epoch's parity
  int32_t* error_codes;
  AccKinds kinds;
  const int* run_if;               // when set: nothing happens unless *run_if != 0
  int8_t* buf;                     // direct pointer, or
  int64_t* const* buf_indirect;    // GROUPBY_BUF-style device array of pointers ([0] is used)
};

__device__ __forceinline__ void store_slot(int8_t* p, int bytes, int padded, int64_t v) {
  if (padded == 8) {
    // a 4-byte aggregate (float) inside an 8-byte slot keeps the upper half of the init pattern (0)
    *reinterpret_cast<int64_t*>(p) = bytes == 8 ? v : int64_t(uint32_t(v));
  } else {
    *reinterpret_cast<int32_t*>(p) = int32_t(v);
  }
}

// key value of bin `comp` (the NULL bin holds max + bucket = the translated NULL, as the reference writes it)
__device__ __forceinline__ int64_t key_of(const DKey& ky, int64_t comp) { return ky.min_val + comp * (ky.bucket ? ky.bucket : 1); }

// merge of one accumulator cell over the ranks' partial tables (kMerged) or the plain work-table cell
template <bool kMerged>
__device__ __forceinline__ int64_t work_cell(const FinalizeArgs& a, int acc, uint64_t e) {
  const uint64_t i = uint64_t(acc) * a.acc_stride + e * a.entry_stride;
  if (!kMerged) return a.work[i];
  const uint8_t kind = a.kinds.kind[acc];
  int64_t x = __ldcg(a.work + i);
  for (uint32_t r = 1; r < a.n_peers; ++r) {
    const int64_t y = __ldcg(a.work + uint64_t(r) * a.slot_cells + i);
    if (kind == ACC_SUM_F) x = __double_as_longlong(__longlong_as_double(x) + __longlong_as_double(y));
    else if (kind == ACC_MIN_I || kind == ACC_MIN_F) x = min(x, y);
    else if (kind == ACC_MAX_I || kind == ACC_MAX_F) x = max(x, y);
    else x += y;   // counters, int64 SUM
  }
  return x;
}

template <bool kMerged>
__global__ void finalize_kernel(const __grid_constant__ FinalizeArgs a) {
  const DLayout& L = a.layout;
  const uint64_t E = a.entry_count;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  if (a.run_if && *a.run_if == 0) return;
  int8_t* const buf = a.buf ? a.buf : reinterpret_cast<int8_t*>(a.buf_indirect[0]);
  if (kMerged) {
    // wait until every rank has published its table of this epoch (flags are written after a system-wide fence)
    if (threadIdx.x == 0) {
      bool ok = true;
      for (uint32_t r = 0; r < a.n_peers && ok; ++r) {
        uint64_t spins = 0;
        while (*reinterpret_cast<const volatile unsigned long long*>(a.flags + r) < a.epoch) {
          if (++spins > (1ull << 22)) { ok = false; break; }   // a few seconds: a peer died or never launched — report, do not hang
          __nanosleep(100);
        }
      }
      if (!ok) record_error(a.error_codes, HDK_B200_ERR_PEER_TIMEOUT);
      __threadfence_system();
    }
    __syncthreads();
  }
  for (uint64_t e = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; e < E; e += step) {
    const int64_t rows = work_cell<kMerged>(a, 0, e);  // accumulator 0 = rows in the group
    const bool empty = rows == 0;
    if (a.baseline && empty) continue;
    int8_t* row = L.columnar ? nullptr : buf + e * L.row_bytes;
    // group keys: the (NULL-translated) key value, or EMPTY_KEY
    int64_t comp[HDK_B200_MAX_KEYS];
    if (!a.baseline)
      for (int k = 0; k < a.n_keys; ++k) comp[k] = a.n_keys == 1 ? int64_t(e) : int64_t((e / uint64_t(a.keys[k].mult)) % uint64_t(a.keys[k].card));
    if (!L.keyless && !a.baseline) {
      for (int k = 0; k < a.n_keys; ++k) {
        const int64_t kv = empty ? HDK_B200_EMPTY_KEY_64 : key_of(a.keys[k], comp[k]);
        if (L.columnar) reinterpret_cast<int64_t*>(buf + size_t(k) * ((8 * E + 7) & ~uint64_t(7)))[e] = kv;
        else reinterpret_cast<int64_t*>(row)[k] = kv;
      }
    }
    for (int s = 0; s < L.slot_count; ++s) {
      const DSlot& sl = L.slots[s];
      if (!sl.padded || (a.baseline && sl.op == SLOT_KEY)) continue;
      int8_t* p = L.columnar ? buf + sl.col_off + e * sl.padded : row + L.key_bytes + sl.off;
      int64_t v = sl.init_val;
      if (!empty) {
        switch (sl.op) {
          case SLOT_KEY: {
            const DKey& ky = a.keys[sl.key_index];
            const bool is_null = ky.has_nulls && comp[sl.key_index] == ky.card - 1;
            v = is_null ? int_null_of(sl.key_width) : key_of(ky, comp[sl.key_index]);
            break;
          }
          case SLOT_COUNT: v = work_cell<kMerged>(a, sl.acc, e); break;
          default: {
            const bool any = !sl.skip_null || work_cell<kMerged>(a, sl.acc_cnt, e) != 0;
            if (any || sl.is_avg_sum) {
              const int64_t cell = work_cell<kMerged>(a, sl.acc, e);
              if (sl.is_fp) {
                const double d = sl.op == SLOT_SUM ? __longlong_as_double(cell) : f64_order_decode(cell);
                v = sl.bytes == 4 ? int64_t(__float_as_uint(float(d))) : __double_as_longlong(d);
              } else {
                v = cell;
              }
              if (!any) v = sl.is_fp ? (sl.bytes == 4 ? int64_t(__float_as_uint(0.f)) : 0) : 0;  // AVG over all-NULL: sum stays 0
            }
          }
        }
      }
      // Keyless + columnar, reference quirk: get_columnar_group_bin_offset (QE/GroupByRuntime.cpp:233-246, picked by
      // RowFuncBuilder::codegenSingleColumnPerfectHash, QE/RowFuncBuilder.cpp:604-607) is handed the buffer's FIRST column
      // — a slot column when there is no key column — and stores the row's (translated) key there while the cell still
      // reads EMPTY_KEY_64.  Only an 8-byte MIN over a NOT NULL int64 starts at that pattern (INT64_MAX): the key then
      // takes part in the minimum.  Reproduced so that the buffer stays the reference's, byte for byte.
      if (s == 0 && !empty && L.columnar && L.keyless && a.n_keys == 1 && sl.padded == 8 && sl.init_val == HDK_B200_EMPTY_KEY_64 &&
          sl.op == SLOT_MIN && !sl.is_fp)
        v = min(v, key_of(a.keys[0], comp[0]));
      store_slot(p, sl.bytes, sl.padded, v);
    }
    if (!L.columnar && !a.baseline) {
      // zero alignment padding inside the row (keys part with 4-byte keys, tail) for deterministic bytes
      // — nothing to do: perfect-hash rows consist of 8-byte keys and 4/8-byte slots; a trailing 4-byte
      // pad exists only when the slot area has an odd number of 4-byte slots.
      uint32_t used = L.key_bytes;
      for (int s = 0; s < L.slot_count; ++s) used = max(used, L.key_bytes + uint32_t(L.slots[s].off) + L.slots[s].padded);
      if (used < L.row_bytes) *reinterpret_cast<int32_t*>(row + used) = 0;
    }
  }
}

int launch_finalize(const Lowered& lw, const int64_t* work_table, int64_t* groups_buffer, int64_t* const* groups_buffer_indirect,
                    cudaStream_t stream, const int* run_if) {
  FinalizeArgs a{};
  a.run_if = run_if;
  a.buf_indirect = groups_buffer_indirect;
  a.layout = lw.layout;
  for (int k = 0; k < lw.plan.n_keys; ++k) a.keys[k] = lw.plan.keys[k];
  a.n_keys = lw.plan.n_keys;
  a.entry_count = lw.plan.entry_count;
  a.work = work_table;
  a.baseline = lw.plan.hash_type == HDK_B200_BASELINE_HASH;
  a.acc_stride = a.baseline ? 1 : a.entry_count;
  a.entry_stride = a.baseline ? uint64_t(lw.plan.n_acc) : 1;
  a.buf = reinterpret_cast<int8_t*>(groups_buffer);
  finalize_kernel<false><<<grid_for(a.entry_count, 128), 128, 0, stream>>>(a);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int launch_finalize_exchange(const Lowered& lw, const int64_t* slots, const unsigned long long* flags, uint32_t n_peers, uint64_t epoch,
                             int32_t* error_codes, int64_t* const* groups_buffer_indirect, cudaStream_t stream) {
  FinalizeArgs a{};
  a.buf_indirect = groups_buffer_indirect;
  a.layout = lw.layout;
  for (int k = 0; k < lw.plan.n_keys; ++k) a.keys[k] = lw.plan.keys[k];
  a.n_keys = lw.plan.n_keys;
  a.entry_count = lw.plan.entry_count;
  a.work = slots;
  a.baseline = 0;
  a.acc_stride = a.entry_count;
  a.entry_stride = 1;
  a.n_peers = n_peers;
  a.slot_cells = (uint64_t(lw.plan.n_acc) * lw.plan.entry_count + 1) & ~uint64_t(1);
  a.epoch = epoch;
  a.flags = flags;
  a.error_codes = error_codes;
  for (int i = 0; i < lw.plan.n_acc; ++i) a.kinds.kind[i] = lw.plan.accs[i].kind;
  // few CTAs: each one spins on the flags first
  finalize_kernel<true><<<std::min(grid_for(a.entry_count, 128), sm_count()), 128, 0, stream>>>(a);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

}  // namespace hb

extern "C" {

int hdk_b200_init_group_by_buffer(const hdk_b200_qmd* qmd, int64_t* groups_buffer, void* stream) {
  // layout only: a plan is not needed, build a minimal Lowered from the descriptor
  if (!qmd || !groups_buffer) { hb::set_error("null argument"); return HDK_B200_E_INVALID; }
  hb::Lowered lw{};
  hb::DLayout& L = lw.layout;
  L.entry_count = qmd->entry_count;
  L.key_count = qmd->key_count;
  L.key_width = qmd->key_width;
  L.keyless = qmd->keyless;
  L.columnar = qmd->output_columnar;
  L.slot_count = qmd->slot_count;
  if (qmd->slot_count < 0 || qmd->slot_count > HDK_B200_MAX_SLOTS || qmd->key_count < 0 || qmd->key_count > HDK_B200_MAX_KEYS) {
    hb::set_error("bad descriptor");
    return HDK_B200_E_INVALID;
  }
  size_t rw = 0, co = 0;
  const size_t E = qmd->entry_count;
  if (qmd->output_columnar && !qmd->keyless) co = size_t(qmd->key_count) * hb::align8h(8 * E);
  L.key_bytes = uint32_t(qmd->keyless ? 0 : hb::align8h(size_t(qmd->key_count) * qmd->key_width));
  for (int s = 0; s < qmd->slot_count; ++s) {
    const int w = qmd->slot_padded[s];
    if (w != 0 && w != 4 && w != 8) { hb::set_error("slot %d: padded width %d unsupported", s, w); return HDK_B200_E_UNSUPPORTED; }
    if (w == 8) rw = hb::align8h(rw);
    L.slots[s].off = int32_t(rw);
    L.slots[s].col_off = co;
    L.slots[s].padded = uint8_t(w);
    L.slots[s].init_val = qmd->init_vals[s];
    rw += w;
    co += hb::align8h(size_t(w) * E);
  }
  L.row_bytes = uint32_t(hb::align8h(L.key_bytes + rw));
  return hb::init_group_by_buffer(lw, groups_buffer, static_cast<cudaStream_t>(stream));
}

int hdk_b200_init_group_by_buffer_on_device(int64_t* groups_buffer, const int64_t* init_vals,
                                            uint32_t groups_buffer_entry_count, uint32_t key_count, uint32_t key_width,
                                            uint32_t row_size_quad, int keyless, int8_t warp_size, size_t block_size_x,
                                            size_t grid_size_x, void* stream) {
  const int block = block_size_x ? int(block_size_x) : 256;
  const uint64_t n = keyless ? uint64_t(groups_buffer_entry_count) * row_size_quad * uint64_t(warp_size ? warp_size : 1)
                             : groups_buffer_entry_count;
  const int grid = grid_size_x ? int(grid_size_x) : hb::grid_for(n, block);
  hb::ref_init_group_by_buffer_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      groups_buffer, init_vals, groups_buffer_entry_count, key_count, key_width, row_size_quad, keyless != 0,
      warp_size ? warp_size : int8_t(1));
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_init_columnar_group_by_buffer_on_device(int64_t* groups_buffer, const int64_t* init_vals,
                                                     uint32_t groups_buffer_entry_count, uint32_t key_count,
                                                     uint32_t agg_col_count, const int8_t* col_sizes, int need_padding,
                                                     int keyless, int8_t key_size, size_t block_size_x,
                                                     size_t grid_size_x, void* stream) {
  const int block = block_size_x ? int(block_size_x) : 256;
  const int grid = grid_size_x ? int(grid_size_x) : hb::grid_for(groups_buffer_entry_count, block);
  hb::ref_init_columnar_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<int8_t*>(groups_buffer), init_vals, groups_buffer_entry_count, key_count, agg_col_count,
      col_sizes, need_padding != 0, keyless != 0, key_size);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

}  // extern "C"
