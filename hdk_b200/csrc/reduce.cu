// hdk_b200/csrc/reduce.cu — device-side ResultSetReduction and result compaction.
//
//   reduce   ResultSetReduction::reduce (QE/ResultSetReduction.cpp:174-330): perfect hash = entry-wise
//            reduceOneSlot (:1234-1320, AGGREGATE_ONE_* :1026-1170) after copying the key; baseline =
//            re-insert each non-empty entry of `that` (reduceOneEntryBaseline :696-760,
//            get_group_value_reduction :560-690) then the same slot reduction.
//   compact  iteration over non-empty entries (ResultSetStorage::isEmptyEntry,
//            omniscidb/ResultSet/ResultSetStorage.cpp:439-525) with target decoding
//            (ResultSet::makeTargetValue, ResultSetIteration.cpp:1264-1360) and AVG finalisation
//            (pair_to_double, ResultSetBufferAccessors.h:168-195).
#include <algorithm>

#include "baseline.cuh"
#include "common.cuh"

namespace hb {

static int grid_for_r(uint64_t n, int block) {
  return int(std::max<uint64_t>(1, std::min<uint64_t>((n + block - 1) / block, uint64_t(sm_count()) * 16)));
}

__device__ __forceinline__ int64_t rd_slot(const int8_t* p, int w) {
  return w == 8 ? *reinterpret_cast<const int64_t*>(p) : int64_t(*reinterpret_cast<const int32_t*>(p));
}
__device__ __forceinline__ void wr_slot(int8_t* p, int w, int64_t v) {
  if (w == 8) *reinterpret_cast<int64_t*>(p) = v; else *reinterpret_cast<int32_t*>(p) = int32_t(v);
}

__device__ __forceinline__ bool entry_is_empty(const DLayout& L, const int8_t* buf, uint64_t E, uint64_t e) {
  if (L.key_count == 0) return false;   // NonGroupedAggregate: the one entry always counts (ResultSetStorage.cpp:440-443)
  if (L.keyless) {
    const DSlot& s = L.slots[L.target_idx_for_key];
    const int8_t* p = L.columnar ? buf + s.col_off + e * s.padded : buf + e * L.row_bytes + s.off;
    return rd_slot(p, s.padded) == (s.padded == 4 ? int64_t(int32_t(s.init_val)) : s.init_val);
  }
  if (L.columnar) return reinterpret_cast<const int64_t*>(buf)[e] == HDK_B200_EMPTY_KEY_64;
  const int8_t* row = buf + e * L.row_bytes;
  return L.key_width == 4 ? *reinterpret_cast<const int32_t*>(row) == HDK_B200_EMPTY_KEY_32
                          : *reinterpret_cast<const int64_t*>(row) == HDK_B200_EMPTY_KEY_64;
}

// this ⊕= that for one slot (no concurrency on an entry: one thread owns it)
__device__ __forceinline__ void reduce_slot(const DSlot& s, int8_t* a, const int8_t* b) {
  const int w = s.op == SLOT_KEY || s.op == SLOT_COUNT ? s.padded : s.bytes;
  const int64_t init = w == 4 ? int64_t(int32_t(s.init_val)) : s.init_val;
  switch (s.op) {
    case SLOT_KEY: {  // projected column: take the right-hand value unless it is the init value
      const int64_t rhs = rd_slot(b, s.padded);
      if (rhs != (s.padded == 4 ? int64_t(int32_t(s.init_val)) : s.init_val)) wr_slot(a, s.padded, rhs);
      return;
    }
    case SLOT_COUNT:
      if (w == 4) *reinterpret_cast<uint32_t*>(a) += *reinterpret_cast<const uint32_t*>(b);
      else *reinterpret_cast<uint64_t*>(a) += *reinterpret_cast<const uint64_t*>(b);
      return;
    default: {
      const bool skip = s.skip_null != 0;
      const int64_t rhs = rd_slot(b, w), lhs = rd_slot(a, w);
      if (skip && rhs == init) return;              // val == skip_val
      if (skip && lhs == init) { wr_slot(a, w, rhs); return; }  // old == skip_val → take val
      int64_t r;
      if (s.is_fp) {
        if (w == 4) {
          const float x = __int_as_float(int32_t(lhs)), y = __int_as_float(int32_t(rhs));
          const float z = s.op == SLOT_SUM ? x + y : s.op == SLOT_MIN ? (y < x ? y : x) : (y > x ? y : x);
          r = int64_t(__float_as_int(z));
        } else {
          const double x = __longlong_as_double(lhs), y = __longlong_as_double(rhs);
          const double z = s.op == SLOT_SUM ? x + y : s.op == SLOT_MIN ? (y < x ? y : x) : (y > x ? y : x);
          r = __double_as_longlong(z);
        }
      } else {
        r = s.op == SLOT_SUM ? int64_t(uint64_t(lhs) + uint64_t(rhs)) : s.op == SLOT_MIN ? min(lhs, rhs) : max(lhs, rhs);
      }
      wr_slot(a, w, r);
    }
  }
}

struct ReduceArgs {
  DLayout layout;       // of `this`
  uint32_t that_entry_count;
  int8_t* this_buf;
  const int8_t* that_buf;
  int32_t* error_codes;
};

__device__ __forceinline__ int8_t* slot_ptr(const DLayout& L, int8_t* buf, uint64_t e, const DSlot& s) {
  return L.columnar ? buf + s.col_off + e * s.padded : buf + e * L.row_bytes + L.key_bytes + s.off;
}

__global__ void reduce_perfect_kernel(const __grid_constant__ ReduceArgs a) {
  const DLayout& L = a.layout;
  const uint64_t E = L.entry_count;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t e = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; e < E; e += step) {
    if (entry_is_empty(L, a.that_buf, E, e)) continue;
    if (!L.keyless) {
      for (int k = 0; k < L.key_count; ++k) {
        if (L.columnar) {
          const uint64_t off = uint64_t(k) * ((8 * E + 7) & ~uint64_t(7)) + 8 * e;
          *reinterpret_cast<int64_t*>(a.this_buf + off) = *reinterpret_cast<const int64_t*>(a.that_buf + off);
        } else {
          reinterpret_cast<int64_t*>(a.this_buf + e * L.row_bytes)[k] = reinterpret_cast<const int64_t*>(a.that_buf + e * L.row_bytes)[k];
        }
      }
    }
    for (int s = 0; s < L.slot_count; ++s) {
      const DSlot& sl = L.slots[s];
      if (!sl.padded) continue;
      reduce_slot(sl, slot_ptr(L, a.this_buf, e, sl), slot_ptr(L, const_cast<int8_t*>(a.that_buf), e, sl));
    }
  }
}

__global__ void reduce_baseline_kernel(const __grid_constant__ ReduceArgs a) {
  const DLayout& L = a.layout;
  const uint64_t TE = a.that_entry_count;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t e = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; e < TE; e += step) {
    // emptiness / keys of `that`
    int64_t keys[HDK_B200_MAX_KEYS];
    if (L.columnar) {
      const int64_t* kb = reinterpret_cast<const int64_t*>(a.that_buf);
      if (kb[e] == HDK_B200_EMPTY_KEY_64) continue;
      for (int k = 0; k < L.key_count; ++k) keys[k] = kb[uint64_t(k) * TE + e];  // 8·TE is 8-aligned
    } else {
      const int8_t* row = a.that_buf + e * L.row_bytes;
      if (L.key_width == 4) {
        if (*reinterpret_cast<const int32_t*>(row) == HDK_B200_EMPTY_KEY_32) continue;
        for (int k = 0; k < L.key_count; ++k) keys[k] = reinterpret_cast<const int32_t*>(row)[k];
      } else {
        if (*reinterpret_cast<const int64_t*>(row) == HDK_B200_EMPTY_KEY_64) continue;
        for (int k = 0; k < L.key_count; ++k) keys[k] = reinterpret_cast<const int64_t*>(row)[k];
      }
    }
    const uint32_t h0 = key_hash_dev(keys, L.key_count, L.key_width) % L.entry_count;
    const int64_t dst = L.columnar ? baseline_claim_columnar(reinterpret_cast<int64_t*>(a.this_buf), L.entry_count, keys, L.key_count, h0)
                        : L.key_width == 4 ? baseline_claim_rowwise<int32_t>(a.this_buf, L.row_bytes, L.entry_count, keys, L.key_count, h0)
                                           : baseline_claim_rowwise<int64_t>(a.this_buf, L.row_bytes, L.entry_count, keys, L.key_count, h0);
    if (dst < 0) { record_error(a.error_codes, dst == kClaimTimedOut ? HDK_B200_ERR_CLAIM_TIMEOUT : -HDK_B200_ERR_OUT_OF_SLOTS); continue; }
    // keys are unique inside `that`, so exactly one thread touches the destination entry's slots;
    // a freshly claimed entry still holds the init values, so reducing into it equals a copy
    for (int s = 0; s < L.slot_count; ++s) {
      const DSlot& sl = L.slots[s];
      if (!sl.padded) continue;
      const int8_t* src;
      if (L.columnar) {
        // slot column offsets of `that`: recompute with its entry count
        uint64_t off = uint64_t(L.key_count) * 8 * TE;
        for (int j = 0; j < s; ++j) off += (uint64_t(L.slots[j].padded) * TE + 7) & ~uint64_t(7);
        src = a.that_buf + off + e * sl.padded;
      } else {
        src = a.that_buf + e * L.row_bytes + L.key_bytes + sl.off;
      }
      reduce_slot(sl, slot_ptr(L, a.this_buf, uint64_t(dst), sl), src);
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct CompactArgs {
  DLayout layout;
  int32_t n_targets;
  // per target: first slot (or -1 → read key column key_index), agg kind, typing
  int16_t slot[HDK_B200_MAX_TARGETS];
  uint8_t agg[HDK_B200_MAX_TARGETS];
  uint8_t chosen_is_fp[HDK_B200_MAX_TARGETS], chosen_width[HDK_B200_MAX_TARGETS], type_width[HDK_B200_MAX_TARGETS];
  uint8_t float_arg[HDK_B200_MAX_TARGETS], sum_is_fp[HDK_B200_MAX_TARGETS];
  int8_t key_index[HDK_B200_MAX_TARGETS];
  const int8_t* buf;
  int64_t* const* out_cols;
  unsigned long long* row_count;
};

// one decoded 8-byte cell of target t of entry e (ResultSet::getTargetValueFromBufferRowwise / Colwise + pair_to_double)
__device__ __forceinline__ int64_t compact_cell(const CompactArgs& a, const DLayout& L, uint64_t E, uint64_t e, int t) {
  int8_t* buf = const_cast<int8_t*>(a.buf);
  if (a.agg[t] == HDK_B200_AGG_AVG) {
    const DSlot& s0 = L.slots[a.slot[t]];
    const DSlot& s1 = L.slots[a.slot[t] + 1];
    const int64_t sum = rd_slot(slot_ptr(L, buf, e, s0), a.float_arg[t] ? 4 : s0.padded);
    const int64_t cnt = rd_slot(slot_ptr(L, buf, e, s1), s1.padded);
    if (cnt == 0) return __double_as_longlong(2.2250738585072014e-308);  // NULL_DOUBLE
    const double dividend = a.float_arg[t] ? double(__int_as_float(int32_t(sum))) : a.sum_is_fp[t] ? __longlong_as_double(sum) : double(sum);
    return __double_as_longlong(dividend / double(cnt));
  }
  int64_t raw;
  if (a.slot[t] >= 0) {
    const DSlot& s = L.slots[a.slot[t]];
    const bool f4 = a.float_arg[t] != 0;
    raw = rd_slot(slot_ptr(L, buf, e, s), f4 ? 4 : s.padded);
    if (f4) raw = __double_as_longlong(double(__int_as_float(int32_t(raw))));
  } else {
    const int k = a.key_index[t];
    if (L.columnar) raw = reinterpret_cast<const int64_t*>(a.buf)[uint64_t(k) * E + e];
    else if (L.key_width == 4) raw = reinterpret_cast<const int32_t*>(a.buf + e * L.row_bytes)[k];
    else raw = reinterpret_cast<const int64_t*>(a.buf + e * L.row_bytes)[k];
  }
  if (!a.chosen_is_fp[t] && int_null_of(a.chosen_width[t]) == resize_int(raw, a.chosen_width[t])) raw = int_null_of(a.type_width[t]);
  return raw;
}

constexpr int kCompactThreads = 256;
constexpr int kCompactItems = 4;     // entries per thread per iteration: one reservation per 1024 entries

__global__ void __launch_bounds__(kCompactThreads) compact_kernel(const __grid_constant__ CompactArgs a) {
  const DLayout& L = a.layout;
  const uint64_t E = L.entry_count;
  const uint64_t step = uint64_t(gridDim.x) * (kCompactThreads * kCompactItems);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ unsigned int warp_cnt[kCompactThreads / 32];
  __shared__ unsigned long long cta_base;
  for (uint64_t base = blockIdx.x * uint64_t(kCompactThreads * kCompactItems); base < E; base += step) {   // whole CTAs iterate together
    // One reservation per CTA iteration: a single counter retires about one atomic per ns, so per-warp reservations bound
    // the kernel at 2e8 entries.  Row = CTA base + non-empty entries of earlier warps + of earlier items / lanes of this warp.
    bool keep[kCompactItems];
    unsigned m[kCompactItems];
    unsigned mine = 0;
#pragma unroll
    for (int j = 0; j < kCompactItems; ++j) {
      const uint64_t e = base + uint64_t(j) * kCompactThreads + threadIdx.x;
      keep[j] = e < E && !entry_is_empty(L, a.buf, E, e);
      m[j] = __ballot_sync(0xffffffffu, keep[j]);
      mine += __popc(m[j]);
    }
    if (lane == 0) warp_cnt[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned total = 0;
      for (int w = 0; w < kCompactThreads / 32; ++w) { const unsigned c = warp_cnt[w]; warp_cnt[w] = total; total += c; }
      cta_base = total ? atomicAdd(a.row_count, (unsigned long long)total) : 0ull;
    }
    __syncthreads();
    unsigned long long wbase = cta_base + warp_cnt[warp];
    __syncthreads();     // warp_cnt / cta_base are rewritten by the next iteration
#pragma unroll
    for (int j = 0; j < kCompactItems; ++j) {
      const unsigned long long r = wbase + __popc(m[j] & ((1u << lane) - 1u));
      wbase += __popc(m[j]);
      if (!keep[j]) continue;
      const uint64_t e = base + uint64_t(j) * kCompactThreads + threadIdx.x;
      for (int t = 0; t < a.n_targets; ++t) a.out_cols[t][r] = compact_cell(a, L, E, e, t);
    }
  }
}

}  // namespace hb

using namespace hb;

extern "C" {

int hdk_b200_reduce(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, int64_t* this_buffer, const int64_t* that_buffer,
                    uint32_t that_entry_count, int32_t* error_codes, void* stream) {
  Lowered lw;
  if (int rc = lower_plan(plan, qmd, &lw)) return rc;
  if (!this_buffer || !that_buffer) { set_error("null buffer"); return HDK_B200_E_INVALID; }
  ReduceArgs a{};
  a.layout = lw.layout;
  a.that_entry_count = that_entry_count;
  a.this_buf = reinterpret_cast<int8_t*>(this_buffer);
  a.that_buf = reinterpret_cast<const int8_t*>(that_buffer);
  a.error_codes = error_codes;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (qmd->hash_type == HDK_B200_PERFECT_HASH) {
    if (that_entry_count != qmd->entry_count) { set_error("perfect-hash reduce needs equal entry counts"); return HDK_B200_E_INVALID; }
    reduce_perfect_kernel<<<grid_for_r(qmd->entry_count, 128), 128, 0, st>>>(a);
  } else {
    if (that_entry_count > qmd->entry_count) { set_error("baseline reduce: that_entry_count > entry_count"); return HDK_B200_E_INVALID; }
    if (!error_codes) { set_error("baseline reduce needs an error code buffer"); return HDK_B200_E_INVALID; }
    reduce_baseline_kernel<<<grid_for_r(that_entry_count, 128), 128, 0, st>>>(a);
  }
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_compact_result(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const int64_t* groups_buffer,
                            int64_t* const* out_cols, uint64_t* row_count, void* stream) {
  Lowered lw;
  if (int rc = lower_plan(plan, qmd, &lw)) return rc;
  if (!groups_buffer || !out_cols || !row_count) { set_error("null argument"); return HDK_B200_E_INVALID; }
  CompactArgs a{};
  a.layout = lw.layout;
  a.n_targets = plan->n_targets;
  for (int t = 0; t < plan->n_targets; ++t) {
    const hdk_b200_target& tg = plan->targets[t];
    const bool has_slot = tg.slot >= 0 && qmd->slot_padded[tg.slot] != 0;
    a.slot[t] = int16_t(has_slot ? tg.slot : -1);
    a.agg[t] = uint8_t(tg.agg);
    a.key_index[t] = int8_t(tg.key_index);
    hdk_b200_type chosen = tg.type;
    const bool minmax = tg.agg == HDK_B200_AGG_MIN || tg.agg == HDK_B200_AGG_MAX;
    if (tg.agg != HDK_B200_AGG_NONE && tg.arg >= 0 && minmax) chosen = tg.arg_type;
    a.chosen_is_fp[t] = uint8_t(chosen.kind == HDK_B200_FP && tg.agg != HDK_B200_AGG_COUNT);
    a.chosen_width[t] = uint8_t(chosen.width);
    a.type_width[t] = uint8_t(tg.type.width);
    const bool float_arg = (tg.agg == HDK_B200_AGG_AVG || tg.agg == HDK_B200_AGG_SUM || minmax) && tg.arg >= 0 &&
                           tg.arg_type.kind == HDK_B200_FP && tg.arg_type.width == 4;
    a.float_arg[t] = uint8_t(float_arg);
    a.sum_is_fp[t] = uint8_t(tg.type.kind == HDK_B200_FP);
  }
  a.buf = reinterpret_cast<const int8_t*>(groups_buffer);
  a.out_cols = out_cols;
  a.row_count = reinterpret_cast<unsigned long long*>(row_count);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CUDA(cudaMemsetAsync(row_count, 0, sizeof(uint64_t), st));
  compact_kernel<<<int(std::min<uint64_t>(148 * 8, (uint64_t(qmd->entry_count) + kCompactThreads * kCompactItems - 1) / (kCompactThreads * kCompactItems))), kCompactThreads, 0, st>>>(a);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

// ---- Arrow buffers of one result column (ArrowResultSetConverter::convertToArrow, omniscidb/ResultSet/
//      ArrowResultSetConverter.cpp: value buffer of the Arrow type + validity bitmap, LSB first) ----------------------------
namespace hb {
struct ArrowColArgs {
  const int64_t* cells;
  uint64_t n;
  int32_t cells_fp, out_fp, out_width, nullable;
  int64_t null_int;
  double null_fp;
  int8_t* values;
  uint32_t* validity;
  unsigned long long* null_count;
};
__global__ void __launch_bounds__(256) arrow_column_kernel(const __grid_constant__ ArrowColArgs a) {
  const uint64_t n32 = (a.n + 31) / 32 * 32;    // whole validity words: every lane of a warp votes
  unsigned long long nulls = 0;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n32; i += uint64_t(gridDim.x) * blockDim.x) {
    bool valid = false;
    if (i < a.n) {
      const int64_t c = a.cells[i];
      valid = !a.nullable || (a.cells_fp ? __longlong_as_double(c) != a.null_fp : c != a.null_int);
      if (a.out_fp) {
        const double d = valid ? __longlong_as_double(c) : 0.0;
        if (a.out_width == 4) reinterpret_cast<float*>(a.values)[i] = float(d);
        else reinterpret_cast<double*>(a.values)[i] = d;
      } else {
        const int64_t v = valid ? c : 0;      // Arrow leaves NULL slots unspecified: zero keeps the buffers deterministic
        if (a.out_width == 8) reinterpret_cast<int64_t*>(a.values)[i] = v;
        else if (a.out_width == 4) reinterpret_cast<int32_t*>(a.values)[i] = int32_t(v);
        else if (a.out_width == 2) reinterpret_cast<int16_t*>(a.values)[i] = int16_t(v);
        else a.values[i] = int8_t(v);
      }
      nulls += !valid;
    }
    const unsigned word = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && a.validity) a.validity[i / 32] = word;
  }
  for (int d = 16; d; d >>= 1) nulls += __shfl_xor_sync(0xffffffffu, nulls, d);
  if ((threadIdx.x & 31) == 0 && nulls && a.null_count) atomicAdd(a.null_count, nulls);
}
}  // namespace hb

int hdk_b200_arrow_column_on_device(const int64_t* cells, uint64_t n_rows, int cells_are_fp, int out_is_fp, int out_width, int nullable,
                                    int64_t null_int, double null_fp, int8_t* values, uint32_t* validity, uint64_t* null_count,
                                    void* stream) {
  if ((!cells && n_rows) || !values || (out_width != 1 && out_width != 2 && out_width != 4 && out_width != 8) ||
      (out_is_fp && out_width < 4) || (out_is_fp && !cells_are_fp)) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (null_count) HB_CUDA(cudaMemsetAsync(null_count, 0, sizeof(uint64_t), st));
  if (!n_rows) return HDK_B200_OK;
  hb::ArrowColArgs a{cells, n_rows, cells_are_fp, out_is_fp, out_width, nullable, null_int, null_fp, values, validity,
                     reinterpret_cast<unsigned long long*>(null_count)};
  const int grid = int(std::min<uint64_t>(uint64_t(hb::sm_count()) * 8, (n_rows + 255) / 256));
  hb::arrow_column_kernel<<<grid, 256, 0, st>>>(a);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

}  // extern "C"
