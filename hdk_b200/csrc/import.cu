// hdk_b200/csrc/import.cu — storage side of the path on the device: Arrow validity bitmap → in-band NULL
// sentinels, and the chunk statistics (min / max / has_nulls) in the same pass.
//
// Restates what ArrowStorage does on the host while importing (omniscidb/ArrowStorage/ArrowStorageUtils.cpp:100-170
// copyArrayDataReplacingNulls; ArrowStorage.cpp:1000-1040 ChunkStats), for data that was copied to the device
// as raw Arrow buffers.  HBM-bound: one read + (for NULL slots) one write of the column.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace hb {

template <typename T, bool kFp>
__global__ void materialize_kernel(T* __restrict__ vals, const uint8_t* __restrict__ validity, int64_t bit_offset, int64_t n,
                                   hdk_b200_chunk_stats* __restrict__ st) {
  using S = typename std::conditional<sizeof(T) == 1, int8_t, typename std::conditional<sizeof(T) == 2, int16_t,
            typename std::conditional<sizeof(T) == 4, int32_t, int64_t>::type>::type>::type;
  const T sentinel = kFp ? (sizeof(T) == 4 ? T(1.17549435e-38f) : T(2.2250738585072014e-308)) : T(int_null_of(sizeof(T)));
  int64_t mn = INT64_MAX, mx = INT64_MIN;
  unsigned long long nulls = 0;
  const int64_t step = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += step) {
    T v = vals[i];
    if (validity) {
      const int64_t b = bit_offset + i;
      if (!((validity[b >> 3] >> (b & 7)) & 1)) {
        v = sentinel;
        vals[i] = v;
      }
    }
    if (v == sentinel) {
      ++nulls;
    } else {
      const int64_t key = kFp ? f64_order_encode(double(v)) : int64_t(S(v));
      mn = min(mn, key);
      mx = max(mx, key);
    }
  }
  for (int d = 16; d; d >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    nulls += __shfl_xor_sync(0xffffffffu, nulls, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (mn != INT64_MAX) {
      atomicMin(reinterpret_cast<long long*>(kFp ? &st->min_f_enc : &st->min_i), (long long)mn);
      atomicMax(reinterpret_cast<long long*>(kFp ? &st->max_f_enc : &st->max_i), (long long)mx);
    }
    if (nulls) atomicAdd(reinterpret_cast<unsigned long long*>(&st->null_count), nulls);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&st->row_count), (unsigned long long)n);
}

__global__ void init_chunk_stats_kernel(hdk_b200_chunk_stats* st) {
  st->min_i = INT64_MAX; st->max_i = INT64_MIN;
  st->min_f_enc = INT64_MAX; st->max_f_enc = INT64_MIN;
  st->null_count = 0; st->row_count = 0;
}

}  // namespace hb

using namespace hb;

extern "C" {

int hdk_b200_init_chunk_stats_on_device(hdk_b200_chunk_stats* stats, void* stream) {
  if (!stats) { set_error("null stats"); return HDK_B200_E_INVALID; }
  init_chunk_stats_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(stats);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_materialize_nulls_on_device(int8_t* values, int elem_width, int is_fp, const uint8_t* validity, int64_t bit_offset,
                                         int64_t num_elems, hdk_b200_chunk_stats* stats, void* stream) {
  if (!values || !stats || num_elems < 0 || bit_offset < 0) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  if (is_fp && elem_width != 4 && elem_width != 8) { set_error("floating point width %d", elem_width); return HDK_B200_E_INVALID; }
  if (num_elems == 0) return HDK_B200_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int block = 256;
  const int grid = int(std::min<int64_t>((num_elems + block - 1) / block, int64_t(sm_count()) * 16));
#define HB_MAT(T, FP) materialize_kernel<T, FP><<<grid, block, 0, st>>>(reinterpret_cast<T*>(values), validity, bit_offset, num_elems, stats)
  if (is_fp) {
    if (elem_width == 4) HB_MAT(float, true); else HB_MAT(double, true);
  } else {
    switch (elem_width) {
      case 1: HB_MAT(uint8_t, false); break;
      case 2: HB_MAT(uint16_t, false); break;
      case 4: HB_MAT(uint32_t, false); break;
      case 8: HB_MAT(uint64_t, false); break;
      default: set_error("element width %d", elem_width); return HDK_B200_E_INVALID;
    }
  }
#undef HB_MAT
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

}  // extern "C"
