// hdk_b200/csrc/join.cu — join hash table build and stand-alone probe kernels.
//
// Replaces (QE = omniscidb/QueryEngine, JHT = QE/JoinHashTable):
//   init_hash_join_buff_on_device                         JHT/Runtime/HashJoinRuntimeGpu.cu:95-106
//   fill_hash_join_buff_on_device[_bucketized]            HashJoinRuntimeGpu.cu:32-93, body JHT/Runtime/HashJoinRuntime.cpp:198-296
//   fill_one_to_many_hash_table_on_device[_bucketized]    HashJoinRuntimeGpu.cu:108-236 (count → scan → positions → row ids)
//   init/fill_baseline_hash_join_buff_on_device_{32,64}   HashJoinRuntimeGpu.cu:238-330, HashJoinRuntime.cpp:298-576
//   fill_one_to_many_baseline_hash_table_on_device_{32,64} HashJoinRuntimeGpu.cu:332-409
//   hash_join_idx / baseline_hash_join_idx_{32,64}        QE/GroupByRuntime.cpp:298-308, JHT/Runtime/JoinHashTableQueryRuntime.cpp:43-98
// The buffer layouts are the reference's (perfect: int32[E]; one-to-many: offsets[E] | counts[E] |
// payload[N]; baseline: E × (key components ‖ payload) of the key width), so PerfectHashTable /
// BaselineHashTable can own the buffers unchanged.
#include <algorithm>
#include <vector>

#include "baseline.cuh"
#include "common.cuh"

namespace hb {

static int grid_for_join(uint64_t n, int block) {
  return int(std::max<uint64_t>(1, std::min<uint64_t>((n + block - 1) / block, uint64_t(sm_count()) * 16)));
}

struct JoinCol {
  const hdk_b200_join_chunk* chunks;  // device
  uint64_t num_chunks, num_elems;
  int elem_sz;
  int column_type;
  int64_t min_val, max_val, null_val, translated_null_val;
  int uses_bw_eq;
};

static JoinCol make_col(const hdk_b200_join_column* jc, const hdk_b200_join_column_type_info* ti) {
  JoinCol c;
  c.chunks = reinterpret_cast<const hdk_b200_join_chunk*>(jc->col_chunks_buff);
  c.num_chunks = jc->num_chunks;
  c.num_elems = jc->num_elems;
  c.elem_sz = int(jc->elem_sz);
  c.column_type = ti->column_type;
  c.min_val = ti->min_val;
  c.max_val = ti->max_val;
  c.null_val = ti->null_val;
  c.translated_null_val = ti->translated_null_val;
  c.uses_bw_eq = ti->uses_bw_eq;
  return c;
}

// JoinColumnIterator / JoinColumnTyped element decode (JHT/Runtime/JoinColumnIterator.h)
__device__ __forceinline__ int64_t join_decode(const int8_t* b, uint64_t i, int w, int column_type, int64_t null_val) {
  int64_t v;
  if (column_type == HDK_B200_UNSIGNED) {
    v = w == 1 ? int64_t(reinterpret_cast<const uint8_t*>(b)[i]) : w == 2 ? int64_t(reinterpret_cast<const uint16_t*>(b)[i])
        : w == 4 ? int64_t(reinterpret_cast<const uint32_t*>(b)[i]) : reinterpret_cast<const int64_t*>(b)[i];
  } else {
    v = w == 1 ? int64_t(b[i]) : w == 2 ? int64_t(reinterpret_cast<const int16_t*>(b)[i])
        : w == 4 ? int64_t(reinterpret_cast<const int32_t*>(b)[i]) : reinterpret_cast<const int64_t*>(b)[i];
    if (column_type == HDK_B200_SMALL_DATE) v = (v == int_null_of(w)) ? null_val : v * 86400;
  }
  return v;
}

// visit every (element, global row index) of a chunked join column, grid-stride within each chunk
template <class F>
__device__ __forceinline__ void for_each_elem(const JoinCol& c, F&& f) {
  const uint64_t start = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t ch = 0; ch < c.num_chunks; ++ch) {
    const hdk_b200_join_chunk chunk = c.chunks[ch];
    for (uint64_t i = start; i < chunk.num_elems; i += step) {
      int64_t elem = join_decode(chunk.col_buff, i, c.elem_sz, c.column_type, c.null_val);
      if (elem == c.null_val) {
        if (c.uses_bw_eq) elem = c.translated_null_val; else continue;   // (the translated NULL has its own entry past max_val)
      } else if (elem < c.min_val || elem > c.max_val) {
        // outside the declared range the element would index outside the table (the reference trusts its metadata and
        // would write out of bounds): ignored here
        continue;
      }
      f(elem, chunk.row_id + i);
    }
  }
}

__global__ void fill_i32_kernel(int32_t* buff, int64_t n, int32_t v) {
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  // 16-byte stores where aligned
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < uint64_t(n); i += step) buff[i] = v;
}

__global__ void fill_one_to_one_kernel(int32_t* buff, int32_t invalid, bool semi, int* err, const __grid_constant__ JoinCol c,
                                       int64_t bucket) {
  for_each_elem(c, [&](int64_t elem, uint64_t index) {
    int32_t* e = buff + (bucket > 1 ? (elem - c.min_val) / bucket : (elem - c.min_val));
    const int32_t old = atomicCAS(e, invalid, int32_t(index));
    if (old != invalid && !semi) atomicCAS(err, 0, -1);  // duplicate key → NeedsOneToManyHash
  });
}

__global__ void count_matches_kernel(int32_t* count_buff, const __grid_constant__ JoinCol c, int64_t bucket) {
  for_each_elem(c, [&](int64_t elem, uint64_t) {
    atomicAdd(count_buff + (bucket > 1 ? (elem - c.min_val) / bucket : (elem - c.min_val)), 1);
  });
}

__global__ void fill_row_ids_kernel(int32_t* buff, int64_t E, const __grid_constant__ JoinCol c, int64_t bucket) {
  int32_t* pos = buff;
  int32_t* cnt = buff + E;
  int32_t* ids = cnt + E;
  for_each_elem(c, [&](int64_t elem, uint64_t index) {
    const int64_t b = bucket > 1 ? (elem - c.min_val) / bucket : (elem - c.min_val);
    const int32_t k = atomicAdd(cnt + b, 1);
    ids[pos[b] + k] = int32_t(index);
  });
}

// ---- exclusive scan of the per-entry counts into positions (-1 where the count is 0) -----------
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__global__ void scan_tile_sums_kernel(const int32_t* cnt, int64_t n, int32_t* tile_sums) {
  __shared__ int32_t warp_sums[kScanBlock / 32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile;
  int32_t s = 0;
  for (int i = threadIdx.x; i < kScanTile; i += kScanBlock) {
    const int64_t idx = base + i;
    if (idx < n) s += cnt[idx];
  }
  for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int w = 0; w < kScanBlock / 32; ++w) t += warp_sums[w];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void scan_tile_offsets_kernel(int32_t* tile_sums, int64_t n_tiles) {
  // single block: exclusive scan of the tile sums in place
  __shared__ int32_t carry_s;
  __shared__ int32_t warp_sums[kScanBlock / 32];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += kScanBlock) {
    const int64_t idx = base + threadIdx.x;
    const int32_t v = idx < n_tiles ? tile_sums[idx] : 0;
    int32_t incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((threadIdx.x & 31) >= d) incl += o;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    int32_t warp_off = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) warp_off += warp_sums[w];
    const int32_t carry = carry_s;
    if (idx < n_tiles) tile_sums[idx] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry_s = carry + warp_off + incl;
    __syncthreads();
  }
}

__global__ void scan_positions_kernel(const int32_t* cnt, int64_t n, const int32_t* tile_offsets, int32_t* pos, int32_t invalid) {
  __shared__ int32_t warp_sums[kScanBlock / 32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  int32_t v[kScanItems];
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? cnt[base + i] : 0;
    s += v[i];
  }
  int32_t incl = s;
  for (int d = 1; d < 32; d <<= 1) {
    const int32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((threadIdx.x & 31) >= d) incl += o;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
  __syncthreads();
  int32_t off = tile_offsets[blockIdx.x] + incl - s;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) off += warp_sums[w];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) pos[base + i] = v[i] ? off : invalid;
    off += v[i];
  }
}

static int build_positions(int32_t* pos, int32_t* cnt, int64_t E, int32_t invalid, cudaStream_t st) {
  const int64_t n_tiles = (E + kScanTile - 1) / kScanTile;
  int32_t* tile_sums = nullptr;
  HB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tile_sums), size_t(n_tiles) * 4, st));
  scan_tile_sums_kernel<<<int(n_tiles), kScanBlock, 0, st>>>(cnt, E, tile_sums);
  HB_LAUNCH_CHECK();
  scan_tile_offsets_kernel<<<1, kScanBlock, 0, st>>>(tile_sums, n_tiles);
  HB_LAUNCH_CHECK();
  scan_positions_kernel<<<int(n_tiles), kScanBlock, 0, st>>>(cnt, E, tile_sums, pos, invalid);
  HB_LAUNCH_CHECK();
  HB_CUDA(cudaFreeAsync(tile_sums, st));
  return HDK_B200_OK;
}

// ---- baseline (composite key) ------------------------------------------------------------------
constexpr int kMaxJoinKeys = 8;  // g_maximum_conditions_to_coalesce (HashJoinRuntime.h:60)
struct JoinCols {
  JoinCol col[kMaxJoinKeys];
  int n;
};

template <typename T>
__global__ void init_baseline_kernel(int8_t* buff, int64_t E, int kc, bool with_val, int32_t invalid) {
  const T empty = sizeof(T) == 4 ? T(HDK_B200_EMPTY_KEY_32) : T(HDK_B200_EMPTY_KEY_64);
  const int n = kc + (with_val ? 1 : 0);
  const uint64_t total = uint64_t(E) * n;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  T* b = reinterpret_cast<T*>(buff);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total; i += step)
    b[i] = (with_val && int(i % n) == kc) ? T(invalid) : empty;
}

// assemble the composite key of flat row `row`; false if a component is NULL (and not bw_eq)
template <typename T>
__device__ __forceinline__ bool composite_key(const JoinCols& cs, uint64_t row, T* key) {
  for (int k = 0; k < cs.n; ++k) {
    const JoinCol& c = cs.col[k];
    // locate the chunk holding `row` (chunks are few: linear walk)
    uint64_t ch = 0;
    while (ch + 1 < c.num_chunks && c.chunks[ch + 1].row_id <= row) ++ch;
    const hdk_b200_join_chunk chunk = c.chunks[ch];
    int64_t elem = join_decode(chunk.col_buff, row - chunk.row_id, c.elem_sz, c.column_type, c.null_val);
    if (elem == c.null_val) {
      if (c.uses_bw_eq) elem = c.translated_null_val; else return false;
    }
    key[k] = T(elem);
  }
  return true;
}

template <typename T>
__global__ void fill_baseline_kernel(int8_t* hash_buff, int64_t E, int32_t invalid, bool semi, bool with_val, int* err,
                                     const __grid_constant__ JoinCols cs) {
  using U = typename std::conditional<sizeof(T) == 4, unsigned int, unsigned long long>::type;
  const size_t entry_sz = size_t(cs.n + (with_val ? 1 : 0)) * sizeof(T);
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t row = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; row < cs.col[0].num_elems; row += step) {
    T key[kMaxJoinKeys];
    if (!composite_key<T>(cs, row, key)) continue;
    T* slot = baseline_slot<T>(hash_buff, E, key, cs.n, entry_sz, true);
    if (!slot) { atomicCAS(err, 0, -2); continue; }  // table full
    if (!with_val) continue;
    const T old = T(atomicCAS(reinterpret_cast<U*>(slot), U(T(invalid)), U(T(row))));
    if (old != T(invalid) && !semi) atomicCAS(err, 0, -1);
  }
}

template <typename T>
__global__ void count_matches_baseline_kernel(int32_t* count_buff, const int8_t* dict, int64_t E, const __grid_constant__ JoinCols cs) {
  const size_t entry_sz = size_t(cs.n) * sizeof(T);
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t row = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; row < cs.col[0].num_elems; row += step) {
    T key[kMaxJoinKeys];
    if (!composite_key<T>(cs, row, key)) continue;
    const T* slot = baseline_slot<T>(const_cast<int8_t*>(dict), E, key, cs.n, entry_sz, false);
    if (!slot) continue;
    const int64_t entry = (reinterpret_cast<const int8_t*>(slot) - dict) / int64_t(entry_sz) - 1;  // slot = row + kc → next entry's start
    atomicAdd(count_buff + entry, 1);
  }
}

template <typename T>
__global__ void fill_row_ids_baseline_kernel(int32_t* buff, const int8_t* dict, int64_t E, const __grid_constant__ JoinCols cs) {
  int32_t* pos = buff;
  int32_t* cnt = buff + E;
  int32_t* ids = cnt + E;
  const size_t entry_sz = size_t(cs.n) * sizeof(T);
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t row = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; row < cs.col[0].num_elems; row += step) {
    T key[kMaxJoinKeys];
    if (!composite_key<T>(cs, row, key)) continue;
    const T* slot = baseline_slot<T>(const_cast<int8_t*>(dict), E, key, cs.n, entry_sz, false);
    if (!slot) continue;
    const int64_t entry = (reinterpret_cast<const int8_t*>(slot) - dict) / int64_t(entry_sz) - 1;
    const int32_t k = atomicAdd(cnt + entry, 1);
    ids[pos[entry] + k] = int32_t(row);
  }
}

__global__ void probe_perfect_kernel(const int32_t* buff, const int64_t* keys, int64_t n, int64_t min_key, int64_t max_key, int64_t* out) {
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < uint64_t(n); i += step) {
    const int64_t k = keys[i];
    out[i] = (k >= min_key && k <= max_key) ? int64_t(buff[k - min_key]) : -1;
  }
}

template <typename T>
__global__ void probe_baseline_kernel(const int8_t* hash_buff, const int8_t* keys, int64_t n, int kc, int64_t E, bool with_val, int64_t* out) {
  const size_t entry_sz = size_t(kc + (with_val ? 1 : 0)) * sizeof(T);
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < uint64_t(n); i += step) {
    const T* key = reinterpret_cast<const T*>(keys) + i * kc;
    T local[kMaxJoinKeys];
    for (int k = 0; k < kc; ++k) local[k] = key[k];
    const T* slot = baseline_slot<T>(const_cast<int8_t*>(hash_buff), E, local, kc, entry_sz, false);
    if (!slot) out[i] = -1;
    else if (with_val) out[i] = int64_t(*slot);
    else out[i] = (reinterpret_cast<const int8_t*>(slot) - hash_buff) / int64_t(entry_sz) - 1;  // entry index of the key dictionary
  }
}

static int make_cols(const hdk_b200_join_column* jcs, const hdk_b200_join_column_type_info* tis, size_t kc, JoinCols* out) {
  if (kc < 1 || kc > size_t(kMaxJoinKeys)) { set_error("key_component_count out of range"); return HDK_B200_E_INVALID; }
  out->n = int(kc);
  for (size_t k = 0; k < kc; ++k) {
    out->col[k] = make_col(&jcs[k], &tis[k]);
    if (jcs[k].num_elems != jcs[0].num_elems) { set_error("join key columns differ in length"); return HDK_B200_E_INVALID; }
  }
  return HDK_B200_OK;
}

// slot-ordered copy of one inner column + presence bitmap (one warp per 32 slots → one bitmap word)
template <typename T>
__global__ void gather_payload_kernel(const int32_t* __restrict__ table, int64_t entries, const T* __restrict__ col,
                                      T* __restrict__ out, uint32_t* __restrict__ bitmap) {
  const int64_t words = (entries + 31) / 32;
  const int lane = threadIdx.x & 31;
  for (int64_t w = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5; w < words; w += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    const int64_t slot = w * 32 + lane;
    const int32_t rid = slot < entries ? table[slot] : -1;
    if (slot < entries) out[slot] = rid >= 0 ? col[rid] : T(0);
    const uint32_t bits = __ballot_sync(0xffffffffu, rid >= 0);
    if (bitmap && lane == 0) bitmap[w] = bits;
  }
}

}  // namespace hb

using namespace hb;

extern "C" {

int hdk_b200_gather_join_payload_on_device(const int32_t* hash_table, int64_t entry_count, const int8_t* inner_col, int elem_width,
                                           int8_t* out_by_slot, uint32_t* present_bitmap, void* stream) {
  if (!hash_table || !inner_col || !out_by_slot || entry_count < 0) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for_join(uint64_t(entry_count), 256);
  switch (elem_width) {
    case 1: gather_payload_kernel<uint8_t><<<grid, 256, 0, st>>>(hash_table, entry_count, reinterpret_cast<const uint8_t*>(inner_col), reinterpret_cast<uint8_t*>(out_by_slot), present_bitmap); break;
    case 2: gather_payload_kernel<uint16_t><<<grid, 256, 0, st>>>(hash_table, entry_count, reinterpret_cast<const uint16_t*>(inner_col), reinterpret_cast<uint16_t*>(out_by_slot), present_bitmap); break;
    case 4: gather_payload_kernel<uint32_t><<<grid, 256, 0, st>>>(hash_table, entry_count, reinterpret_cast<const uint32_t*>(inner_col), reinterpret_cast<uint32_t*>(out_by_slot), present_bitmap); break;
    case 8: gather_payload_kernel<uint64_t><<<grid, 256, 0, st>>>(hash_table, entry_count, reinterpret_cast<const uint64_t*>(inner_col), reinterpret_cast<uint64_t*>(out_by_slot), present_bitmap); break;
    default: set_error("element width %d", elem_width); return HDK_B200_E_INVALID;
  }
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_init_hash_join_buff_on_device(int32_t* buff, int64_t entry_count, int32_t invalid_slot_val, void* stream) {
  if (!buff || entry_count < 0) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  fill_i32_kernel<<<grid_for_join(uint64_t(entry_count), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(buff, entry_count, invalid_slot_val);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_fill_hash_join_buff_on_device(int32_t* buff, int32_t invalid_slot_val, int for_semi_join, int* dev_err_buff,
                                           const hdk_b200_join_column* join_column,
                                           const hdk_b200_join_column_type_info* type_info, int64_t bucket_normalization,
                                           void* stream) {
  if (!buff || !dev_err_buff || !join_column || !type_info) { set_error("null argument"); return HDK_B200_E_INVALID; }
  const JoinCol c = make_col(join_column, type_info);
  fill_one_to_one_kernel<<<grid_for_join(c.num_elems, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      buff, invalid_slot_val, for_semi_join != 0, dev_err_buff, c, bucket_normalization);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_fill_one_to_many_hash_table_on_device(int32_t* buff, int64_t hash_entry_count, int32_t invalid_slot_val,
                                                   const hdk_b200_join_column* join_column,
                                                   const hdk_b200_join_column_type_info* type_info,
                                                   int64_t bucket_normalization, void* stream) {
  if (!buff || !join_column || !type_info || hash_entry_count <= 0) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const JoinCol c = make_col(join_column, type_info);
  int32_t* pos = buff;
  int32_t* cnt = buff + hash_entry_count;
  HB_CUDA(cudaMemsetAsync(cnt, 0, size_t(hash_entry_count) * 4, st));
  count_matches_kernel<<<grid_for_join(c.num_elems, 256), 256, 0, st>>>(cnt, c, bucket_normalization);
  HB_LAUNCH_CHECK();
  if (int rc = build_positions(pos, cnt, hash_entry_count, invalid_slot_val, st)) return rc;
  HB_CUDA(cudaMemsetAsync(cnt, 0, size_t(hash_entry_count) * 4, st));
  fill_row_ids_kernel<<<grid_for_join(c.num_elems, 256), 256, 0, st>>>(buff, hash_entry_count, c, bucket_normalization);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_init_baseline_hash_join_buff_on_device(int8_t* hash_join_buff, int64_t entry_count, size_t key_component_count,
                                                    int with_val_slot, int32_t invalid_slot_val, int key_width, void* stream) {
  if (!hash_join_buff || (key_width != 4 && key_width != 8)) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t total = uint64_t(entry_count) * (key_component_count + (with_val_slot ? 1 : 0));
  if (key_width == 4)
    init_baseline_kernel<int32_t><<<grid_for_join(total, 256), 256, 0, st>>>(hash_join_buff, entry_count, int(key_component_count), with_val_slot != 0, invalid_slot_val);
  else
    init_baseline_kernel<int64_t><<<grid_for_join(total, 256), 256, 0, st>>>(hash_join_buff, entry_count, int(key_component_count), with_val_slot != 0, invalid_slot_val);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_fill_baseline_hash_join_buff_on_device(int8_t* hash_buff, int64_t entry_count, int32_t invalid_slot_val,
                                                    int for_semi_join, size_t key_component_count, int with_val_slot,
                                                    int* dev_err_buff, const hdk_b200_join_column* join_columns,
                                                    const hdk_b200_join_column_type_info* type_infos, int key_width,
                                                    void* stream) {
  if (!hash_buff || !dev_err_buff || !join_columns || !type_infos || (key_width != 4 && key_width != 8) || entry_count <= 0) {
    set_error("bad argument");
    return HDK_B200_E_INVALID;
  }
  JoinCols cs;
  if (int rc = make_cols(join_columns, type_infos, key_component_count, &cs)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for_join(cs.col[0].num_elems, 256);
  if (key_width == 4)
    fill_baseline_kernel<int32_t><<<grid, 256, 0, st>>>(hash_buff, entry_count, invalid_slot_val, for_semi_join != 0, with_val_slot != 0, dev_err_buff, cs);
  else
    fill_baseline_kernel<int64_t><<<grid, 256, 0, st>>>(hash_buff, entry_count, invalid_slot_val, for_semi_join != 0, with_val_slot != 0, dev_err_buff, cs);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_fill_one_to_many_baseline_hash_table_on_device(int32_t* buff, const int8_t* composite_key_dict,
                                                            int64_t hash_entry_count, int32_t invalid_slot_val,
                                                            size_t key_component_count,
                                                            const hdk_b200_join_column* join_columns,
                                                            const hdk_b200_join_column_type_info* type_infos, int key_width,
                                                            void* stream) {
  if (!buff || !composite_key_dict || !join_columns || !type_infos || (key_width != 4 && key_width != 8) || hash_entry_count <= 0) {
    set_error("bad argument");
    return HDK_B200_E_INVALID;
  }
  JoinCols cs;
  if (int rc = make_cols(join_columns, type_infos, key_component_count, &cs)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for_join(cs.col[0].num_elems, 256);
  int32_t* pos = buff;
  int32_t* cnt = buff + hash_entry_count;
  HB_CUDA(cudaMemsetAsync(cnt, 0, size_t(hash_entry_count) * 4, st));
  if (key_width == 4) count_matches_baseline_kernel<int32_t><<<grid, 256, 0, st>>>(cnt, composite_key_dict, hash_entry_count, cs);
  else count_matches_baseline_kernel<int64_t><<<grid, 256, 0, st>>>(cnt, composite_key_dict, hash_entry_count, cs);
  HB_LAUNCH_CHECK();
  if (int rc = build_positions(pos, cnt, hash_entry_count, invalid_slot_val, st)) return rc;
  HB_CUDA(cudaMemsetAsync(cnt, 0, size_t(hash_entry_count) * 4, st));
  if (key_width == 4) fill_row_ids_baseline_kernel<int32_t><<<grid, 256, 0, st>>>(buff, composite_key_dict, hash_entry_count, cs);
  else fill_row_ids_baseline_kernel<int64_t><<<grid, 256, 0, st>>>(buff, composite_key_dict, hash_entry_count, cs);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_probe_hash_join_on_device(const int32_t* buff, const int64_t* keys, int64_t n, int64_t min_key,
                                       int64_t max_key, int64_t* out, void* stream) {
  if (!buff || !keys || !out || n < 0) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  probe_perfect_kernel<<<grid_for_join(uint64_t(n), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(buff, keys, n, min_key, max_key, out);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_probe_baseline_hash_join_on_device(const int8_t* hash_buff, const int8_t* keys, int64_t n,
                                                size_t key_component_count, int key_width, int64_t entry_count,
                                                int with_val_slot, int64_t* out, void* stream) {
  if (!hash_buff || !keys || !out || n < 0 || (key_width != 4 && key_width != 8) || key_component_count < 1 ||
      key_component_count > size_t(kMaxJoinKeys) || entry_count <= 0) {
    set_error("bad argument");
    return HDK_B200_E_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for_join(uint64_t(n), 256);
  if (key_width == 4) probe_baseline_kernel<int32_t><<<grid, 256, 0, st>>>(hash_buff, keys, n, int(key_component_count), entry_count, with_val_slot != 0, out);
  else probe_baseline_kernel<int64_t><<<grid, 256, 0, st>>>(hash_buff, keys, n, int(key_component_count), entry_count, with_val_slot != 0, out);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

}  // extern "C"
