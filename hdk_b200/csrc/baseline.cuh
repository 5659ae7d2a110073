// hdk_b200/csrc/baseline.cuh — device side of baseline-hash group-by: open-addressing probe / claim in
// the reference-encoded global buffer.  (Aggregates accumulate in a neutral work table, scan.cu / finalize.cu.)
//
//   hash           key_hash = MurmurHash3_x86_32 over the key bytes, seed 0  (QE/GroupByRuntime.cpp:24-29)
//   probe          h % E, linear probing, NULL ⇒ out of slots                (QE/GroupByRuntime.cpp:31-54, 90-112)
//   claim          CAS on the first key component; the winner publishes the rest; others wait until
//                  the rest is published, then compare                      (QE/cuda_mapd_rt.cu:176-236, 240-321)
#pragma once
#include "common.cuh"

namespace hb {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __forceinline__ uint32_t mm3_block(uint32_t h, uint32_t k) {
  k *= 0xcc9e2d51u;
  k = rotl32(k, 15);
  k *= 0x1b873593u;
  h ^= k;
  h = rotl32(h, 13);
  return h * 5u + 0xe6546b64u;
}
__device__ __forceinline__ uint32_t mm3_final(uint32_t h, uint32_t len) {
  h ^= len;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
// keys[] hold the components widened to int64; key_width selects how many bytes of each are hashed
__device__ __forceinline__ uint32_t key_hash_dev(const int64_t* keys, int key_count, int key_width) {
  uint32_t h = 0;
  for (int i = 0; i < key_count; ++i) {
    h = mm3_block(h, uint32_t(uint64_t(keys[i])));
    if (key_width == 8) h = mm3_block(h, uint32_t(uint64_t(keys[i]) >> 32));
  }
  return mm3_final(h, uint32_t(key_count * key_width));
}

// MurmurHash64A over 64-bit-widened keys: the reference's partition function for partitioned
// aggregation (QE/RowFuncBuilder.cpp:516-577)
__device__ __forceinline__ uint64_t murmur64a_keys(const int64_t* keys, int key_count) {
  const uint64_t m = 0xc6a4a7935bd1e995ULL;
  const int r = 47;
  uint64_t h = 0 ^ (uint64_t(key_count) * 8 * m);
  for (int i = 0; i < key_count; ++i) {
    uint64_t k = uint64_t(keys[i]);
    k *= m;
    k ^= k >> r;
    k *= m;
    h ^= k;
    h *= m;
  }
  h ^= h >> r;
  h *= m;
  h ^= h >> r;
  return h;
}

template <typename T>
__device__ __forceinline__ T ld_volatile(const T* p) { return *reinterpret_cast<const volatile T*>(p); }

// Find or claim the entry for `keys` (row-wise layout).  Returns the entry index or -1 (table full).
template <typename T>
__device__ __forceinline__ int64_t baseline_claim_rowwise(int8_t* buf, uint32_t row_bytes, uint32_t E, const int64_t* keys,
                                                          int key_count, uint32_t h0) {
  const T empty = sizeof(T) == 4 ? T(HDK_B200_EMPTY_KEY_32) : T(HDK_B200_EMPTY_KEY_64);
  using U = typename std::conditional<sizeof(T) == 4, unsigned int, unsigned long long>::type;
  uint32_t h = h0;
  do {
    T* row = reinterpret_cast<T*>(buf + size_t(h) * row_bytes);
    const T k0 = T(keys[0]);
    T first = ld_volatile(row);
    if (first == empty) first = T(atomicCAS(reinterpret_cast<U*>(row), U(empty), U(k0)));
    if (first == empty) {  // we own the entry: publish the remaining components
      for (int i = 1; i < key_count; ++i) atomicExch(reinterpret_cast<U*>(row + i), U(T(keys[i])));
      return h;
    }
    if (first == k0) {
      bool match = true;
      for (int i = 1; i < key_count && match; ++i) {
        T v;
        while ((v = ld_volatile(row + i)) == empty) {
        }
        match = v == T(keys[i]);
      }
      if (match) return h;
    }
    h = h + 1 == E ? 0 : h + 1;
  } while (h != h0);
  return -1;
}

// columnar layout: 8-byte key columns, component i at buf64[i * E + h]
__device__ __forceinline__ int64_t baseline_claim_columnar(int64_t* buf64, uint32_t E, const int64_t* keys, int key_count,
                                                           uint32_t h0) {
  uint32_t h = h0;
  do {
    int64_t first = ld_volatile(buf64 + h);
    if (first == HDK_B200_EMPTY_KEY_64)
      first = int64_t(atomicCAS(reinterpret_cast<unsigned long long*>(buf64 + h), (unsigned long long)HDK_B200_EMPTY_KEY_64,
                                (unsigned long long)keys[0]));
    if (first == HDK_B200_EMPTY_KEY_64) {
      for (int i = 1; i < key_count; ++i)
        atomicExch(reinterpret_cast<unsigned long long*>(buf64 + size_t(i) * E + h), (unsigned long long)keys[i]);
      return h;
    }
    if (first == keys[0]) {
      bool match = true;
      for (int i = 1; i < key_count && match; ++i) {
        int64_t v;
        while ((v = ld_volatile(buf64 + size_t(i) * E + h)) == HDK_B200_EMPTY_KEY_64) {
        }
        match = v == keys[i];
      }
      if (match) return h;
    }
    h = h + 1 == E ? 0 : h + 1;
  } while (h != h0);
  return -1;
}

}  // namespace hb
