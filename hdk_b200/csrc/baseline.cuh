// hdk_b200/csrc/baseline.cuh — device side of baseline-hash group-by: open-addressing probe / claim in
// the reference-encoded global buffer.  (Aggregates accumulate in a neutral work table, scan.cu / finalize.cu.)
//
//   hash           key_hash = MurmurHash3_x86_32 over the key bytes, seed 0  (QE/GroupByRuntime.cpp:24-29)
//   probe          h % E, linear probing, NULL ⇒ out of slots                (QE/GroupByRuntime.cpp:31-54, 90-112)
//   claim          the reference's protocol (QE/cuda_mapd_rt.cu:176-236, 240-321): CAS on the first key component; the
//                  winner publishes the rest; others wait until the rest is published, then compare.  Here a key that
//                  fits one atomic is claimed WHOLE — two 4-byte components with one 64-bit CAS, two 8-byte components of
//                  a 16-byte aligned row with one atom.cas.b128 — so there is no half-published entry and nobody waits;
//                  the reference protocol remains for wider keys and the columnar layout, its wait backs off
//                  (__nanosleep) and is bounded (kClaimTimedOut instead of a hang).
#pragma once
#ifndef __CUDACC_RTC__
#include <type_traits>
#endif

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

constexpr int64_t kClaimTimedOut = -2;   // a winner never published the rest of its key (→ error HDK_B200_ERR_CLAIM_TIMEOUT)

// wait until a claimed entry's component is published: poll, back off, give up after ~1 s
template <typename T>
__device__ __forceinline__ bool wait_published(const T* p, T empty, T* out) {
  T v = ld_volatile(p);
  for (uint32_t spins = 0; v == empty; ++spins) {
    if (spins >= (1u << 22)) return false;
    if (spins >= 16) __nanosleep(spins < 1024 ? 32 : 256);
    v = ld_volatile(p);
  }
  *out = v;
  return true;
}

// 16-byte compare-and-swap (sm_90+): returns the old value
__device__ __forceinline__ void cas_b128(void* addr, uint64_t cmp_lo, uint64_t cmp_hi, uint64_t val_lo, uint64_t val_hi, uint64_t& old_lo,
                                         uint64_t& old_hi) {
  asm volatile(
      "{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
      "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
      : "=l"(old_lo), "=l"(old_hi)
      : "l"(cmp_lo), "l"(cmp_hi), "l"(val_lo), "l"(val_hi), "l"(addr)
      : "memory");
}

// Find or claim the entry for `keys` (row-wise layout).  Returns the entry index, -1 (table full) or kClaimTimedOut.
template <typename T>
__device__ __forceinline__ int64_t baseline_claim_rowwise(int8_t* buf, uint32_t row_bytes, uint32_t E, const int64_t* keys,
                                                          int key_count, uint32_t h0) {
  const T empty = sizeof(T) == 4 ? T(HDK_B200_EMPTY_KEY_32) : T(HDK_B200_EMPTY_KEY_64);
  using U = typename std::conditional<sizeof(T) == 4, unsigned int, unsigned long long>::type;
  uint32_t h = h0;
  if (sizeof(T) == 4 && key_count == 2) {
    // two 4-byte components: the whole key is one 64-bit word of the (8-byte aligned) row
    const unsigned long long e64 = (unsigned long long)uint32_t(HDK_B200_EMPTY_KEY_32) * 0x100000001ull;
    const unsigned long long mine = (unsigned long long)uint32_t(keys[0]) | ((unsigned long long)uint32_t(keys[1]) << 32);
    do {
      unsigned long long* row = reinterpret_cast<unsigned long long*>(buf + size_t(h) * row_bytes);
      unsigned long long cur = ld_volatile(row);
      if (cur == e64) cur = atomicCAS(row, e64, mine);
      if (cur == e64 || cur == mine) return h;
      h = h + 1 == E ? 0 : h + 1;
    } while (h != h0);
    return -1;
  }
  if (sizeof(T) == 8 && key_count == 2 && (row_bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(buf) & 15u) == 0) {
    // two 8-byte components of a 16-byte aligned row: one 128-bit CAS.  A half only ever holds EMPTY or its final value,
    // so a plain look that shows both halves equal to the key is a match; anything with an EMPTY half is settled by the CAS.
    const uint64_t e = uint64_t(HDK_B200_EMPTY_KEY_64), k0 = uint64_t(keys[0]), k1 = uint64_t(keys[1]);
    do {
      uint64_t* row = reinterpret_cast<uint64_t*>(buf + size_t(h) * row_bytes);
      uint64_t c0 = ld_volatile(row), c1 = ld_volatile(row + 1);
      if (c0 == e || c1 == e) {
        cas_b128(row, e, e, k0, k1, c0, c1);
        if (c0 == e && c1 == e) return h;   // claimed
      }
      if (c0 == k0 && c1 == k1) return h;
      h = h + 1 == E ? 0 : h + 1;
    } while (h != h0);
    return -1;
  }
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
        if (!wait_published(row + i, empty, &v)) return kClaimTimedOut;
        match = v == T(keys[i]);
      }
      if (match) return h;
    }
    h = h + 1 == E ? 0 : h + 1;
  } while (h != h0);
  return -1;
}

// columnar layout: 8-byte key columns, component i at buf64[i * E + h] (components of one entry lie E words apart: no
// packed claim)
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
        if (!wait_published(buf64 + size_t(i) * E + h, int64_t(HDK_B200_EMPTY_KEY_64), &v)) return kClaimTimedOut;
        match = v == keys[i];
      }
      if (match) return h;
    }
    h = h + 1 == E ? 0 : h + 1;
  } while (h != h0);
  return -1;
}

// ---- baseline JOIN tables (JHT/BaselineJoinHashTable.cpp): E x (key components ‖ payload), MurmurHash1 ----------
__device__ __forceinline__ uint32_t murmur1_dev(const void* key, int len) {
  // MurmurHash1 (QE/MurmurHash1Inl.h), seed 0; len is a multiple of 4 here
  const unsigned int m = 0xc6a4a793u;
  unsigned int h = 0u ^ (unsigned(len) * m);
  const unsigned int* d = static_cast<const unsigned int*>(key);
  for (int i = 0; i < len / 4; ++i) {
    h += d[i];
    h *= m;
    h ^= h >> 16;
  }
  h *= m;
  h ^= h >> 10;
  h *= m;
  h ^= h >> 17;
  return h;
}

// claim-or-find the entry of `key`: CAS the first component, publish the rest, wait for the rest
// (get_matching_baseline_hash_slot_at, HashJoinRuntime.cpp:359-394)
template <typename T>
__device__ __forceinline__ T* baseline_slot(int8_t* hash_buff, int64_t E, const T* key, int kc, size_t entry_sz, bool insert) {
  using U = typename std::conditional<sizeof(T) == 4, unsigned int, unsigned long long>::type;
  const T empty = sizeof(T) == 4 ? T(HDK_B200_EMPTY_KEY_32) : T(HDK_B200_EMPTY_KEY_64);
  const uint32_t h0 = murmur1_dev(key, kc * int(sizeof(T))) % uint32_t(E);
  uint32_t h = h0;
  do {
    T* row = reinterpret_cast<T*>(hash_buff + size_t(h) * entry_sz);
    T first = *reinterpret_cast<volatile T*>(row);
    if (first == empty) {
      if (!insert) return nullptr;
      first = T(atomicCAS(reinterpret_cast<U*>(row), U(empty), U(key[0])));
      if (first == empty) {
        for (int i = 1; i < kc; ++i) atomicExch(reinterpret_cast<U*>(row + i), U(key[i]));
        return row + kc;
      }
    }
    if (first == key[0]) {
      bool match = true;
      for (int i = 1; i < kc && match; ++i) {
        T v;
        if (!wait_published(row + i, empty, &v)) return nullptr;   // (a winner that never publishes: treated as a miss / full table)
        match = v == key[i];
      }
      if (match) return row + kc;
    }
    h = h + 1 == uint32_t(E) ? 0 : h + 1;
  } while (h != h0);
  return nullptr;
}


// probe of the one-to-one layout from the fused scan kernel: row id of the matching inner row, or -1
// (baseline_hash_join_idx_{32,64}, JHT/Runtime/JoinHashTableQueryRuntime.cpp:43-98)
template <typename T>
__device__ __forceinline__ int64_t baseline_join_probe(const int8_t* table, int64_t E, const int64_t* key64, int kc) {
  T key[HDK_B200_MAX_KEYS];
  for (int i = 0; i < kc; ++i) key[i] = T(key64[i]);
  const T* slot = baseline_slot<T>(const_cast<int8_t*>(table), E, key, kc, size_t(kc + 1) * sizeof(T), false);
  return slot ? int64_t(*slot) : -1;
}

// position of the key in the composite-key dictionary of a one-to-many baseline table, or -1
// (get_composite_key_index_{32,64}, JHT/Runtime/JoinHashTableQueryRuntime.cpp:130-171); the table is read-only here
template <typename T>
__device__ __forceinline__ int64_t baseline_dict_index(const int8_t* dict, int64_t E, const int64_t* key64, int kc) {
  const T empty = sizeof(T) == 4 ? T(HDK_B200_EMPTY_KEY_32) : T(HDK_B200_EMPTY_KEY_64);
  T key[HDK_B200_MAX_KEYS];
  for (int i = 0; i < kc; ++i) key[i] = T(key64[i]);
  const uint32_t h0 = murmur1_dev(key, kc * int(sizeof(T))) % uint32_t(E);
  uint32_t h = h0;
  do {
    const T* row = reinterpret_cast<const T*>(dict) + size_t(h) * kc;
    const T first = __ldg(row);
    if (first == empty) return -1;
    bool match = first == key[0];
    for (int i = 1; i < kc && match; ++i) match = __ldg(row + i) == key[i];
    if (match) return int64_t(h);
    h = h + 1 == uint32_t(E) ? 0 : h + 1;
  } while (h != h0);
  return -1;
}

}  // namespace hb
