/*
 * oracle/rt_port.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the per-row runtime functions of intel/hdk that sit on
 * the fused scan/filter/group-by/aggregate and hash-join path.  The function NAMES and
 * SIGNATURES match the reference's extern "C" runtime so that oracle/driver.inc can be
 * compiled either against this port (liboracle_port.so) or against the reference's own
 * sources compiled where they lie (oracle/_ref/liboracle_ref.so, see Makefile).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use it.  Each function cites the reference file:line it follows
 * (QE = omniscidb/QueryEngine).
 */
#pragma once
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <limits>

#define ORC_EMPTY_KEY_64 INT64_C(9223372036854775807) /* QE/GpuRtConstants.h:29 */
#define ORC_EMPTY_KEY_32 2147483647                   /* QE/GpuRtConstants.h:30 */

/* ---------------------------------------------------------------- hashes ---
 * MurmurHash3_x86_32 (QE/MurmurHash3Inl.h), MurmurHash1 (QE/MurmurHash1Inl.h) and
 * MurmurHash64A (QE/MurmurHash1Inl.h:66-…) are Austin Appleby's published
 * algorithms; restated from the published definitions. */
static inline uint32_t orc_rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

extern "C" inline uint32_t MurmurHash3(const void* key, int64_t len, const uint32_t seed) {
  const uint8_t* p = static_cast<const uint8_t*>(key);
  const int64_t nblocks = len / 4;
  uint32_t h = seed;
  for (int64_t i = 0; i < nblocks; ++i) {
    uint32_t k;
    std::memcpy(&k, p + 4 * i, 4);
    k *= 0xcc9e2d51u;
    k = orc_rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = orc_rotl32(h, 13);
    h = h * 5 + 0xe6546b64u;
  }
  const uint8_t* tail = p + 4 * nblocks;
  uint32_t k = 0;
  switch (len & 3) {
    case 3: k ^= uint32_t(tail[2]) << 16; /* fallthrough */
    case 2: k ^= uint32_t(tail[1]) << 8;  /* fallthrough */
    case 1:
      k ^= tail[0];
      k *= 0xcc9e2d51u;
      k = orc_rotl32(k, 15);
      k *= 0x1b873593u;
      h ^= k;
  }
  h ^= uint32_t(len);
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

extern "C" inline uint32_t MurmurHash1(const void* key, int len, const uint32_t seed) {
  const unsigned int m = 0xc6a4a793u;
  const int r = 16;
  unsigned int h = seed ^ (unsigned(len) * m);
  const unsigned char* data = static_cast<const unsigned char*>(key);
  while (len >= 4) {
    unsigned int k;
    std::memcpy(&k, data, 4);
    h += k;
    h *= m;
    h ^= h >> 16;
    data += 4;
    len -= 4;
  }
  switch (len) {
    case 3: h += unsigned(data[2]) << 16; /* fallthrough */
    case 2: h += unsigned(data[1]) << 8;  /* fallthrough */
    case 1:
      h += data[0];
      h *= m;
      h ^= h >> r;
  }
  h *= m;
  h ^= h >> 10;
  h *= m;
  h ^= h >> 17;
  return h;
}

extern "C" inline uint64_t MurmurHash64A(const void* key, int len, uint64_t seed) {
  const uint64_t m = 0xc6a4a7935bd1e995ULL;
  const int r = 47;
  uint64_t h = seed ^ (uint64_t(len) * m);
  const unsigned char* data = static_cast<const unsigned char*>(key);
  const int nblk = len / 8;
  for (int i = 0; i < nblk; ++i) {
    uint64_t k;
    std::memcpy(&k, data + 8 * i, 8);
    k *= m;
    k ^= k >> r;
    k *= m;
    h ^= k;
    h *= m;
  }
  const unsigned char* t = data + 8 * nblk;
  switch (len & 7) {
    case 7: h ^= uint64_t(t[6]) << 48; /* fallthrough */
    case 6: h ^= uint64_t(t[5]) << 40; /* fallthrough */
    case 5: h ^= uint64_t(t[4]) << 32; /* fallthrough */
    case 4: h ^= uint64_t(t[3]) << 24; /* fallthrough */
    case 3: h ^= uint64_t(t[2]) << 16; /* fallthrough */
    case 2: h ^= uint64_t(t[1]) << 8;  /* fallthrough */
    case 1:
      h ^= uint64_t(t[0]);
      h *= m;
  }
  h ^= h >> r;
  h *= m;
  h ^= h >> r;
  return h;
}

/* QE/GroupByRuntime.cpp:24-29 */
extern "C" inline uint32_t key_hash(const int64_t* key, const uint32_t key_count,
                                    const uint32_t key_byte_width) {
  return MurmurHash3(key, int64_t(key_byte_width) * key_count, 0);
}

/* ------------------------------------------------------------- decoders ---
 * QE/DecodersImpl.h:31-60 (signed), :62-92 (unsigned), :122-140 (fp), :153-161 (date) */
extern "C" inline int64_t fixed_width_int_decode(const int8_t* s, const int32_t w, const int64_t pos) {
  switch (w) {
    case 1: return s[pos];
    case 2: { int16_t v; std::memcpy(&v, s + 2 * pos, 2); return v; }
    case 4: { int32_t v; std::memcpy(&v, s + 4 * pos, 4); return v; }
    case 8: { int64_t v; std::memcpy(&v, s + 8 * pos, 8); return v; }
    default: return std::numeric_limits<int64_t>::min() + 1;
  }
}
extern "C" inline int64_t fixed_width_unsigned_decode(const int8_t* s, const int32_t w, const int64_t pos) {
  switch (w) {
    case 1: return uint8_t(s[pos]);
    case 2: { uint16_t v; std::memcpy(&v, s + 2 * pos, 2); return v; }
    case 4: { uint32_t v; std::memcpy(&v, s + 4 * pos, 4); return v; }
    case 8: { uint64_t v; std::memcpy(&v, s + 8 * pos, 8); return int64_t(v); }
    default: return std::numeric_limits<int64_t>::min() + 1;
  }
}
extern "C" inline float fixed_width_float_decode(const int8_t* s, const int64_t pos) {
  float v; std::memcpy(&v, s + 4 * pos, 4); return v;
}
extern "C" inline double fixed_width_double_decode(const int8_t* s, const int64_t pos) {
  double v; std::memcpy(&v, s + 8 * pos, 8); return v;
}
extern "C" inline int64_t fixed_width_small_date_decode(const int8_t* s, const int32_t w,
                                                        const int32_t null_val,
                                                        const int64_t ret_null_val, const int64_t pos) {
  const int64_t v = fixed_width_int_decode(s, w, pos);
  return v == null_val ? ret_null_val : v * 86400;
}

/* ---------------------------------------------------------- time extract ---
 * omniscidb/Utils/ExtractFromTime.cpp:156-163 (fast path), :260-271; constants
 * Utils/ExtractFromTime.h:33-72 */
extern "C" inline int64_t extract_year(const int64_t timeval) {
  const uint32_t kEpochOffsetYear1900 = 2208988800u, kSecsJanToMar1900 = 5097600u;
  const uint32_t kSecondsPer4YearCycle = 126230400u, kSecondsPerNonLeapYear = 31536000u;
  if (timeval >= 0 && timeval <= int64_t(UINT32_MAX - kEpochOffsetYear1900)) {
    const uint32_t s1900 = uint32_t(timeval) + kEpochOffsetYear1900;
    const uint32_t leap = (s1900 - kSecsJanToMar1900) / kSecondsPer4YearCycle;
    return (s1900 - leap * 86400u) / kSecondsPerNonLeapYear + 1900;
  }
  auto floor_div = [](int64_t a, int64_t b) { return (a < 0 ? a - (b - 1) : a) / b; };
  const int64_t kEpochAdjustedDays = 11017, kDaysPer400Years = 146097;
  const unsigned MARJAN = 31 + 30 + 31 + 30 + 31 + 31 + 30 + 31 + 30 + 31;
  const int64_t day = floor_div(timeval, 86400);
  const int64_t era = floor_div(day - kEpochAdjustedDays, kDaysPer400Years);
  const unsigned doe = unsigned(day - kEpochAdjustedDays - era * kDaysPer400Years);
  const unsigned yoe = (doe - doe / 1460 + doe / 36524 - (doe == 146096)) / 365;
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  return 2000 + era * 400 + yoe + (MARJAN <= doy);
}

/* QE/RuntimeFunctions.cpp:259-271 */
extern "C" inline int64_t floor_div_lhs(const int64_t dividend, const int64_t divisor) {
  return (dividend < 0 ? dividend - (divisor - 1) : dividend) / divisor;
}
extern "C" inline int64_t floor_div_nullable_lhs(const int64_t dividend, const int64_t divisor,
                                                 const int64_t null_val) {
  return dividend == null_val ? null_val : floor_div_lhs(dividend, divisor);
}

/* ------------------------------------------------------------ aggregates ---
 * QE/RuntimeFunctions.cpp:388-391, 456-476, 528-560 (ints), 612-703 (skip_val ints),
 * 708-736, 773-800 (fp), 821-880 (skip_val fp) */
extern "C" inline uint64_t agg_count(uint64_t* agg, const int64_t) { return (*agg)++; }
extern "C" inline uint32_t agg_count_int32(uint32_t* agg, const int32_t) { return (*agg)++; }
extern "C" inline uint64_t agg_count_double(uint64_t* agg, const double) { return (*agg)++; }
extern "C" inline uint32_t agg_count_float(uint32_t* agg, const float) { return (*agg)++; }
extern "C" inline int64_t agg_sum(int64_t* agg, const int64_t val) { const int64_t old = *agg; *agg = int64_t(uint64_t(old) + uint64_t(val)); return old; }
extern "C" inline int32_t agg_sum_int32(int32_t* agg, const int32_t val) { const int32_t old = *agg; *agg = int32_t(uint32_t(old) + uint32_t(val)); return old; }
extern "C" inline void agg_max(int64_t* agg, const int64_t val) { *agg = std::max(*agg, val); }
extern "C" inline void agg_min(int64_t* agg, const int64_t val) { *agg = std::min(*agg, val); }
extern "C" inline void agg_id(int64_t* agg, const int64_t val) { *agg = val; }
extern "C" inline void agg_max_int32(int32_t* agg, const int32_t val) { *agg = std::max(*agg, val); }
extern "C" inline void agg_min_int32(int32_t* agg, const int32_t val) { *agg = std::min(*agg, val); }
extern "C" inline void agg_id_int32(int32_t* agg, const int32_t val) { *agg = val; }

extern "C" inline int64_t agg_sum_skip_val(int64_t* agg, const int64_t val, const int64_t skip_val) {
  const int64_t old = *agg;
  if (val != skip_val) {
    if (old != skip_val) return agg_sum(agg, val);
    *agg = val;
  }
  return old;
}
extern "C" inline int32_t agg_sum_int32_skip_val(int32_t* agg, const int32_t val, const int32_t skip_val) {
  const int32_t old = *agg;
  if (val != skip_val) {
    if (old != skip_val) return agg_sum_int32(agg, val);
    *agg = val;
  }
  return old;
}
extern "C" inline uint64_t agg_count_skip_val(uint64_t* agg, const int64_t val, const int64_t skip_val) {
  return val != skip_val ? agg_count(agg, val) : *agg;
}
extern "C" inline uint32_t agg_count_int32_skip_val(uint32_t* agg, const int32_t val, const int32_t skip_val) {
  return val != skip_val ? agg_count_int32(agg, val) : *agg;
}
#define ORC_SKIP_INT(NAME, T)                                                         \
  extern "C" inline void NAME##_skip_val(T* agg, const T val, const T skip_val) {     \
    if (val != skip_val) {                                                            \
      if (*agg != skip_val) NAME(agg, val); else *agg = val;                          \
    }                                                                                 \
  }
ORC_SKIP_INT(agg_max, int64_t)
ORC_SKIP_INT(agg_min, int64_t)
ORC_SKIP_INT(agg_max_int32, int32_t)
ORC_SKIP_INT(agg_min_int32, int32_t)
#undef ORC_SKIP_INT

template <class F, class I> static inline F orc_bits_to(I bits) { F f; std::memcpy(&f, &bits, sizeof(F)); return f; }
template <class I, class F> static inline I orc_to_bits(F f) { I b; std::memcpy(&b, &f, sizeof(I)); return b; }

extern "C" inline void agg_sum_double(int64_t* agg, const double val) { *agg = orc_to_bits<int64_t>(orc_bits_to<double>(*agg) + val); }
extern "C" inline void agg_max_double(int64_t* agg, const double val) { *agg = orc_to_bits<int64_t>(std::max(orc_bits_to<double>(*agg), val)); }
extern "C" inline void agg_min_double(int64_t* agg, const double val) { *agg = orc_to_bits<int64_t>(std::min(orc_bits_to<double>(*agg), val)); }
extern "C" inline void agg_id_double(int64_t* agg, const double val) { *agg = orc_to_bits<int64_t>(val); }
extern "C" inline void agg_sum_float(int32_t* agg, const float val) { *agg = orc_to_bits<int32_t>(orc_bits_to<float>(*agg) + val); }
extern "C" inline void agg_max_float(int32_t* agg, const float val) { *agg = orc_to_bits<int32_t>(std::max(orc_bits_to<float>(*agg), val)); }
extern "C" inline void agg_min_float(int32_t* agg, const float val) { *agg = orc_to_bits<int32_t>(std::min(orc_bits_to<float>(*agg), val)); }
extern "C" inline void agg_id_float(int32_t* agg, const float val) { *agg = orc_to_bits<int32_t>(val); }
extern "C" inline uint64_t agg_count_double_skip_val(uint64_t* agg, const double val, const double skip_val) {
  return val != skip_val ? agg_count_double(agg, val) : *agg;
}
extern "C" inline uint32_t agg_count_float_skip_val(uint32_t* agg, const float val, const float skip_val) {
  return val != skip_val ? agg_count_float(agg, val) : *agg;
}
#define ORC_SKIP_FP(NAME, ADDR_T, DATA_T)                                                       \
  extern "C" inline void NAME##_skip_val(ADDR_T* agg, const DATA_T val, const DATA_T skip_val) { \
    if (val != skip_val) {                                                                       \
      if (*agg != orc_to_bits<ADDR_T>(skip_val)) NAME(agg, val);                                 \
      else *agg = orc_to_bits<ADDR_T>(val);                                                      \
    }                                                                                            \
  }
ORC_SKIP_FP(agg_sum_double, int64_t, double)
ORC_SKIP_FP(agg_max_double, int64_t, double)
ORC_SKIP_FP(agg_min_double, int64_t, double)
ORC_SKIP_FP(agg_sum_float, int32_t, float)
ORC_SKIP_FP(agg_max_float, int32_t, float)
ORC_SKIP_FP(agg_min_float, int32_t, float)
#undef ORC_SKIP_FP

/* ----------------------------------------------------- group-by probes ---
 * QE/RuntimeFunctions.cpp:1210-1250 (row-wise match/claim), :1252-1300 (columnar slot),
 * :1339-1382 (perfect hash multi-key), :1387-1395 (keyless);
 * QE/GroupByRuntime.cpp:31-54, 90-112, 198-246 */
template <typename T>
static inline int64_t* orc_match_rowwise(int64_t* buf, const uint32_t h, const T* key,
                                         const uint32_t key_count, const uint32_t row_size_quad) {
  const T empty = sizeof(T) == 4 ? T(ORC_EMPTY_KEY_32) : T(ORC_EMPTY_KEY_64);
  T* row = reinterpret_cast<T*>(buf + uint64_t(h) * row_size_quad);
  auto slots = [&]() {
    uintptr_t p = reinterpret_cast<uintptr_t>(row + key_count);
    return reinterpret_cast<int64_t*>((p + 7) & ~uintptr_t(7));
  };
  if (*row == empty) {
    std::memcpy(row, key, key_count * sizeof(T));
    return slots();
  }
  if (std::memcmp(row, key, key_count * sizeof(T)) == 0) return slots();
  return nullptr;
}
extern "C" inline int64_t* get_matching_group_value(int64_t* buf, const uint32_t h, const int64_t* key,
                                                    const uint32_t key_count, const uint32_t key_width,
                                                    const uint32_t row_size_quad) {
  if (key_width == 4) return orc_match_rowwise(buf, h, reinterpret_cast<const int32_t*>(key), key_count, row_size_quad);
  if (key_width == 8) return orc_match_rowwise(buf, h, key, key_count, row_size_quad);
  return nullptr;
}
extern "C" inline int64_t* get_group_value(int64_t* buf, const uint32_t entry_count, const int64_t* key,
                                           const uint32_t key_count, const uint32_t key_width,
                                           const uint32_t row_size_quad) {
  const uint32_t h0 = key_hash(key, key_count, key_width) % entry_count;
  uint32_t h = h0;
  do {
    if (int64_t* g = get_matching_group_value(buf, h, key, key_count, key_width, row_size_quad)) return g;
    h = (h + 1) % entry_count;
  } while (h != h0);
  return nullptr;
}
template <typename T>
static inline int32_t orc_match_columnar(int64_t* buf, const uint32_t entry_count, const uint32_t h,
                                         const T* key, const uint32_t key_count) {
  const T empty = sizeof(T) == 4 ? T(ORC_EMPTY_KEY_32) : T(ORC_EMPTY_KEY_64);
  T* kb = reinterpret_cast<T*>(buf);
  if (kb[h] == empty) {
    for (uint32_t i = 0; i < key_count; ++i) kb[h + uint64_t(i) * entry_count] = key[i];
    return int32_t(h);
  }
  for (uint32_t i = 0; i < key_count; ++i)
    if (kb[h + uint64_t(i) * entry_count] != key[i]) return -1;
  return int32_t(h);
}
extern "C" inline int32_t get_group_value_columnar_slot(int64_t* buf, const uint32_t entry_count,
                                                        const int64_t* key, const uint32_t key_count,
                                                        const uint32_t key_width) {
  const uint32_t h0 = key_hash(key, key_count, key_width) % entry_count;
  uint32_t h = h0;
  do {
    const int32_t m = key_width == 4
                          ? orc_match_columnar(buf, entry_count, h, reinterpret_cast<const int32_t*>(key), key_count)
                          : orc_match_columnar(buf, entry_count, h, key, key_count);
    if (m != -1) return int32_t(h);
    h = (h + 1) % entry_count;
  } while (h != h0);
  return -1;
}
extern "C" inline int64_t* get_group_value_fast(int64_t* buf, const int64_t key, const int64_t min_key,
                                                const int64_t bucket, const uint32_t row_size_quad) {
  int64_t d = key - min_key;
  if (bucket) d /= bucket;
  const int64_t off = d * row_size_quad;
  if (buf[off] == ORC_EMPTY_KEY_64) buf[off] = key;
  return buf + off + 1;
}
extern "C" inline int64_t* get_group_value_fast_keyless(int64_t* buf, const int64_t key, const int64_t min_key,
                                                        const int64_t, const uint32_t row_size_quad) {
  return buf + row_size_quad * (key - min_key);
}
extern "C" inline uint32_t get_columnar_group_bin_offset(int64_t* key_base, const int64_t key,
                                                         const int64_t min_key, const int64_t bucket) {
  int64_t off = key - min_key;
  if (bucket) off /= bucket;
  if (key_base[off] == ORC_EMPTY_KEY_64) key_base[off] = key;
  return uint32_t(off);
}
extern "C" inline int64_t* get_matching_group_value_perfect_hash(int64_t* buf, const uint32_t h,
                                                                 const int64_t* key, const uint32_t key_count,
                                                                 const uint32_t row_size_quad) {
  const uint32_t off = h * row_size_quad;
  if (buf[off] == ORC_EMPTY_KEY_64)
    for (uint32_t i = 0; i < key_count; ++i) buf[off + i] = key[i];
  return buf + off + key_count;
}
extern "C" inline int64_t* get_matching_group_value_perfect_hash_keyless(int64_t* buf, const uint32_t h,
                                                                         const uint32_t row_size_quad) {
  return buf + uint64_t(row_size_quad) * h;
}
extern "C" inline void set_matching_group_value_perfect_hash_columnar(int64_t* buf, const uint32_t h,
                                                                      const int64_t* key, const uint32_t key_count,
                                                                      const uint32_t entry_count) {
  if (buf[h] == ORC_EMPTY_KEY_64)
    for (uint32_t i = 0; i < key_count; ++i) buf[uint64_t(i) * entry_count + h] = key[i];
}

/* QE/GroupByRuntime.cpp:298-329 + JHT/Runtime/JoinHashImpl.h:86-97 */
extern "C" inline int64_t hash_join_idx(int64_t hash_buff, const int64_t key, const int64_t min_key,
                                        const int64_t max_key) {
  if (key >= min_key && key <= max_key) return reinterpret_cast<const int32_t*>(hash_buff)[key - min_key];
  return -1;
}
extern "C" inline int64_t hash_join_idx_nullable(int64_t hash_buff, const int64_t key, const int64_t min_key,
                                                 const int64_t max_key, const int64_t null_val) {
  return key != null_val ? hash_join_idx(hash_buff, key, min_key, max_key) : -1;
}

/* JHT/Runtime/JoinHashTableQueryRuntime.cpp:43-98 */
template <typename T>
static inline int64_t orc_baseline_probe(const int8_t* hash_buff, const int8_t* key, const size_t key_bytes,
                                         const size_t entry_count) {
  /* entry = key bytes followed by one T payload; -1 = no match (kNoMatch), -2 = hit an
   * empty entry (kNotPresent) — both mean "no join partner" to the caller. */
  if (!entry_count) return -1;
  const T empty = sizeof(T) == 4 ? T(ORC_EMPTY_KEY_32) : T(ORC_EMPTY_KEY_64);
  const uint32_t h0 = MurmurHash1(key, int(key_bytes), 0) % entry_count;
  uint32_t h = h0;
  do {
    const int8_t* row = hash_buff + size_t(h) * (key_bytes + sizeof(T));
    T first, payload;
    std::memcpy(&first, row, sizeof(T));
    if (std::memcmp(row, key, key_bytes) == 0) {
      std::memcpy(&payload, row + key_bytes, sizeof(T));
      return int64_t(payload);
    }
    if (first == empty) return -2;
    h = (h + 1) % entry_count;
  } while (h != h0);
  return -1;
}
extern "C" inline int64_t baseline_hash_join_idx_32(const int8_t* hash_buff, const int8_t* key,
                                                    const size_t key_bytes, const size_t entry_count) {
  return orc_baseline_probe<int32_t>(hash_buff, key, key_bytes, entry_count);
}
extern "C" inline int64_t baseline_hash_join_idx_64(const int8_t* hash_buff, const int8_t* key,
                                                    const size_t key_bytes, const size_t entry_count) {
  return orc_baseline_probe<int64_t>(hash_buff, key, key_bytes, entry_count);
}

/* get_composite_key_index_{32,64} (JHT/Runtime/JoinHashTableQueryRuntime.cpp:130-171): index of the key in the
 * composite-key dictionary of a one-to-many baseline table (entries = the key components only), or -1 */
template <typename T>
static inline int64_t orc_composite_key_index(const T* key, const size_t kc, const T* dict, const size_t entry_count) {
  if (!entry_count) return -1;
  const T empty = sizeof(T) == 4 ? T(ORC_EMPTY_KEY_32) : T(ORC_EMPTY_KEY_64);
  const uint32_t h = MurmurHash1(key, int(kc * sizeof(T)), 0) % entry_count;
  if (std::memcmp(dict + size_t(h) * kc, key, kc * sizeof(T)) == 0) return h;
  uint32_t hp = (h + 1) % entry_count;
  while (hp != h) {
    const T* row = dict + size_t(hp) * kc;
    if (std::memcmp(row, key, kc * sizeof(T)) == 0) return hp;
    if (row[0] == empty) return -1;
    hp = (hp + 1) % entry_count;
  }
  return -1;
}
extern "C" inline int64_t get_composite_key_index_32(const int32_t* key, const size_t kc, const int32_t* dict,
                                                     const size_t entry_count) {
  return orc_composite_key_index<int32_t>(key, kc, dict, entry_count);
}
extern "C" inline int64_t get_composite_key_index_64(const int64_t* key, const size_t kc, const int64_t* dict,
                                                     const size_t entry_count) {
  return orc_composite_key_index<int64_t>(key, kc, dict, entry_count);
}
