// hdk_b200/csrc/partagg.cu — radix-partitioned baseline-hash aggregation (high-cardinality group-by).
//
// The reference aggregates a baseline-hash group-by by probing ONE open-addressing table per kernel
// (get_group_value, QE/GroupByRuntime.cpp:31-54; GPU claim protocol QE/cuda_mapd_rt.cu:176-236) — on a GPU that is a
// random DRAM sector plus several global atomics per row once the table outgrows L2 (config 4: 2e8 entries).  Its own
// CPU answer to that shape is partitioned aggregation (QE/RelAlgExecutor.cpp:691-838: partition the rows by a hash of
// the key — QE/RowFuncBuilder.cpp:516-577 — so that every partition aggregates on its own).  Here the same idea runs
// INSIDE one GPU, down to partitions whose groups fit a CTA's shared memory:
//
//   1. count      every row's key → partition = key hash range-reduced to P = F1 x F2 (<= 255 x 128); per-CTA shared histogram
//   2. offsets    exclusive scan of the P counters (one CTA)
//   3. scatter    one or two passes, each splitting <= 256 ways.  A row that passes the filters becomes a packed RECORD
//                 (key values ‖ aggregate arguments, 4-byte words).  A CTA regroups a tile of 4096 rows by destination
//                 in shared memory, reserves one run per destination with a single global atomic and writes whole runs —
//                 consecutive lanes, consecutive words — so every pass reads and writes HBM sequentially.  (One direct
//                 32768-way scatter was measured first: 5 partial-sector stores and one returning global atomic per row
//                 ran at 7.8 G rows/s, the stores alone at the same speed, the atomics alone at 20 G rows/s.)
//   4. aggregate  one CTA per partition: records streamed into shared memory with cp.async, open addressing in SHARED
//                 memory keyed by the row id of the group's first ("representative") record — claiming is one 32-bit CAS
//                 and keys are compared against the representative's record, so there is no multi-word publish protocol —
//                 neutral accumulators beside it, native 32-bit shared atomics (64-bit integer SUM = two 32-bit adds with
//                 carry).  The finished groups are written straight into the caller's group-by buffer in the reference
//                 layout (keys, "skip_val" NULL protocol, compact slot widths: the same encoding finalize.cu produces),
//                 at consecutive entries reserved with one global atomic per partition.  The order of entries in a
//                 baseline-hash buffer is free (ResultSet iteration skips EMPTY keys, the reduction re-inserts by key);
//                 the rest of the buffer keeps the caller's init pattern.
//
// A partition whose groups do not fit the shared table is split by further hash bits and re-read (no global fallback).
// Pre-compiled plan shapes (static_shapes.inc) instantiate every kernel with the plan's structure as a constant; any
// other eligible plan runs the same kernels over the interpreter.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "accum.cuh"
#include "baseline.cuh"
#include "common.cuh"
#include "eval.cuh"
#include "partagg.cuh"
#include "shape.cuh"

namespace hb {

constexpr uint32_t kPaMaxPartitions = 255 * 128;   // <= 255 level-1 destinations x <= 128 level-2 destinations
constexpr uint32_t kPaMaxFragments = 4096;
constexpr uint32_t kPaFanout = 255;          // destinations of one scatter pass (a tile's per-vector destination is a byte, 0xff = padding)
constexpr int kPaThreads = 512;
constexpr int kPaMaxTileRows = 4096;         // rows regrouped per tile (ranks fit 16 bits, destinations 8)
constexpr int kPaCountRows = 8;              // rows per thread per tile of the count pass
constexpr int kAggThreads = 1024;
constexpr uint32_t kEmptyId = 0xffffffffu;

struct PaArgs {
  DPlan plan;
  PaLayout lay;
  int32_t key_width;
  // level 1 source: the fragments' columns
  const int8_t* const* col_buffers;
  const int64_t* num_rows;
  uint32_t num_fragments;
  // level 2 source: the records level 1 wrote, one region per level-1 destination
  const uint32_t* src_recs;
  const uint32_t* src_count;       // [F1] records actually written per level-1 destination
  uint32_t P, F1, F2_log2;
  uint32_t n_dest;                 // destinations of this pass
  uint32_t tile_rows;
  uint32_t* counts;                // [P]      rows per final partition (count pass)
  const unsigned long long* base;  // [P + 1]  first record of each final partition
  uint32_t* cursor;                // this pass's cursors: level 1 [F1] (or [P] when it is the only level), level 2 [P]
  uint32_t* dst_recs;
  int32_t* error_codes;
  const int* skewed;               // set by the offsets kernel: a partition is too heavy for one CTA, the fallback path runs instead
};

// Hash of the (cast) key values: 32 x 32 → 64-bit multiply mixing (the scheme of wyhash32: xor the key's halves into two
// 32-bit lanes, multiply the lanes, take the product's halves as the new lanes — three instructions per step).  The
// reference's partitioned aggregation hashes with MurmurHash64A (QE/RowFuncBuilder.cpp:516-577; shuffle.cu keeps it for
// the multi-GPU exchange); inside one GPU any deterministic function of the key serves, and this one costs a third of
// the instructions — it is evaluated four times per row (count, two scatter levels, aggregation).
// Bits: upper half → partition (range reduction of its top bits) and fingerprint / split bits (its low bits);
// lower half → bucket of the shared table (top bits) and further split bits (low bits).
struct KeyHasher {
  uint32_t a, b;
  __device__ __forceinline__ void mix() {
    const uint64_t c = uint64_t(a ^ 0x53c5ca59u) * uint64_t(b ^ 0x74743c1bu);
    a = uint32_t(c);
    b = uint32_t(c >> 32);
  }
  __device__ __forceinline__ explicit KeyHasher(int n_keys) : a(0x9E3779B9u + uint32_t(n_keys)), b(0x85EBCA6Bu) { mix(); }
  __device__ __forceinline__ void add(int64_t key) {
    a ^= uint32_t(uint64_t(key));
    b ^= uint32_t(uint64_t(key) >> 32);
    mix();
  }
  // each half of the result is lane ^ lane: the upper half of a product alone is far from uniform (measured: partition
  // sizes with a standard deviation 20 x Poisson's when it fed the range reduction directly)
  __device__ __forceinline__ uint64_t finish() {
    mix();
    mix();
    const uint32_t hi = a ^ b;
    mix();
    return (uint64_t(hi) << 32) | (a ^ b);
  }
};
// 7 bits: an id whose top byte is 0xff is always the EMPTY id
__device__ __forceinline__ uint32_t pa_fingerprint(uint64_t h) { return (uint32_t(h >> 32) & 0x7fu) << 24; }
// bits that decide the sub-pass of a split partition: independent of the partition (top of the upper half) and of the
// bucket (top of the lower half)
__device__ __forceinline__ uint32_t pa_split_bits(uint64_t h) { return ((uint32_t(h >> 40) & 0xffu) | ((uint32_t(h) & 0x3ffu) << 8)); }
__device__ __forceinline__ uint32_t pa_partition(uint64_t h, uint32_t P) { return __umulhi(uint32_t(h >> 32), P); }
__device__ __forceinline__ int64_t pa_key_cast(int64_t v, int key_width) { return key_width == 4 ? int64_t(int32_t(v)) : v; }

__device__ __forceinline__ uint64_t pa_ld_elem(const int8_t* ptr, int w) {
  return w == 8 ? *reinterpret_cast<const uint64_t*>(ptr) : w == 4 ? uint64_t(*reinterpret_cast<const uint32_t*>(ptr))
         : w == 2 ? uint64_t(*reinterpret_cast<const uint16_t*>(ptr)) : uint64_t(*reinterpret_cast<const uint8_t*>(ptr));
}

// ---- row evaluation ----------------------------------------------------------------------------------------------
// Same error rules as the scan kernel (scan.cu): an error inside a qual is raised whether or not the row passes, any
// other only for rows that pass.  kFull = false (count pass, first half of a scatter tile): only "does the row reach the
// aggregation" and its keys are wanted — errors are left to the pass that writes the records, and with a compile-time
// plan the loads and nodes nobody uses are dropped by the compiler.  Returns false when the row contributes nothing.
template <class Shape, bool kFull>
__device__ __forceinline__ bool pa_eval_row(const DPlan& rp, const int8_t* const* cols, uint64_t pos, V* vals, int32_t& my_err) {
  auto load_inner = [&](int, int, int) -> uint64_t { return 0; };
  int32_t row_err = 0, qual_err = 0;
  if constexpr (Shape::is_static) {
    constexpr DPlan sp = Shape::get();
    auto load_outer = [&](int c, int w) -> uint64_t { return pa_ld_elem(cols[c] + pos * uint64_t(w), w); };
    static_for<0, sp.n_exprs>([&](auto I) {
      constexpr int n = decltype(I)::value;
      constexpr DPlan sp = Shape::get();
      DExpr e = sp.exprs[n];
      if constexpr (sp.exprs[n].op == HDK_B200_OP_CONST) e.imm = rp.exprs[n].imm;   // literals are run-time
      int32_t err = 0;
      vals[n] = eval_node(sp, e, vals, err, load_outer, load_inner);
      if constexpr ((sp.exprs[n].aux & kAuxInQual) != 0) { if (err && !qual_err) qual_err = err; }
      else { if (err && !row_err) row_err = err; }
    });
    if (qual_err) { if (kFull) my_err = my_err > 0 ? my_err : qual_err; return false; }
    bool pass = true;
    static_for<0, sp.n_filters>([&](auto F) {
      constexpr DPlan sp = Shape::get();
      pass = pass && (vals[sp.filters[decltype(F)::value]].i > 0);
    });
    if (!pass) return false;
  } else {
    auto load_outer = [&](int c, int w) -> uint64_t { return pa_ld_elem(cols[c] + pos * uint64_t(w), w); };
    for (int n = 0; n < rp.n_exprs; ++n) {
      int32_t e = 0;
      vals[n] = eval_node(rp, rp.exprs[n], vals, e, load_outer, load_inner);
      if (e) { int32_t& dst = (rp.exprs[n].aux & kAuxInQual) ? qual_err : row_err; if (!dst) dst = e; }
    }
    if (qual_err) { if (kFull) my_err = my_err > 0 ? my_err : qual_err; return false; }
    for (int f = 0; f < rp.n_filters; ++f)
      if (!(vals[rp.filters[f]].i > 0)) return false;
  }
  if (kFull && row_err) { my_err = my_err > 0 ? my_err : row_err; return false; }
  return true;
}

template <class Shape>
__device__ __forceinline__ uint64_t pa_hash_vals(const PaLayout& L, int key_width, const V* vals) {
  if constexpr (Shape::is_static) {
    constexpr PaLayout SL = pa_layout_of(Shape::get());
    KeyHasher kh(SL.n_keys);
    static_for<0, SL.n_keys>([&](auto K) {
      constexpr PaLayout SL = pa_layout_of(Shape::get());
      kh.add(pa_key_cast(vals[SL.f[decltype(K)::value].expr].i, key_width));
    });
    return kh.finish();
  } else {
    KeyHasher kh(L.n_keys);
    for (int k = 0; k < L.n_keys; ++k) kh.add(pa_key_cast(vals[L.f[k].expr].i, key_width));
    return kh.finish();
  }
}

__device__ __forceinline__ int64_t pa_field_value(const uint32_t* rec, const PaField& f) {
  if (f.words == 2) return int64_t(uint64_t(rec[f.off]) | (uint64_t(rec[f.off + 1]) << 32));
  return int64_t(int32_t(rec[f.off]));
}

template <class Shape>
__device__ __forceinline__ uint64_t pa_hash_record(const PaLayout& L, const uint32_t* rec) {
  if constexpr (Shape::is_static) {
    constexpr PaLayout SL = pa_layout_of(Shape::get());
    KeyHasher kh(SL.n_keys);
    static_for<0, SL.n_keys>([&](auto K) {
      constexpr PaLayout SL = pa_layout_of(Shape::get());
      kh.add(pa_field_value(rec, SL.f[decltype(K)::value]));
    });
    return kh.finish();
  } else {
    KeyHasher kh(L.n_keys);
    for (int k = 0; k < L.n_keys; ++k) kh.add(pa_field_value(rec, L.f[k]));
    return kh.finish();
  }
}

template <class Shape>
__device__ __forceinline__ void pa_write_record(const PaLayout& L, int key_width, const V* vals, uint32_t* dst) {
  if constexpr (Shape::is_static) {
    constexpr PaLayout SL = pa_layout_of(Shape::get());
    static_for<0, SL.n_fields>([&](auto I) {
      constexpr int i = decltype(I)::value;
      constexpr PaLayout SL = pa_layout_of(Shape::get());
      int64_t v = vals[SL.f[i].expr].i;
      if (i < SL.n_keys) v = pa_key_cast(v, key_width);
      dst[SL.f[i].off] = uint32_t(uint64_t(v));
      if constexpr (SL.f[i].words == 2) dst[SL.f[i].off + 1] = uint32_t(uint64_t(v) >> 32);
    });
  } else {
    for (int i = 0; i < L.n_fields; ++i) {
      int64_t v = vals[L.f[i].expr].i;
      if (i < L.n_keys) v = pa_key_cast(v, key_width);
      dst[L.f[i].off] = uint32_t(uint64_t(v));
      if (L.f[i].words == 2) dst[L.f[i].off + 1] = uint32_t(uint64_t(v) >> 32);
    }
  }
}

template <class Shape>
struct PaVals {   // vals[] sized by the shape (registers) or by the ABI limit (local memory, interpreter)
  static constexpr int N = Shape::is_static ? (Shape::get().n_exprs > 0 ? Shape::get().n_exprs : 1) : HDK_B200_MAX_EXPRS;
  V v[N];
};

// tiles over a list of sources (fragments or level-1 regions): prefix of tiles per source, binary search per tile
__device__ __forceinline__ uint32_t pa_find_source(const uint32_t* tile_prefix, uint32_t n_src, uint64_t tile) {
  uint32_t lo = 0, hi = n_src;   // last source whose first tile is <= tile
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (tile_prefix[mid] <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- pass 1: count ------------------------------------------------------------------------------------------------
template <class Shape>
__global__ void __launch_bounds__(kPaThreads, 2) pa_count_kernel(const __grid_constant__ PaArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  uint32_t* tile_prefix = reinterpret_cast<uint32_t*>(dsm);                               // [num_fragments + 1]
  uint32_t* hist = reinterpret_cast<uint32_t*>(dsm) + ((a.num_fragments + 1 + 3) & ~3u);  // [P]
  const int tid = threadIdx.x;
  constexpr uint32_t kTile = kPaThreads * kPaCountRows;
  if (tid == 0) {
    uint32_t acc = 0;
    tile_prefix[0] = 0;
    for (uint32_t f = 0; f < a.num_fragments; ++f) {
      const int64_t rows = a.num_rows[f];
      acc += rows > 0 ? uint32_t((rows + kTile - 1) / kTile) : 0;
      tile_prefix[f + 1] = acc;
    }
  }
  for (uint32_t i = tid; i < a.P; i += kPaThreads) hist[i] = 0;
  __syncthreads();
  const uint32_t total_tiles = tile_prefix[a.num_fragments];
  PaVals<Shape> vals;
  int32_t err = 0;
  for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const uint32_t frag = pa_find_source(tile_prefix, a.num_fragments, tile);
    const uint64_t row0 = (tile - tile_prefix[frag]) * uint64_t(kTile);
    const uint64_t rows = uint64_t(a.num_rows[frag]);
    const int8_t* const* cols = a.col_buffers + size_t(frag) * a.plan.n_cols;
    const int8_t* cptr[HDK_B200_MAX_COLS];
    if constexpr (Shape::is_static) {
      constexpr DPlan sp = Shape::get();
      static_for<0, sp.n_cols>([&](auto Cc) { cptr[decltype(Cc)::value] = cols[decltype(Cc)::value]; });
    }
    if constexpr (Shape::is_static) {
      // all rows of the thread are evaluated before any is hashed: their loads are in flight together
      PaVals<Shape> v[kPaCountRows];
      bool ok[kPaCountRows];
#pragma unroll
      for (int r = 0; r < kPaCountRows; ++r) {
        const uint64_t pos = row0 + uint64_t(r) * kPaThreads + tid;
        ok[r] = pos < rows && pa_eval_row<Shape, false>(a.plan, cptr, pos < rows ? pos : row0, v[r].v, err);
      }
#pragma unroll
      for (int r = 0; r < kPaCountRows; ++r)
        if (ok[r]) atomicAdd(&hist[pa_partition(pa_hash_vals<Shape>(a.lay, a.key_width, v[r].v), a.P)], 1u);
    } else {
      for (int r = 0; r < kPaCountRows; ++r) {
        const uint64_t pos = row0 + uint64_t(r) * kPaThreads + tid;
        if (pos >= rows) continue;
        if (!pa_eval_row<Shape, false>(a.plan, cols, pos, vals.v, err)) continue;
        atomicAdd(&hist[pa_partition(pa_hash_vals<Shape>(a.lay, a.key_width, vals.v), a.P)], 1u);
      }
    }
  }
  __syncthreads();
  for (uint32_t i = tid; i < a.P; i += kPaThreads)
    if (hist[i]) atomicAdd(a.counts + i, hist[i]);
}

// ---- pass 2: exclusive scan of the partition counters (one CTA) -------------------------------------------------------
// Also the skew check: hashing spreads the GROUPS evenly, so a partition with far more rows than the others holds hot keys.
// One CTA aggregates a partition; a partition beyond `heavy_rows` would serialise the launch on one SM, so the whole
// launch falls back to the global-table probe (whose atomics spread a hot key's rows over all SMs): *skewed = 1.
__global__ void __launch_bounds__(1024) pa_offsets_kernel(const uint32_t* counts, unsigned long long* base, uint32_t P, uint32_t heavy_rows,
                                                          unsigned long long capacity_rows, int* skewed) {
  __shared__ unsigned long long warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t chunk = (P + 1023) / 1024, lo = min(uint32_t(tid) * chunk, P), hi = min(lo + chunk, P);
  unsigned long long sum = 0;
  for (uint32_t i = lo; i < hi; ++i) { sum += counts[i]; if (counts[i] > heavy_rows) *skewed = 1; }
  unsigned long long incl = sum;
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = warp_tot[lane], wi = w;
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    warp_tot[lane] = wi - w;
  }
  __syncthreads();
  unsigned long long run = warp_tot[warp] + incl - sum;
  for (uint32_t i = lo; i < hi; ++i) { base[i] = run; run += counts[i]; }
  if (hi == P && lo < P) {   // the thread owning the last counter also writes the total
    base[P] = run;
    // more rows than the caller's total_rows_hint sized the record area for: stand down as well (the fallback path does
    // not depend on the hint)
    if (run > capacity_rows) *skewed = 1;
  }
}

// ---- pass 3: scatter (level 1 from the columns, level 2 from level 1's records) ---------------------------------------
// A tile of R rows is regrouped by destination in shared memory, then written out run by run.  Every run is placed in
// the staging area at a word offset CONGRUENT (mod 4) to the global word offset it was reserved at, each in its own
// 16-byte-aligned slot: the copy-out then moves whole 16-byte vectors (LDS.128 → STG.128), word by word only at the
// two ends of a run.  Shared memory: the staging area, per vector its destination, per row its destination and its rank
// inside the destination's run, per destination the run's staging start / length and its global offset.
template <class Shape, int kLevel>
__global__ void __launch_bounds__(kPaThreads, 2) pa_scatter_kernel(const __grid_constant__ PaArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const PaLayout& L = a.lay;
  const uint32_t R = a.tile_rows;
  const uint32_t n_src = kLevel == 1 ? a.num_fragments : a.F1;
  int RW;
  if constexpr (Shape::is_static) RW = pa_layout_of(Shape::get()).rec_words; else RW = L.rec_words;
  const uint32_t stage_words = (R * uint32_t(RW) + 8u * a.n_dest + 3u) & ~3u;
  uint32_t* stage = reinterpret_cast<uint32_t*>(dsm);                                    // [stage_words]
  long long* delta = reinterpret_cast<long long*>(stage + stage_words);                  // [n_dest] global word - staging word
  uint32_t* hist = reinterpret_cast<uint32_t*>(delta + a.n_dest);                        // [n_dest]
  uint32_t* run_ws = hist + a.n_dest;                                                    // [n_dest] first staging word of the run
  uint32_t* run_len = run_ws + a.n_dest;                                                 // [n_dest] records in the run
  uint32_t* slot_ps = run_len + a.n_dest;                                                // [n_dest + 1] staging slots (multiples of 4 words)
  uint32_t* tile_prefix = slot_ps + a.n_dest + 1;                                        // [n_src + 1]
  uint32_t* row_info = tile_prefix + n_src + 1;                                          // [R]  rank | destination << 16; rank 0xffff = dropped
  uint8_t* dest_of_vec = reinterpret_cast<uint8_t*>(row_info + R);                       // [stage_words / 4]  0xff = padding
  const int tid = threadIdx.x, lane = tid & 31;
  if (*a.skewed) return;
  if (tid == 0) {
    uint32_t acc = 0;
    tile_prefix[0] = 0;
    for (uint32_t s = 0; s < n_src; ++s) {
      const int64_t rows = kLevel == 1 ? a.num_rows[s] : int64_t(a.src_count[s]);
      acc += rows > 0 ? uint32_t((rows + R - 1) / R) : 0;
      tile_prefix[s + 1] = acc;
    }
  }
  for (uint32_t i = tid; i < a.n_dest; i += kPaThreads) hist[i] = 0;
  __syncthreads();
  const uint32_t total_tiles = tile_prefix[n_src];
  const uint32_t F2_mask = (1u << a.F2_log2) - 1u;
  PaVals<Shape> vals;
  int32_t my_err = 0;
  for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const uint32_t src = pa_find_source(tile_prefix, n_src, tile);
    const uint64_t row0 = (tile - tile_prefix[src]) * uint64_t(R);
    const uint64_t src_rows = kLevel == 1 ? uint64_t(a.num_rows[src]) : uint64_t(a.src_count[src]);
    const uint32_t n_tile = uint32_t(min(uint64_t(R), src_rows - row0));
    const int8_t* const* cols = nullptr;
    const int8_t* cptr[HDK_B200_MAX_COLS];
    const uint32_t* recs = nullptr;
    if constexpr (kLevel == 1) {
      cols = a.col_buffers + size_t(src) * a.plan.n_cols;
      if constexpr (Shape::is_static) {
        constexpr DPlan sp = Shape::get();
        static_for<0, sp.n_cols>([&](auto Cc) { cptr[decltype(Cc)::value] = cols[decltype(Cc)::value]; });
      }
    } else {
      recs = a.src_recs + (a.base[size_t(src) << a.F2_log2] + row0) * uint64_t(RW);
    }
    // ---- A: destination of every row and its rank inside the destination's run (kBatch rows per thread at a time: their
    //         loads are in flight together)
    constexpr int kBatch = 4;
    for (uint32_t i0 = 0; i0 < n_tile; i0 += kPaThreads * kBatch) {
      uint64_t h[kBatch];
      bool ok[kBatch];
      if constexpr (kLevel == 1 && Shape::is_static) {
        PaVals<Shape> v[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const uint32_t i = i0 + uint32_t(b) * kPaThreads + tid;
          int32_t e = 0;
          ok[b] = i < n_tile && pa_eval_row<Shape, false>(a.plan, cptr, row0 + (i < n_tile ? i : 0), v[b].v, e);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) h[b] = pa_hash_vals<Shape>(L, a.key_width, v[b].v);
      } else if constexpr (kLevel == 1) {
        for (int b = 0; b < kBatch; ++b) {
          const uint32_t i = i0 + uint32_t(b) * kPaThreads + tid;
          int32_t e = 0;
          ok[b] = i < n_tile && pa_eval_row<Shape, false>(a.plan, cols, row0 + i, vals.v, e);
          h[b] = ok[b] ? pa_hash_vals<Shape>(L, a.key_width, vals.v) : 0;
        }
      } else {
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const uint32_t i = i0 + uint32_t(b) * kPaThreads + tid;
          ok[b] = i < n_tile;
          h[b] = pa_hash_record<Shape>(L, recs + size_t(ok[b] ? i : 0) * RW);
        }
      }
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const uint32_t i = i0 + uint32_t(b) * kPaThreads + tid;
        if (i >= n_tile) continue;
        uint32_t d = 0, rank = 0xffffu;
        if (ok[b]) {
          const uint32_t p = pa_partition(h[b], a.P);
          d = kLevel == 1 ? (p >> a.F2_log2) : (p & F2_mask);
          rank = atomicAdd(&hist[d], 1u);
        }
        row_info[i] = rank | (d << 16);
      }
    }
    __syncthreads();
    // ---- one reservation per destination; the run's staging slot holds it at the same offset (mod 4 words) as in global memory
    for (uint32_t d = tid; d < a.n_dest; d += kPaThreads) {
      const uint32_t n = hist[d];
      unsigned long long gword = 0;
      if (n) {
        // level 1 writes into the region of its destination's first final partition; level 2 into the final partition
        const uint32_t cidx = kLevel == 1 ? d : ((src << a.F2_log2) | d);
        const uint32_t pidx = kLevel == 1 ? (d << a.F2_log2) : cidx;
        gword = (a.base[pidx] + atomicAdd(a.cursor + cidx, n)) * uint64_t(RW);
      }
      run_len[d] = n;
      run_ws[d] = uint32_t(gword & 3u);                       // offset inside the slot, completed below
      delta[d] = static_cast<long long>(gword);
      hist[d] = 0;
    }
    for (uint32_t q = tid; q < stage_words / 16; q += kPaThreads) reinterpret_cast<uint32_t*>(dest_of_vec)[q] = 0xffffffffu;
    __syncthreads();
    if (tid < 32) {   // exclusive scan of the slot sizes: one warp, contiguous chunks
      const uint32_t chunk = (a.n_dest + 31) / 32, lo = min(uint32_t(lane) * chunk, a.n_dest), hi = min(lo + chunk, a.n_dest);
      uint32_t sum = 0;
      for (uint32_t i = lo; i < hi; ++i) sum += run_len[i] ? ((run_len[i] * uint32_t(RW) + run_ws[i] + 3u) & ~3u) : 0u;
      uint32_t incl = sum;
      for (int dd = 1; dd < 32; dd <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, dd);
        if (lane >= dd) incl += o;
      }
      uint32_t run = incl - sum;
      for (uint32_t i = lo; i < hi; ++i) {
        slot_ps[i] = run;
        run += run_len[i] ? ((run_len[i] * uint32_t(RW) + run_ws[i] + 3u) & ~3u) : 0u;
      }
      if (lane == 31) slot_ps[a.n_dest] = incl;
    }
    __syncthreads();
    for (uint32_t d = tid; d < a.n_dest; d += kPaThreads) {
      const uint32_t ws = slot_ps[d] + run_ws[d];
      run_ws[d] = ws;
      delta[d] -= static_cast<long long>(ws);
    }
    __syncthreads();
    // ---- B: records into their regrouped position
    for (uint32_t i0 = 0; i0 < n_tile; i0 += kPaThreads * kBatch) {
      uint32_t* dst[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const uint32_t i = i0 + uint32_t(b) * kPaThreads + tid;
        dst[b] = nullptr;
        if (i < n_tile) {
          const uint32_t info = row_info[i], rank = info & 0xffffu;
          if (rank != 0xffffu) {
            const uint32_t d = info >> 16;
            const uint32_t w = run_ws[d] + rank * uint32_t(RW);
            dst[b] = stage + w;
            const uint32_t q0 = w >> 2, q1 = (w + uint32_t(RW) - 1u) >> 2;
            dest_of_vec[q0] = uint8_t(d);
            dest_of_vec[q1] = uint8_t(d);
            for (uint32_t q = q0 + 1; q < q1; ++q) dest_of_vec[q] = uint8_t(d);
          }
        }
      }
      if constexpr (kLevel == 1 && Shape::is_static) {
        PaVals<Shape> v[kBatch];
        bool ok[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const uint32_t i = i0 + uint32_t(b) * kPaThreads + tid;
          ok[b] = dst[b] != nullptr && pa_eval_row<Shape, true>(a.plan, cptr, row0 + (dst[b] ? i : 0), v[b].v, my_err);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          if (!dst[b]) continue;
          if (ok[b]) pa_write_record<Shape>(L, a.key_width, v[b].v, dst[b]);
          else for (int w = 0; w < RW; ++w) dst[b][w] = 0;
        }
      } else if constexpr (kLevel == 1) {
        for (int b = 0; b < kBatch; ++b) {
          if (!dst[b]) continue;
          const uint32_t i = i0 + uint32_t(b) * kPaThreads + tid;
          // (an error in a key / aggregate argument of a row that passes is reported through my_err and fails the query;
          //  its slot's words stay defined)
          if (pa_eval_row<Shape, true>(a.plan, cols, row0 + i, vals.v, my_err)) pa_write_record<Shape>(L, a.key_width, vals.v, dst[b]);
          else for (int w = 0; w < RW; ++w) dst[b][w] = 0;
        }
      } else if constexpr (Shape::is_static) {
        constexpr int SRW = pa_layout_of(Shape::get()).rec_words;
        uint32_t wv[kBatch][SRW];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const uint32_t* rec = recs + size_t(dst[b] ? i0 + uint32_t(b) * kPaThreads + tid : 0) * SRW;
#pragma unroll
          for (int w = 0; w < SRW; ++w) wv[b][w] = rec[w];
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          if (!dst[b]) continue;
#pragma unroll
          for (int w = 0; w < SRW; ++w) dst[b][w] = wv[b][w];
        }
      } else {
        for (int b = 0; b < kBatch; ++b) {
          if (!dst[b]) continue;
          const uint32_t* rec = recs + size_t(i0 + uint32_t(b) * kPaThreads + tid) * RW;
          for (int w = 0; w < RW; ++w) dst[b][w] = rec[w];
        }
      }
    }
    __syncthreads();
    // ---- copy out: one 16-byte vector per thread and step, consecutive threads consecutive vectors
    const uint32_t n_vec = slot_ps[a.n_dest] >> 2;
    for (uint32_t q = tid; q < n_vec; q += kPaThreads) {
      const uint32_t d = dest_of_vec[q];
      if (d == 0xffu) continue;
      const uint4 v = reinterpret_cast<const uint4*>(stage)[q];
      const uint32_t w = q << 2, ws = run_ws[d], we = ws + run_len[d] * uint32_t(RW);
      uint32_t* out = a.dst_recs + (delta[d] + static_cast<long long>(w));
      if (w >= ws && w + 4 <= we) {
        *reinterpret_cast<uint4*>(out) = v;
      } else {
        if (w >= ws && w < we) out[0] = v.x;
        if (w + 1 >= ws && w + 1 < we) out[1] = v.y;
        if (w + 2 >= ws && w + 2 < we) out[2] = v.z;
        if (w + 3 >= ws && w + 3 < we) out[3] = v.w;
      }
    }
    __syncthreads();
  }
  if (kLevel == 1 && my_err) record_error(a.error_codes, my_err);
}

// ---- pass 4: per-partition aggregation in shared memory ---------------------------------------------------------------
constexpr int kAggWarps = kAggThreads / 32;
constexpr int kAggIlp = 2;                       // records per lane in flight
constexpr int kAggWarpChunk = 32 * kAggIlp;      // records per warp and cp.async chunk

struct PaAggArgs {
  DPlan plan;
  DLayout layout;
  PaLayout lay;
  uint32_t P, NB;                   // partitions, buckets (of 4 slots) of the shared table
  const uint32_t* counts;           // records actually written per partition
  const unsigned long long* base;
  const uint32_t* recs;
  unsigned int* work_counter;       // next partition to take
  unsigned long long* out_cursor;   // entries of the group-by buffer handed out so far
  int64_t* const* groupby_buf;
  int32_t* error_codes;
  const int* skewed;
  uint32_t off_chunks;              // byte offset of the warps' record chunks inside dynamic shared memory
  uint32_t chunk_words;             // words per chunk buffer (incl. alignment slack); every warp owns two
  uint32_t acc_off[kMaxAcc];        // byte offset of accumulator a's cells (ids at 0)
};

__device__ __forceinline__ void pa_store_slot(int8_t* p, int bytes, int padded, int64_t v) {
  if (padded == 8) *reinterpret_cast<int64_t*>(p) = bytes == 8 ? v : int64_t(uint32_t(v));
  else *reinterpret_cast<int32_t*>(p) = int32_t(v);
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds_volatile_v4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
// bit 7 of every byte of x that is zero, exact for the lowest such byte (a borrow can only flag bytes above a zero byte)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) { return (x - 0x01010101u) & ~x & 0x80808080u; }

// one accumulator update of one record (kind / argument type are compile-time constants for pre-compiled shapes)
__device__ __forceinline__ void pa_accumulate(const DPlan& p, const PaLayout& F, int k, const DAcc acc, const uint32_t* rec, uint8_t* cell) {
  if (acc.kind == ACC_CNT_ALL) { atomicAdd(reinterpret_cast<uint32_t*>(cell), 1u); return; }
  const int64_t v = pa_field_value(rec, F.f[F.acc_field[k]]);
  bool is_null = false;
  if (acc.arg_nullable) {
    const DExpr& te = p.exprs[acc.arg];
    is_null = te.kind == HDK_B200_FP ? __longlong_as_double(v) == fp_null_of(te.width)
                                     : (v == int_null_of(te.width) || (acc.arg_nullable == 2 && int32_t(v) == INT32_MIN));
  }
  // a COUNT(arg) cell counts the NULL rows (rare: no atomic for most rows); the emit step turns it into rows - NULLs
  if (acc.kind == ACC_CNT_NN) { if (is_null) atomicAdd(reinterpret_cast<uint32_t*>(cell), 1u); return; }
  if (is_null) return;
  switch (acc.kind) {
    case ACC_SUM_I: smem_add_i64(cell, v); break;
    case ACC_SUM_F: atomicAdd(reinterpret_cast<double*>(cell), __longlong_as_double(v)); break;
    case ACC_MIN_I: case ACC_MAX_I: bin_update_shared_atomic(acc.kind, cell, v); break;
    default: bin_update_shared_atomic(acc.kind, cell, f64_order_encode(__longlong_as_double(v))); break;   // MIN_F / MAX_F
  }
}

// The shared table: NB buckets of four 32-bit ids (one 16-byte vector), accumulator cells beside them indexed by slot
// (= 4 x bucket + position).  An id is the index of the group's representative record inside the partition, with an
// 8-bit fingerprint of the key hash in its top byte when the partition has < 2^24 records.  A record reads its bucket
// with one LDS.128, compares the four fingerprints at once, verifies a match against the representative's key words in
// global memory (L2), claims the first empty slot with one 32-bit CAS when nothing matches, and moves to the next
// bucket only when the bucket is full: at load 0.6 about one bucket probe per record, the warp's slowest record two or
// three — against seven slot probes with linear probing over single slots.
template <class Shape>
__global__ void __launch_bounds__(kAggThreads, 1) pa_aggregate_kernel(const __grid_constant__ PaAggArgs a) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint32_t* ids = reinterpret_cast<uint32_t*>(sm);
  __shared__ uint32_t s_part, s_overflow, s_ngroups, s_emitted, s_top;
  __shared__ unsigned long long s_outbase;
  __shared__ uint32_t s_stack[64][2];   // (mod, rem) sub-passes still to run for the current partition
  const DPlan& p = a.plan;
  const PaLayout& F = a.lay;
  const DLayout& L = a.layout;
  const uint32_t NB = a.NB, T = NB * 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (*a.skewed) return;
  int RW, KW;
  if constexpr (Shape::is_static) { RW = pa_layout_of(Shape::get()).rec_words; KW = pa_layout_of(Shape::get()).key_words; }
  else { RW = F.rec_words; KW = F.key_words; }
  auto reset_slot = [&](uint32_t s) {
    ids[s] = kEmptyId;
    for (int k = 0; k < p.n_acc; ++k) {
      if (p.accs[k].bytes == 4) reinterpret_cast<uint32_t*>(sm + a.acc_off[k])[s] = 0;
      else reinterpret_cast<int64_t*>(sm + a.acc_off[k])[s] = acc_identity(p.accs[k].kind);
    }
  };
  for (uint32_t s = tid; s < T; s += kAggThreads) reset_slot(s);
  int8_t* const buf = reinterpret_cast<int8_t*>(a.groupby_buf[0]);
  const uint64_t E = L.entry_count;
  uint32_t* const my_chunks = reinterpret_cast<uint32_t*>(sm + a.off_chunks) + size_t(warp) * 2 * a.chunk_words;
  const uint32_t ids_saddr = static_cast<uint32_t>(__cvta_generic_to_shared(ids));
  for (;;) {
    __syncthreads();
    if (tid == 0) s_part = atomicAdd(a.work_counter, 1u);
    __syncthreads();
    const uint32_t part = s_part;
    if (part >= a.P) break;
    const uint32_t n = a.counts[part];
    if (n == 0) continue;
    const uint64_t first_word = a.base[part] * uint64_t(RW);
    const uint32_t* rows = a.recs + first_word;
    const bool has_fp = n < (1u << 24) - 1;                 // spare bits of the id word hold a fingerprint of the key hash
    const uint32_t row_mask = has_fp ? 0x00ffffffu : 0xffffffffu;
    // this warp's slice of the partition, streamed in chunks of kAggWarpChunk records
    const uint32_t w_lo = uint32_t(uint64_t(n) * warp / kAggWarps), w_hi = uint32_t(uint64_t(n) * (warp + 1) / kAggWarps);
    const uint32_t n_chunks = (w_hi - w_lo + kAggWarpChunk - 1) / kAggWarpChunk;
    // a chunk is copied from the 16-byte aligned address at or below its first word
    auto issue_chunk = [&](uint32_t c) {
      const uint32_t r0 = w_lo + c * kAggWarpChunk;
      const uint64_t w0 = first_word + uint64_t(r0) * RW;
      const uint32_t nrec = min(uint32_t(kAggWarpChunk), w_hi - r0);
      const uint64_t wa = w0 & ~uint64_t(3);
      const uint32_t pieces = uint32_t((w0 - wa) + uint64_t(nrec) * RW + 3) / 4;
      const uint32_t* src = a.recs + wa;
      uint32_t* dst = my_chunks + (c & 1) * a.chunk_words;
      for (uint32_t i = lane; i < pieces; i += 32) cp_async16(dst + 4 * i, src + 4 * i);
    };
    if (tid == 0) { s_stack[0][0] = 1; s_stack[0][1] = 0; s_top = 1; }
    for (;;) {
      __syncthreads();
      if (s_top == 0) break;
      const uint32_t mod = s_stack[s_top - 1][0], rem = s_stack[s_top - 1][1];
      __syncthreads();
      if (tid == 0) { --s_top; s_overflow = 0; s_ngroups = 0; s_emitted = 0; }
      __syncthreads();
      // ---- insert + accumulate: every warp on its own, no CTA barrier inside (the next chunk streams in meanwhile)
      if (n_chunks) issue_chunk(0);
      cp_async_commit();
      for (uint32_t c = 0; c < n_chunks; ++c) {
        if (c + 1 < n_chunks) issue_chunk(c + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const uint32_t r0 = w_lo + c * kAggWarpChunk;
        const uint32_t nrec = min(uint32_t(kAggWarpChunk), w_hi - r0);
        const uint32_t* cbase = my_chunks + (c & 1) * a.chunk_words + uint32_t((first_word + uint64_t(r0) * RW) & 3u);
        const bool overflow_now = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(&s_overflow), 0) != 0;
        // Two records per lane, every lane runs the same number of steps: the probe loop is warp-synchronous — all lanes
        // iterate until the warp's last record has found its slot — so the warp never splits into groups that execute
        // the body at different times (measured with a per-lane loop: 10 of 32 lanes active on average).
        const uint32_t* rec[kAggIlp];
        uint32_t bucket[kAggIlp], slot[kAggIlp], mine[kAggIlp], fp[kAggIlp], probes[kAggIlp], excl[kAggIlp];
        bool act[kAggIlp], pend[kAggIlp];
#pragma unroll
        for (int u = 0; u < kAggIlp; ++u) {
          const uint32_t j = uint32_t(u) * 32 + lane;
          act[u] = j < nrec && !overflow_now;
          rec[u] = cbase + size_t(act[u] ? j : 0) * RW;
          const uint64_t h = pa_hash_record<Shape>(F, rec[u]);
          if (mod > 1 && (pa_split_bits(h) & (mod - 1)) != rem) act[u] = false;
          fp[u] = has_fp ? pa_fingerprint(h) : 0u;
          mine[u] = fp[u] | (r0 + j);
          bucket[u] = __umulhi(uint32_t(h), NB);
          slot[u] = 0;
          probes[u] = 0;
          excl[u] = 0;
          pend[u] = act[u];
        }
        bool any_pend = false;
#pragma unroll
        for (int u = 0; u < kAggIlp; ++u) any_pend = any_pend || pend[u];
        while (__any_sync(0xffffffffu, any_pend)) {
          uint32_t cand[kAggIlp];     // position (0-3) of the slot to verify, 4 = none
          uint32_t empty[kAggIlp];    // position of the first empty slot, 4 = none
          uint32_t cand_id[kAggIlp];
#pragma unroll
          for (int u = 0; u < kAggIlp; ++u) {
            cand[u] = empty[u] = 4;
            cand_id[u] = 0;
            if (!pend[u]) continue;
            const uint4 q = lds_volatile_v4(ids_saddr + bucket[u] * 16u);
            if (has_fp) {
              // the four top bytes in one word: fingerprint matches and empties found with byte arithmetic
              const uint32_t tb = __byte_perm(__byte_perm(q.x, q.y, 0x0073), __byte_perm(q.z, q.w, 0x0073), 0x5410);
              const uint32_t zm = zero_bytes((tb ^ ((fp[u] >> 24) * 0x01010101u)) | excl[u]);   // excl: 0xff in verified-unequal bytes
              const uint32_t ze = zero_bytes(~tb);
              if (zm) {
                cand[u] = (__ffs(zm) - 1) >> 3;
                cand_id[u] = cand[u] == 0 ? q.x : cand[u] == 1 ? q.y : cand[u] == 2 ? q.z : q.w;
              }
              if (ze) empty[u] = (__ffs(ze) - 1) >> 3;
            } else {
              const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int s = 3; s >= 0; --s) {
                if (w[s] == kEmptyId) empty[u] = s;
                else if (!((excl[u] >> (8 * s)) & 1u)) { cand[u] = s; cand_id[u] = w[s]; }
              }
            }
          }
          // a slot with the record's fingerprint: compare with the representative's key words (both records' loads in flight)
          if constexpr (Shape::is_static) {
            constexpr int SKW = pa_layout_of(Shape::get()).key_words;
            uint32_t rk[kAggIlp][SKW];
#pragma unroll
            for (int u = 0; u < kAggIlp; ++u) {
              const uint32_t* rep = rows + uint64_t(cand[u] < 4 ? (cand_id[u] & row_mask) : 0u) * RW;
#pragma unroll
              for (int w = 0; w < SKW; ++w) rk[u][w] = cand[u] < 4 ? __ldg(rep + w) : 0u;
            }
#pragma unroll
            for (int u = 0; u < kAggIlp; ++u) {
              if (cand[u] == 4) continue;
              bool eq = true;
#pragma unroll
              for (int w = 0; w < SKW; ++w) eq = eq && rk[u][w] == rec[u][w];
              if (eq) { pend[u] = false; slot[u] = bucket[u] * 4 + cand[u]; }
              else excl[u] |= 0xffu << (8 * cand[u]);
            }
          } else {
#pragma unroll
            for (int u = 0; u < kAggIlp; ++u) {
              if (cand[u] == 4) continue;
              const uint32_t* rep = rows + uint64_t(cand_id[u] & row_mask) * RW;
              bool eq = true;
              for (int w = 0; w < KW && eq; ++w) eq = __ldg(rep + w) == rec[u][w];
              if (eq) { pend[u] = false; slot[u] = bucket[u] * 4 + cand[u]; }
              else excl[u] |= 0xffu << (8 * cand[u]);
            }
          }
          any_pend = false;
#pragma unroll
          for (int u = 0; u < kAggIlp; ++u) {
            if (pend[u] && cand[u] == 4) {
              if (empty[u] < 4) {
                // nothing in the bucket matches: claim its first empty slot; on a lost race look at the bucket again
                // (the winner may hold this record's key)
                const uint32_t s = bucket[u] * 4 + empty[u];
                if (atomicCAS(ids + s, kEmptyId, mine[u]) == kEmptyId) { pend[u] = false; slot[u] = s; }
              } else {
                bucket[u] = bucket[u] + 1 == NB ? 0 : bucket[u] + 1;
                excl[u] = 0;
                if (++probes[u] >= NB) { pend[u] = false; act[u] = false; s_overflow = 1; }
              }
            }
            any_pend = any_pend || pend[u];
          }
        }
#pragma unroll
        for (int u = 0; u < kAggIlp; ++u) {
          if (!act[u]) continue;
          if constexpr (Shape::is_static) {
            constexpr DPlan sp = Shape::get();
            static_for<0, sp.n_acc>([&](auto A) {
              constexpr int k = decltype(A)::value;
              constexpr DPlan sp = Shape::get();
              constexpr PaLayout SL = pa_layout_of(Shape::get());
              pa_accumulate(sp, SL, k, sp.accs[k], rec[u], sm + a.acc_off[k] + size_t(slot[u]) * sp.accs[k].bytes);
            });
          } else {
            for (int k = 0; k < p.n_acc; ++k) pa_accumulate(p, F, k, p.accs[k], rec[u], sm + a.acc_off[k] + size_t(slot[u]) * p.accs[k].bytes);
          }
        }
        __syncwarp();   // every lane is done with this chunk buffer before chunk c + 2 lands in it
      }
      cp_async_wait<0>();
      __syncthreads();
      if (s_overflow) {
        // the groups of this sub-pass do not fit: forget them and split the sub-pass by two more hash bits
        for (uint32_t s = tid; s < T; s += kAggThreads) reset_slot(s);
        if (tid == 0) {
          if (mod >= (1u << 16) || s_top + 4 > 64) {
            record_error(a.error_codes, -HDK_B200_ERR_OUT_OF_SLOTS);
          } else {
            for (uint32_t j = 0; j < 4; ++j) { s_stack[s_top][0] = mod * 4; s_stack[s_top][1] = rem + j * mod; ++s_top; }
          }
        }
        continue;
      }
      // ---- emit: count the groups, reserve entries, encode them in the reference layout, reset the slots
      uint32_t mine_n = 0;
      for (uint32_t s = tid; s < T; s += kAggThreads) mine_n += ids[s] != kEmptyId;
      for (int d = 16; d; d >>= 1) mine_n += __shfl_xor_sync(0xffffffffu, mine_n, d);
      if (lane == 0 && mine_n) atomicAdd(&s_ngroups, mine_n);
      __syncthreads();
      if (tid == 0) {
        s_outbase = atomicAdd(a.out_cursor, (unsigned long long)s_ngroups);
        if (s_outbase + s_ngroups > E) record_error(a.error_codes, -HDK_B200_ERR_OUT_OF_SLOTS);   // more groups than entries
      }
      __syncthreads();
      const unsigned long long outbase = s_outbase;
      for (uint32_t s0 = 0; s0 < T; s0 += kAggThreads) {
        const uint32_t s = s0 + tid;
        const uint32_t id = s < T ? ids[s] : kEmptyId;
        const unsigned live = __ballot_sync(0xffffffffu, id != kEmptyId);
        uint32_t wbase = 0;
        if (lane == 0 && live) wbase = atomicAdd(&s_emitted, (uint32_t)__popc(live));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (id == kEmptyId) continue;
        const uint64_t e = outbase + wbase + __popc(live & ((1u << lane) - 1u));
        const uint32_t* rep = rows + uint64_t(id & row_mask) * RW;
        if (e < E) {
          int8_t* row = L.columnar ? nullptr : buf + e * L.row_bytes;
          for (int k = 0; k < F.n_keys; ++k) {
            const int64_t kv = pa_field_value(rep, F.f[k]);
            if (L.columnar) reinterpret_cast<int64_t*>(buf + size_t(k) * ((8 * E + 7) & ~uint64_t(7)))[e] = kv;
            else if (L.key_width == 4) reinterpret_cast<int32_t*>(row)[k] = int32_t(kv);
            else reinterpret_cast<int64_t*>(row)[k] = kv;
          }
          auto cell = [&](int k) -> int64_t {
            if (p.accs[k].bytes == 8) return reinterpret_cast<const int64_t*>(sm + a.acc_off[k])[s];
            const int64_t c = int64_t(reinterpret_cast<const uint32_t*>(sm + a.acc_off[k])[s]);
            // (accumulator 0 is always the group's row count, lower.cu)
            return p.accs[k].kind == ACC_CNT_NN ? int64_t(reinterpret_cast<const uint32_t*>(sm + a.acc_off[0])[s]) - c : c;
          };
          for (int si = 0; si < L.slot_count; ++si) {
            const DSlot& sl = L.slots[si];
            if (!sl.padded || sl.op == SLOT_KEY) continue;
            int8_t* dst = L.columnar ? buf + sl.col_off + e * sl.padded : row + L.key_bytes + sl.off;
            int64_t v = sl.init_val;
            if (sl.op == SLOT_COUNT) {
              v = cell(sl.acc);
            } else {
              // SUM / MIN / MAX with the reference's "skip_val" protocol (finalize.cu finalize_kernel)
              const bool any = !sl.skip_null || cell(sl.acc_cnt) != 0;
              if (any || sl.is_avg_sum) {
                const int64_t c = cell(sl.acc);
                if (sl.is_fp) {
                  const double d = sl.op == SLOT_SUM ? __longlong_as_double(c) : f64_order_decode(c);
                  v = sl.bytes == 4 ? int64_t(__float_as_uint(float(d))) : __double_as_longlong(d);
                } else {
                  v = c;
                }
                if (!any) v = 0;   // AVG over all-NULL: the sum slot stays 0 (0.0 / 0.f have all-zero bits)
              }
            }
            pa_store_slot(dst, sl.bytes, sl.padded, v);
          }
        }
        reset_slot(s);
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static size_t pa_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct PaGeometry {
  uint32_t P, F1, F2_log2, T, tile_rows;
  size_t scatter_smem[2];           // level 1 (n_src = fragments), level 2 (n_src = F1)
  size_t agg_smem;
  uint32_t off_chunks, chunk_words;
  uint32_t acc_off[kMaxAcc];
  size_t header_bytes, rec_bytes, total_bytes;
  uint32_t heavy_rows;
};

static int partagg_geometry(const Lowered& lw, const PaLayout& L, uint64_t total_rows, uint32_t num_fragments, PaGeometry* g) {
  const DPlan& p = lw.plan;
  int dev = 0, max_smem = 0;
  HB_CUDA(cudaGetDevice(&dev));
  HB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // ---- aggregation: per warp two record chunks, then the table (buckets of four ids + the accumulator cells)
  g->chunk_words = uint32_t(pa_align(size_t(kAggWarpChunk) * L.rec_words + 8, 4));
  const size_t chunk_bytes = size_t(g->chunk_words) * 4 * 2 * kAggWarps;
  size_t per_slot = 4;
  for (int a = 0; a < p.n_acc; ++a) per_slot += p.accs[a].bytes;
  const size_t budget = size_t(max_smem) - 2048;
  if (chunk_bytes + 128 * per_slot + 64 * size_t(p.n_acc + 2) > budget) { set_error("partitioned aggregation: records too wide for shared memory"); return HDK_B200_E_UNSUPPORTED; }
  uint32_t T = uint32_t(std::min<size_t>((budget - chunk_bytes - 64 * size_t(p.n_acc + 2)) / per_slot, 1u << 20));
  T &= ~127u;
  if (g_debug.pa_slots > 0) T = std::min<uint32_t>(T, std::max<uint32_t>(128u, uint32_t(g_debug.pa_slots) & ~127u));
  if (T < 128) { set_error("partitioned aggregation: accumulators do not fit in shared memory"); return HDK_B200_E_UNSUPPORTED; }
  g->T = T;
  size_t off = pa_align(size_t(T) * 4, 16);
  for (int a = 0; a < p.n_acc; ++a) { g->acc_off[a] = uint32_t(off); off = pa_align(off + size_t(T) * p.accs[a].bytes, 16); }
  g->off_chunks = uint32_t(off);
  g->agg_smem = off + chunk_bytes;
  // ---- partitions.  The caller sizes the table at 2 x the estimated number of groups (QE/RelAlgExecutor.cpp:1553-1557):
  // aim at 60 % full shared tables for that estimate; a partition that still overflows is split by further hash bits
  const uint64_t groups_est = std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(p.entry_count) / 2 + 1, total_rows));
  const uint64_t per_part = std::max<uint64_t>(1, uint64_t(T) * 6 / 10);   // buckets of four tolerate a load of 0.6
  uint64_t P = (groups_est + per_part - 1) / per_part;
  // enough partitions to occupy the GPU even with few groups, as long as they keep a few thousand rows each
  P = std::max<uint64_t>(P, std::min<uint64_t>(uint64_t(sm_count()) * 8, total_rows / 4096));
  if (g_debug.pa_partitions > 0) P = uint64_t(g_debug.pa_partitions);
  P = std::max<uint64_t>(1, std::min<uint64_t>(P, kPaMaxPartitions));
  uint32_t f2 = 0;
  while (((P + (uint64_t(1) << f2) - 1) >> f2) > kPaFanout) ++f2;   // F2 = 2^f2 final partitions per level-1 destination; F1 = ceil(P / F2) <= 255
  const uint32_t F1 = uint32_t((P + (uint64_t(1) << f2) - 1) >> f2);
  g->F1 = F1;
  g->F2_log2 = f2;
  g->P = F1 << f2;
  // ---- scatter tiles: R rows x (record + rank + destination) + per destination (slot padding, descriptors) + per vector a byte
  auto scatter_bytes = [&](size_t R, size_t n_src, size_t n_dest) {
    const size_t stage_words = (R * L.rec_words + 8 * n_dest + 3) & ~size_t(3);
    return stage_words * 4 + n_dest * (8 + 4 + 4 + 4 + 4) + 4 + (n_src + 1) * 4 + R * 4 + stage_words / 4 + 64;
  };
  g->tile_rows = 0;
  for (size_t R = kPaMaxTileRows; R >= size_t(kPaThreads); R -= kPaThreads) {
    const size_t need = std::max(scatter_bytes(R, num_fragments, F1), scatter_bytes(R, F1, size_t(1) << f2));
    // two CTAs per SM when the tile stays at least half the maximum, else one
    if (need + 1024 <= size_t(max_smem) / 2 || (R <= kPaMaxTileRows / 2 && need + 1024 <= size_t(max_smem))) { g->tile_rows = uint32_t(R); break; }
  }
  if (!g->tile_rows) { set_error("partitioned aggregation: records too wide for a scatter tile"); return HDK_B200_E_UNSUPPORTED; }
  g->scatter_smem[0] = scatter_bytes(g->tile_rows, num_fragments, F1);
  g->scatter_smem[1] = scatter_bytes(g->tile_rows, F1, size_t(1) << f2);
  // header: counts u32[P] | cursor2 u32[P] | cursor1 u32[F1] | base u64[P + 1] | work counter, out cursor
  g->header_bytes = pa_align(size_t(g->P) * 8 + size_t(F1) * 4 + 8 + (size_t(g->P) + 1) * 8 + 64, 256);
  g->rec_bytes = pa_align(size_t(total_rows) * L.rec_words * 4 + 64, 256);
  // (the fallback's work table lives where the records would: only one of the two paths runs)
  g->total_bytes = g->header_bytes + std::max(g->rec_bytes * (f2 ? 2 : 1), pa_align(lw.work_table_bytes, 256)) + 256;
  // heavy partition: more rows than one SM can take without stretching the launch (a quarter of an even share of the
  // SMs' work) and far beyond the average
  const uint64_t avg = total_rows / g->P + 1;
  uint64_t heavy = std::max<uint64_t>(8 * avg + 65536, total_rows / (uint64_t(sm_count()) * 4));
  if (g_debug.pa_heavy_rows > 0) heavy = uint64_t(g_debug.pa_heavy_rows);
  g->heavy_rows = uint32_t(std::min<uint64_t>(heavy, 0xffffffffu));
  return HDK_B200_OK;
}

int partagg_scratch_bytes(const Lowered& lw, uint64_t total_rows, size_t* bytes) {
  const PaLayout L = pa_layout_of(lw.plan);
  if (!L.ok) { set_error("plan shape not eligible for partitioned aggregation"); return HDK_B200_E_UNSUPPORTED; }
  if (total_rows >= 0xfffffff0ull) { set_error("partitioned aggregation: too many rows per launch"); return HDK_B200_E_UNSUPPORTED; }
  PaGeometry g;
  if (int rc = partagg_geometry(lw, L, total_rows, 1, &g)) return rc;
  *bytes = g.total_bytes;
  return HDK_B200_OK;
}

// kernels of one plan shape
struct PaKernels {
  uint64_t sig;
  void (*count)(const PaArgs);
  void (*scatter1)(const PaArgs);
  void (*scatter2)(const PaArgs);
  void (*aggregate)(const PaAggArgs);
};
template <class Shape>
constexpr PaKernels pa_kernels_of(uint64_t sig) {
  return PaKernels{sig, pa_count_kernel<Shape>, pa_scatter_kernel<Shape, 1>, pa_scatter_kernel<Shape, 2>, pa_aggregate_kernel<Shape>};
}
template <int ID>
constexpr PaKernels pa_static_kernels(uint64_t sig) {
  if constexpr (pa_layout_of(StaticShape<ID>::get()).ok != 0) return pa_kernels_of<StaticShape<ID>>(sig);
  else return PaKernels{0, nullptr, nullptr, nullptr, nullptr};
}
#define HB_STATIC_SHAPE(ID, SIG, NAME, RPI, ...) pa_static_kernels<ID>(SIG),
static const PaKernels kPaStatic[] = {
#include "static_shapes.inc"
    PaKernels{0, nullptr, nullptr, nullptr, nullptr}};
#undef HB_STATIC_SHAPE

int launch_partagg(const Lowered& lw, const hdk_b200_kernel_params* params, void* scratch, size_t scratch_bytes, cudaStream_t st,
                   hdk_b200_launch_info* info, const int** fallback_flag, int64_t** fallback_work) {
  const PaLayout L = pa_layout_of(lw.plan);
  if (!L.ok) { set_error("plan shape not eligible for partitioned aggregation"); return HDK_B200_E_UNSUPPORTED; }
  const uint64_t total_rows = params->total_rows_hint;
  if (params->num_fragments > kPaMaxFragments) { set_error("more than %u fragments per launch", kPaMaxFragments); return HDK_B200_E_UNSUPPORTED; }
  PaGeometry g;
  if (int rc = partagg_geometry(lw, L, total_rows, uint32_t(params->num_fragments), &g)) return rc;
  if (scratch_bytes < g.total_bytes || !scratch) { set_error("partitioned aggregation needs %zu scratch bytes, got %zu", g.total_bytes, scratch_bytes); return HDK_B200_E_INVALID; }
  PaKernels kern = pa_kernels_of<GenericShape>(0);
  int variant = 0;
  if (!g_debug.force_generic) {
    const uint64_t sig = plan_signature(lw.plan);
    for (int i = 0; i < int(sizeof(kPaStatic) / sizeof(kPaStatic[0])); ++i)
      if (kPaStatic[i].count && kPaStatic[i].sig == sig) { kern = kPaStatic[i]; variant = i + 1; break; }
  }
  uint8_t* s = static_cast<uint8_t*>(scratch);
  const bool two_level = g.F2_log2 != 0;
  PaArgs a{};
  a.plan = lw.plan;
  a.lay = L;
  a.key_width = lw.layout.key_width;
  a.col_buffers = params->col_buffers;
  a.num_rows = params->num_rows;
  a.num_fragments = uint32_t(params->num_fragments);
  a.P = g.P;
  a.F1 = g.F1;
  a.F2_log2 = g.F2_log2;
  a.tile_rows = g.tile_rows;
  a.counts = reinterpret_cast<uint32_t*>(s);
  uint32_t* cursor2 = a.counts + g.P;
  uint32_t* cursor1 = cursor2 + g.P;
  unsigned long long* base = reinterpret_cast<unsigned long long*>(s + pa_align(size_t(g.P) * 8 + size_t(g.F1) * 4, 8));
  a.base = base;
  unsigned int* work_counter = reinterpret_cast<unsigned int*>(base + g.P + 1);
  unsigned long long* out_cursor = reinterpret_cast<unsigned long long*>(work_counter + 2);
  int* skewed = reinterpret_cast<int*>(out_cursor + 1);
  a.skewed = skewed;
  *fallback_flag = skewed;
  *fallback_work = reinterpret_cast<int64_t*>(s + g.header_bytes);
  uint32_t* recs_a = reinterpret_cast<uint32_t*>(s + g.header_bytes);               // level 1 output
  uint32_t* recs_b = reinterpret_cast<uint32_t*>(s + g.header_bytes + g.rec_bytes);   // level 2 output
  a.error_codes = params->error_codes;
  HB_CUDA(cudaMemsetAsync(s, 0, g.header_bytes, st));
  const size_t count_smem = pa_align((size_t(a.num_fragments) + 1) * 4, 16) + size_t(g.P) * 4;
  HB_CUDA(cudaFuncSetAttribute(kern.count, cudaFuncAttributeMaxDynamicSharedMemorySize, int(count_smem)));
  kern.count<<<sm_count() * (count_smem > 100 * 1024 ? 1 : 2), kPaThreads, count_smem, st>>>(a);
  HB_LAUNCH_CHECK();
  pa_offsets_kernel<<<1, 1024, 0, st>>>(a.counts, base, g.P, g.heavy_rows, total_rows, skewed);
  HB_LAUNCH_CHECK();
  // level 1: fragments → F1 destinations (the final partitions themselves when one level suffices)
  a.n_dest = g.F1;
  a.cursor = two_level ? cursor1 : cursor2;
  a.dst_recs = recs_a;
  HB_CUDA(cudaFuncSetAttribute(kern.scatter1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(g.scatter_smem[0])));
  kern.scatter1<<<sm_count() * 2, kPaThreads, g.scatter_smem[0], st>>>(a);
  HB_LAUNCH_CHECK();
  int launches = 3;
  if (two_level) {
    a.src_recs = recs_a;
    a.src_count = cursor1;
    a.n_dest = 1u << g.F2_log2;
    a.cursor = cursor2;
    a.dst_recs = recs_b;
    HB_CUDA(cudaFuncSetAttribute(kern.scatter2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(g.scatter_smem[1])));
    kern.scatter2<<<sm_count() * 2, kPaThreads, g.scatter_smem[1], st>>>(a);
    HB_LAUNCH_CHECK();
    ++launches;
  }
  PaAggArgs ag{};
  ag.plan = lw.plan;
  ag.layout = lw.layout;
  ag.lay = L;
  ag.P = g.P;
  ag.NB = g.T / 4;
  ag.counts = cursor2;
  ag.base = base;
  ag.recs = two_level ? recs_b : recs_a;
  ag.work_counter = work_counter;
  ag.out_cursor = out_cursor;
  ag.groupby_buf = params->groupby_buf;
  ag.error_codes = params->error_codes;
  ag.skewed = skewed;
  ag.off_chunks = g.off_chunks;
  ag.chunk_words = g.chunk_words;
  for (int k = 0; k < lw.plan.n_acc; ++k) ag.acc_off[k] = g.acc_off[k];
  HB_CUDA(cudaFuncSetAttribute(kern.aggregate, cudaFuncAttributeMaxDynamicSharedMemorySize, int(g.agg_smem)));
  kern.aggregate<<<int(std::min<uint32_t>(g.P, uint32_t(sm_count()))), kAggThreads, g.agg_smem, st>>>(ag);
  HB_LAUNCH_CHECK();
  ++launches;
  if (info) {
    info->variant = variant;
    info->strategy = HDK_B200_STRATEGY_PARTITIONED;
    info->n_launches = launches;
    info->grid = sm_count() * 2;
    info->block = kAggThreads;
    info->smem_bytes = int(g.agg_smem);
    info->n_accumulators = lw.plan.n_acc;
    info->tile_rows = int(g.P);
  }
  return HDK_B200_OK;
}

}  // namespace hb
