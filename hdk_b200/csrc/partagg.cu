// hdk_b200/csrc/partagg.cu — radix-partitioned baseline-hash aggregation (high-cardinality group-by).
//
// The reference aggregates a baseline-hash group-by by probing ONE open-addressing table per kernel
// (get_group_value, QE/GroupByRuntime.cpp:31-54; GPU claim protocol QE/cuda_mapd_rt.cu:176-236) — on a GPU that is a
// random DRAM sector plus several global atomics per row once the table outgrows L2 (config 4: 2e8 entries).  Its own
// CPU answer to that shape is partitioned aggregation (QE/RelAlgExecutor.cpp:691-838: partition the rows by a hash of
// the key — QE/RowFuncBuilder.cpp:516-577 — so that every partition aggregates on its own).  Here the same idea runs
// INSIDE one GPU, down to partitions whose groups fit a CTA's shared memory:
//
//   1. count      every row's key → partition = key hash range-reduced to P = F1 x F2 (<= 32768); per-CTA shared histogram
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

constexpr uint32_t kPaMaxPartitions = 32768;
constexpr uint32_t kPaMaxFragments = 4096;
constexpr uint32_t kPaFanout = 256;          // destinations of one scatter pass
constexpr int kPaThreads = 512;
constexpr int kPaMaxTileRows = 4096;         // rows regrouped per tile (ranks fit 16 bits, destinations 8)
constexpr int kPaCountRows = 8;              // rows per thread per tile of the count pass
constexpr int kAggThreads = 1024;
constexpr int kAggChunkRows = 2048;          // records staged per cp.async chunk
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
};

// MurmurHash64A over the 64-bit widened key values (the reference's partition hash, QE/RowFuncBuilder.cpp:516-577),
// computed incrementally so that the keys need not sit in an array
struct KeyHasher {
  uint64_t h;
  __device__ __forceinline__ explicit KeyHasher(int n_keys) : h(uint64_t(n_keys) * 8 * 0xc6a4a7935bd1e995ULL) {}
  __device__ __forceinline__ void add(int64_t key) {
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    uint64_t k = uint64_t(key) * m;
    k ^= k >> 47;
    k *= m;
    h ^= k;
    h *= m;
  }
  __device__ __forceinline__ uint64_t finish() const {
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    uint64_t x = h;
    x ^= x >> 47;
    x *= m;
    x ^= x >> 47;
    return x;
  }
};
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
#pragma unroll
    for (int r = 0; r < kPaCountRows; ++r) {
      const uint64_t pos = row0 + uint64_t(r) * kPaThreads + tid;
      if (pos >= rows) continue;
      if (!pa_eval_row<Shape, false>(a.plan, Shape::is_static ? cptr : cols, pos, vals.v, err)) continue;
      atomicAdd(&hist[pa_partition(pa_hash_vals<Shape>(a.lay, a.key_width, vals.v), a.P)], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = tid; i < a.P; i += kPaThreads)
    if (hist[i]) atomicAdd(a.counts + i, hist[i]);
}

// ---- pass 2: exclusive scan of the partition counters (one CTA) -------------------------------------------------------
__global__ void __launch_bounds__(1024) pa_offsets_kernel(const uint32_t* counts, unsigned long long* base, uint32_t P) {
  __shared__ unsigned long long warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t chunk = (P + 1023) / 1024, lo = min(uint32_t(tid) * chunk, P), hi = min(lo + chunk, P);
  unsigned long long sum = 0;
  for (uint32_t i = lo; i < hi; ++i) sum += counts[i];
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
  if (hi == P && lo < P) base[P] = run;   // the thread owning the last counter also writes the total
}

// ---- pass 3: scatter (level 1 from the columns, level 2 from level 1's records) ---------------------------------------
// Shared memory of a tile of R rows: the regrouped records (R x rec_words words), per row its destination and its rank
// inside the destination's run, per regrouped position its destination; per destination the run start, the reservation.
template <class Shape, int kLevel>
__global__ void __launch_bounds__(kPaThreads, 2) pa_scatter_kernel(const __grid_constant__ PaArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const PaLayout& L = a.lay;
  const uint32_t R = a.tile_rows;
  const uint32_t n_src = kLevel == 1 ? a.num_fragments : a.F1;
  int RW;
  if constexpr (Shape::is_static) RW = pa_layout_of(Shape::get()).rec_words; else RW = L.rec_words;
  uint32_t* stage = reinterpret_cast<uint32_t*>(dsm);                                    // [R * RW]
  long long* delta = reinterpret_cast<long long*>(stage + ((size_t(R) * RW + 3) & ~size_t(3)));   // [n_dest] words: global - local
  uint32_t* hist = reinterpret_cast<uint32_t*>(delta + a.n_dest);                        // [n_dest]
  uint32_t* run_start = hist + a.n_dest;                                                 // [n_dest + 1]
  uint32_t* tile_prefix = run_start + a.n_dest + 1;                                      // [n_src + 1]
  uint16_t* rank_of_row = reinterpret_cast<uint16_t*>(tile_prefix + n_src + 1);          // [R]  0xffff = dropped
  uint8_t* dest_of_row = reinterpret_cast<uint8_t*>(rank_of_row + R);                    // [R]
  uint8_t* dest_of_pos = dest_of_row + R;                                                // [R]
  const int tid = threadIdx.x, lane = tid & 31;
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
    // ---- A: destination of every row and its rank inside the destination's run
    for (uint32_t i = tid; i < n_tile; i += kPaThreads) {
      uint64_t h;
      bool ok = true;
      if constexpr (kLevel == 1) {
        int32_t e = 0;
        ok = pa_eval_row<Shape, false>(a.plan, Shape::is_static ? cptr : cols, row0 + i, vals.v, e);
        h = ok ? pa_hash_vals<Shape>(L, a.key_width, vals.v) : 0;
      } else {
        h = pa_hash_record<Shape>(L, recs + size_t(i) * RW);
      }
      uint32_t d = 0;
      uint32_t rank = 0xffffu;
      if (ok) {
        const uint32_t p = pa_partition(h, a.P);
        d = kLevel == 1 ? (p >> a.F2_log2) : (p & F2_mask);
        rank = atomicAdd(&hist[d], 1u);
      }
      dest_of_row[i] = uint8_t(d);
      rank_of_row[i] = uint16_t(rank);
    }
    __syncthreads();
    // ---- run starts (exclusive scan of the histogram: one warp, contiguous chunks) and one reservation per destination
    if (tid < 32) {
      const uint32_t chunk = (a.n_dest + 31) / 32, lo = min(uint32_t(lane) * chunk, a.n_dest), hi = min(lo + chunk, a.n_dest);
      uint32_t sum = 0;
      for (uint32_t i = lo; i < hi; ++i) sum += hist[i];
      uint32_t incl = sum;
      for (int dd = 1; dd < 32; dd <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, dd);
        if (lane >= dd) incl += o;
      }
      uint32_t run = incl - sum;
      for (uint32_t i = lo; i < hi; ++i) { run_start[i] = run; run += hist[i]; }
      if (lane == 31) run_start[a.n_dest] = incl;
    }
    __syncthreads();
    for (uint32_t d = tid; d < a.n_dest; d += kPaThreads) {
      const uint32_t n = hist[d];
      if (n) {
        // level 1 writes into the region of its destination's first final partition; level 2 into the final partition
        const uint32_t cidx = kLevel == 1 ? d : ((src << a.F2_log2) | d);
        const uint32_t pidx = kLevel == 1 ? (d << a.F2_log2) : cidx;
        const unsigned long long g = a.base[pidx] + atomicAdd(a.cursor + cidx, n);
        delta[d] = (static_cast<long long>(g) - static_cast<long long>(run_start[d])) * RW;
      }
      hist[d] = 0;
    }
    // ---- B: records into their regrouped position
    for (uint32_t i = tid; i < n_tile; i += kPaThreads) {
      const uint32_t rank = rank_of_row[i];
      if (rank == 0xffffu) continue;
      const uint32_t d = dest_of_row[i];
      const uint32_t pos = run_start[d] + rank;
      uint32_t* dst = stage + size_t(pos) * RW;
      if constexpr (kLevel == 1) {
        if (pa_eval_row<Shape, true>(a.plan, Shape::is_static ? cptr : cols, row0 + i, vals.v, my_err)) {
          pa_write_record<Shape>(L, a.key_width, vals.v, dst);
        } else {
          // an error in a key / aggregate argument of a row that passes: reported (my_err), the query fails; keep the
          // slot's words defined
          for (int w = 0; w < RW; ++w) dst[w] = 0;
        }
      } else {
        const uint32_t* rec = recs + size_t(i) * RW;
        if constexpr (Shape::is_static) {
          constexpr int SRW = pa_layout_of(Shape::get()).rec_words;
#pragma unroll
          for (int w = 0; w < SRW; ++w) dst[w] = rec[w];
        } else {
          for (int w = 0; w < RW; ++w) dst[w] = rec[w];
        }
      }
      dest_of_pos[pos] = uint8_t(d);
    }
    __syncthreads();
    // ---- copy the runs out: consecutive threads, consecutive words
    const uint32_t n_words = run_start[a.n_dest] * uint32_t(RW);
    for (uint32_t w = tid; w < n_words; w += kPaThreads) {
      uint32_t i;
      if constexpr (Shape::is_static) i = w / uint32_t(pa_layout_of(Shape::get()).rec_words); else i = w / uint32_t(RW);
      a.dst_recs[delta[dest_of_pos[i]] + static_cast<long long>(w)] = stage[w];
    }
    __syncthreads();
  }
  if (kLevel == 1 && my_err) record_error(a.error_codes, my_err);
}

// ---- pass 4: per-partition aggregation in shared memory ---------------------------------------------------------------
struct PaAggArgs {
  DPlan plan;
  DLayout layout;
  PaLayout lay;
  uint32_t P, T;                    // partitions, shared-table slots
  const uint32_t* counts;           // records actually written per partition
  const unsigned long long* base;
  const uint32_t* recs;
  unsigned int* work_counter;       // next partition to take
  unsigned long long* out_cursor;   // entries of the group-by buffer handed out so far
  int64_t* const* groupby_buf;
  int32_t* error_codes;
  uint32_t off_chunks;              // byte offset of the two record chunks inside dynamic shared memory
  uint32_t chunk_words;             // words per chunk buffer (incl. alignment slack)
  uint32_t acc_off[kMaxAcc];        // byte offset of accumulator a's cells (ids at 0)
};

// 64-bit integer SUM in shared memory with native 32-bit atomics: add the low half, carry into the high half.  Every
// carry is added exactly once by the thread whose addition produced it, additions commute, so the cell ends up exact
// modulo 2^64 whatever the interleaving (a 64-bit shared atomicAdd is a CAS loop: SASS ATOMS.CAST.SPIN.64).
__device__ __forceinline__ void smem_add_i64(uint8_t* cell, int64_t x) {
  uint32_t* w = reinterpret_cast<uint32_t*>(cell);
  const uint32_t lo = uint32_t(uint64_t(x)), hi = uint32_t(uint64_t(x) >> 32);
  uint32_t carry = 0;
  if (lo) {
    const uint32_t old = atomicAdd(w, lo);
    carry = (old + lo) < old ? 1u : 0u;
  }
  const uint32_t h = hi + carry;
  if (h) atomicAdd(w + 1, h);
}

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

// one accumulator update of one record (kind / argument type are compile-time constants for pre-compiled shapes)
__device__ __forceinline__ void pa_accumulate(const DPlan& p, const PaLayout& F, int k, const DAcc acc, const uint32_t* rec, uint8_t* cell) {
  if (acc.kind == ACC_CNT_ALL) { atomicAdd(reinterpret_cast<uint32_t*>(cell), 1u); return; }
  const int64_t v = pa_field_value(rec, F.f[F.acc_field[k]]);
  if (acc.arg_nullable) {
    const DExpr& te = p.exprs[acc.arg];
    const bool is_null = te.kind == HDK_B200_FP ? __longlong_as_double(v) == fp_null_of(te.width)
                                                : (v == int_null_of(te.width) || (acc.arg_nullable == 2 && int32_t(v) == INT32_MIN));
    if (is_null) return;
  }
  switch (acc.kind) {
    case ACC_CNT_NN: atomicAdd(reinterpret_cast<uint32_t*>(cell), 1u); break;
    case ACC_SUM_I: smem_add_i64(cell, v); break;
    case ACC_SUM_F: atomicAdd(reinterpret_cast<double*>(cell), __longlong_as_double(v)); break;
    case ACC_MIN_I: case ACC_MAX_I: bin_update_shared_atomic(acc.kind, cell, v); break;
    default: bin_update_shared_atomic(acc.kind, cell, f64_order_encode(__longlong_as_double(v))); break;   // MIN_F / MAX_F
  }
}

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
  const uint32_t T = a.T;
  const int tid = threadIdx.x, lane = tid & 31;
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
  uint32_t* chunk_buf[2] = {reinterpret_cast<uint32_t*>(sm + a.off_chunks), reinterpret_cast<uint32_t*>(sm + a.off_chunks) + a.chunk_words};
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
    const int row_bits = n < (1u << 24) - 1 ? 24 : 32;      // spare bits of the id word hold a fingerprint of the key hash
    const uint32_t row_mask = row_bits == 32 ? 0xffffffffu : (1u << row_bits) - 1u;
    const uint32_t n_chunks = (n + kAggChunkRows - 1) / kAggChunkRows;
    // chunk c = records [c * kAggChunkRows, ...): copied from the 16-byte aligned address at or below its first word
    auto issue_chunk = [&](uint32_t c) {
      const uint64_t w0 = first_word + uint64_t(c) * kAggChunkRows * RW;
      const uint32_t nrec = min(uint32_t(kAggChunkRows), n - c * kAggChunkRows);
      const uint64_t wa = w0 & ~uint64_t(3);
      const uint32_t pieces = uint32_t((w0 - wa) + uint64_t(nrec) * RW + 3) / 4;
      const uint32_t* src = a.recs + wa;
      uint32_t* dst = chunk_buf[c & 1];
      for (uint32_t i = tid; i < pieces; i += kAggThreads) cp_async16(dst + 4 * i, src + 4 * i);
    };
    if (tid == 0) { s_stack[0][0] = 1; s_stack[0][1] = 0; s_top = 1; }
    for (;;) {
      __syncthreads();
      if (s_top == 0) break;
      const uint32_t mod = s_stack[s_top - 1][0], rem = s_stack[s_top - 1][1];
      __syncthreads();
      if (tid == 0) { --s_top; s_overflow = 0; s_ngroups = 0; s_emitted = 0; }
      issue_chunk(0);
      cp_async_commit();
      __syncthreads();
      // ---- insert + accumulate, chunk by chunk (the next chunk streams in meanwhile)
      for (uint32_t c = 0; c < n_chunks; ++c) {
        if (c + 1 < n_chunks) issue_chunk(c + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint32_t c0 = c * kAggChunkRows;
        const uint32_t nrec = min(uint32_t(kAggChunkRows), n - c0);
        const uint32_t* cbase = chunk_buf[c & 1] + uint32_t((first_word + uint64_t(c0) * RW) & 3u);
        if (!s_overflow)
        for (uint32_t j = tid; j < nrec; j += kAggThreads) {
          const uint32_t* rec = cbase + size_t(j) * RW;
          const uint32_t i = c0 + j;
          const uint64_t h = pa_hash_record<Shape>(F, rec);
          if (mod > 1 && (uint32_t(h >> 13) & (mod - 1)) != rem) continue;
          const uint32_t fp = row_bits == 32 ? 0u : (uint32_t(h >> 45) & 0xffu) << 24;
          const uint32_t mine = fp | i;
          uint32_t slot = __umulhi(uint32_t(h), T);
          uint32_t probes = 0;
          bool lost = false;
          for (;;) {
            uint32_t id = *reinterpret_cast<volatile uint32_t*>(ids + slot);
            if (id == kEmptyId) {
              id = atomicCAS(ids + slot, kEmptyId, mine);
              if (id == kEmptyId) break;      // claimed: this record represents the group
            }
            if ((id & ~row_mask) == fp) {     // same fingerprint: compare with the representative's key words
              const uint32_t* rep = rows + uint64_t(id & row_mask) * RW;
              bool eq = true;
              if constexpr (Shape::is_static) {
                constexpr int SKW = pa_layout_of(Shape::get()).key_words;
                uint32_t rk[SKW];
#pragma unroll
                for (int w = 0; w < SKW; ++w) rk[w] = __ldg(rep + w);
#pragma unroll
                for (int w = 0; w < SKW; ++w) eq = eq && rk[w] == rec[w];
              } else {
                for (int w = 0; w < KW && eq; ++w) eq = __ldg(rep + w) == rec[w];
              }
              if (eq) break;
            }
            slot = slot + 1 == T ? 0 : slot + 1;
            if (++probes >= T) { lost = true; break; }
          }
          if (lost) { s_overflow = 1; break; }
          if constexpr (Shape::is_static) {
            constexpr DPlan sp = Shape::get();
            static_for<0, sp.n_acc>([&](auto A) {
              constexpr int k = decltype(A)::value;
              constexpr DPlan sp = Shape::get();
              constexpr PaLayout SL = pa_layout_of(Shape::get());
              pa_accumulate(sp, SL, k, sp.accs[k], rec, sm + a.acc_off[k] + size_t(slot) * sp.accs[k].bytes);
            });
          } else {
            for (int k = 0; k < p.n_acc; ++k) pa_accumulate(p, F, k, p.accs[k], rec, sm + a.acc_off[k] + size_t(slot) * p.accs[k].bytes);
          }
        }
        __syncthreads();   // every thread is done with this chunk buffer before chunk c + 2 lands in it
      }
      cp_async_wait<0>();
      __syncthreads();
      if (s_overflow) {
        // the groups of this sub-pass do not fit: forget them and split the sub-pass by two more hash bits
        for (uint32_t s = tid; s < T; s += kAggThreads) reset_slot(s);
        if (tid == 0) {
          if (mod >= (1u << 18) || s_top + 4 > 64) {
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
            return p.accs[k].bytes == 4 ? int64_t(reinterpret_cast<const uint32_t*>(sm + a.acc_off[k])[s])
                                        : reinterpret_cast<const int64_t*>(sm + a.acc_off[k])[s];
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
};

static int partagg_geometry(const Lowered& lw, const PaLayout& L, uint64_t total_rows, uint32_t num_fragments, PaGeometry* g) {
  const DPlan& p = lw.plan;
  int dev = 0, max_smem = 0;
  HB_CUDA(cudaGetDevice(&dev));
  HB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // ---- aggregation: two record chunks + the table
  g->chunk_words = uint32_t(pa_align(size_t(kAggChunkRows) * L.rec_words + 8, 4));
  const size_t chunk_bytes = size_t(g->chunk_words) * 4 * 2;
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
  // aim at half-full shared tables for that estimate; a partition that still overflows is split by further hash bits
  const uint64_t groups_est = std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(p.entry_count) / 2 + 1, total_rows));
  uint64_t P = (groups_est + T / 2 - 1) / (T / 2);
  // enough partitions to occupy the GPU even with few groups, as long as they keep a few thousand rows each
  P = std::max<uint64_t>(P, std::min<uint64_t>(uint64_t(sm_count()) * 8, total_rows / 4096));
  if (g_debug.pa_partitions > 0) P = uint64_t(g_debug.pa_partitions);
  P = std::max<uint64_t>(1, std::min<uint64_t>(P, kPaMaxPartitions));
  uint32_t f2 = 0;
  while ((P >> f2) > kPaFanout) ++f2;                         // F2 = 2^f2 final partitions per level-1 destination
  const uint32_t F1 = uint32_t((P + (uint64_t(1) << f2) - 1) >> f2);
  g->F1 = F1;
  g->F2_log2 = f2;
  g->P = F1 << f2;
  // ---- scatter tiles: R rows x (record + rank + 2 destination bytes)
  const size_t per_dest = 8 + 4 + 4;
  for (int lvl = 0; lvl < 2; ++lvl) {
    const size_t n_src = lvl == 0 ? num_fragments : F1;
    const size_t n_dest = lvl == 0 ? F1 : (size_t(1) << f2);
    const size_t fixed = n_dest * per_dest + 8 + (n_src + 1) * 4 + 64;
    const size_t room = (size_t(max_smem) / 2 > fixed + 4096 ? size_t(max_smem) / 2 : size_t(max_smem)) - 1024 - fixed;   // two CTAs per SM when they fit
    size_t R = room / (size_t(L.rec_words) * 4 + 4);
    R = std::min<size_t>(R, kPaMaxTileRows) / kPaThreads * kPaThreads;
    if (R < size_t(kPaThreads)) { set_error("partitioned aggregation: records too wide for a scatter tile"); return HDK_B200_E_UNSUPPORTED; }
    if (lvl == 0) g->tile_rows = uint32_t(R); else g->tile_rows = std::min<uint32_t>(g->tile_rows, uint32_t(R));
  }
  for (int lvl = 0; lvl < 2; ++lvl) {
    const size_t n_src = lvl == 0 ? num_fragments : F1;
    const size_t n_dest = lvl == 0 ? F1 : (size_t(1) << f2);
    g->scatter_smem[lvl] = pa_align(size_t(g->tile_rows) * L.rec_words * 4, 16) + n_dest * per_dest + 8 + (n_src + 1) * 4 + size_t(g->tile_rows) * 4 + 64;
  }
  // header: counts u32[P] | cursor2 u32[P] | cursor1 u32[F1] | base u64[P + 1] | work counter, out cursor
  g->header_bytes = pa_align(size_t(g->P) * 8 + size_t(F1) * 4 + 8 + (size_t(g->P) + 1) * 8 + 64, 256);
  g->rec_bytes = pa_align(size_t(total_rows) * L.rec_words * 4 + 64, 256);
  g->total_bytes = g->header_bytes + g->rec_bytes * (f2 ? 2 : 1) + 256;
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
                   hdk_b200_launch_info* info) {
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
  uint32_t* recs_a = reinterpret_cast<uint32_t*>(s + g.header_bytes);               // level 1 output
  uint32_t* recs_b = reinterpret_cast<uint32_t*>(s + g.header_bytes + g.rec_bytes);   // level 2 output
  a.error_codes = params->error_codes;
  HB_CUDA(cudaMemsetAsync(s, 0, g.header_bytes, st));
  const size_t count_smem = pa_align((size_t(a.num_fragments) + 1) * 4, 16) + size_t(g.P) * 4;
  HB_CUDA(cudaFuncSetAttribute(kern.count, cudaFuncAttributeMaxDynamicSharedMemorySize, int(count_smem)));
  kern.count<<<sm_count() * (count_smem > 100 * 1024 ? 1 : 2), kPaThreads, count_smem, st>>>(a);
  HB_LAUNCH_CHECK();
  pa_offsets_kernel<<<1, 1024, 0, st>>>(a.counts, base, g.P);
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
  ag.T = g.T;
  ag.counts = cursor2;
  ag.base = base;
  ag.recs = two_level ? recs_b : recs_a;
  ag.work_counter = work_counter;
  ag.out_cursor = out_cursor;
  ag.groupby_buf = params->groupby_buf;
  ag.error_codes = params->error_codes;
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
