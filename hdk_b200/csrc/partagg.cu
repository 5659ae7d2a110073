// hdk_b200/csrc/partagg.cu — radix-partitioned baseline-hash aggregation (high-cardinality group-by).
//
// The reference aggregates a baseline-hash group-by by probing ONE open-addressing table per kernel
// (get_group_value, QE/GroupByRuntime.cpp:31-54; GPU claim protocol QE/cuda_mapd_rt.cu:176-236) — on a GPU that is a
// random DRAM sector plus several global atomics per row once the table outgrows L2 (config 4: 2e8 entries).  Its own
// CPU answer to that shape is partitioned aggregation (QE/RelAlgExecutor.cpp:691-838: partition the rows by a hash of
// the key — QE/RowFuncBuilder.cpp:516-577 — so that every partition aggregates on its own).  Here the same idea runs
// INSIDE one GPU, down to partitions whose groups fit a CTA's shared memory:
//
//   1. count    every row's key → partition = hash range-reduced to P (≤ 32768); per-CTA shared histogram
//   2. offsets  exclusive scan of the P counters (one CTA)
//   3. scatter  every row that passes the filters becomes a packed RECORD (key values ‖ aggregate arguments, 4-byte
//               words) written at its partition's cursor; partitions are contiguous runs of the record buffer
//   4. aggregate one CTA per partition: open addressing in SHARED memory keyed by the row id of the group's first
//               ("representative") record — claiming is one 32-bit CAS, keys are compared against the representative's
//               record, so there is no multi-word publish protocol — neutral accumulators beside it, native 32-bit
//               shared atomics (64-bit integer SUM = two 32-bit adds with carry).  The finished groups are written
//               straight into the caller's group-by buffer in the reference layout (keys, "skip_val" NULL protocol,
//               compact slot widths: the same encoding finalize.cu produces), at consecutive entries reserved with one
//               global atomic per partition.  The order of entries in a baseline-hash buffer is free
//               (ResultSet iteration skips EMPTY keys; the reduction re-inserts by key), the rest keeps the init pattern.
//
// A partition whose groups do not fit the shared table is split by further hash bits and re-read (no global fallback).
// Traffic per row: columns once, record written once and read once; no global atomics on the aggregates.
#include <algorithm>
#include <cstdio>

#include "accum.cuh"
#include "baseline.cuh"
#include "common.cuh"
#include "eval.cuh"
#include "partagg.cuh"

namespace hb {

constexpr uint32_t kPaMaxPartitions = 32768;
constexpr uint32_t kPaMaxFragments = 4096;
constexpr int kPaThreads = 512;
constexpr int kPaRowsPerThread = 4;
constexpr int kPaTileRows = kPaThreads * kPaRowsPerThread;
constexpr int kAggThreads = 1024;
constexpr uint32_t kEmptyId = 0xffffffffu;
constexpr int kPaDirectMaxFields = 8;

struct PaArgs {
  DPlan plan;
  PaLayout lay;
  const int8_t* const* col_buffers;
  const int64_t* num_rows;
  uint32_t num_fragments;
  uint32_t P;
  uint32_t* counts;            // [P] rows per partition
  unsigned long long* base;    // [P + 1] first record of each partition
  uint32_t* cursor;            // [P]
  uint32_t* recs;              // records, rec_words words each
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

__device__ __forceinline__ int64_t pa_decode_col(const int8_t* base, uint64_t pos, int w, bool is_fp) {
  const int8_t* ptr = base + pos * uint64_t(w);
  if (is_fp) return w == 4 ? __double_as_longlong(double(*reinterpret_cast<const float*>(ptr))) : *reinterpret_cast<const int64_t*>(ptr);
  return w == 8 ? *reinterpret_cast<const int64_t*>(ptr) : w == 4 ? int64_t(*reinterpret_cast<const int32_t*>(ptr))
         : w == 2 ? int64_t(*reinterpret_cast<const int16_t*>(ptr)) : int64_t(*ptr);
}

// Evaluate one row with the interpreter: filters, keys, aggregate arguments.  Same error rules as the scan kernel
// (scan.cu process_row_generic): an error inside a qual is raised whether or not the row passes, any other only for rows
// that pass.  Returns false when the row contributes nothing.
__device__ __forceinline__ bool pa_eval_generic(const DPlan& p, const int8_t* const* cols, uint64_t pos, V* vals, int32_t& my_err) {
  auto load_outer = [&](int c, int w) -> uint64_t {
    const int8_t* ptr = cols[c] + pos * uint64_t(w);
    return w == 8 ? *reinterpret_cast<const uint64_t*>(ptr) : w == 4 ? uint64_t(*reinterpret_cast<const uint32_t*>(ptr))
           : w == 2 ? uint64_t(*reinterpret_cast<const uint16_t*>(ptr)) : uint64_t(*reinterpret_cast<const uint8_t*>(ptr));
  };
  auto load_inner = [&](int, int, int) -> uint64_t { return 0; };
  int32_t row_err = 0, qual_err = 0;
  for (int n = 0; n < p.n_exprs; ++n) {
    int32_t e = 0;
    vals[n] = eval_node(p, p.exprs[n], vals, e, load_outer, load_inner);
    if (e) { int32_t& dst = (p.exprs[n].aux & kAuxInQual) ? qual_err : row_err; if (!dst) dst = e; }
  }
  if (qual_err) { my_err = my_err > 0 ? my_err : qual_err; return false; }
  for (int f = 0; f < p.n_filters; ++f)
    if (!(vals[p.filters[f]].i > 0)) return false;
  if (row_err) { my_err = my_err > 0 ? my_err : row_err; return false; }
  return true;
}

__device__ __forceinline__ int64_t pa_key_cast(int64_t v, int key_width) { return key_width == 4 ? int64_t(int32_t(v)) : v; }

struct PaTileWalk {
  uint32_t frag;
  uint64_t row0, rows;
  const int8_t* const* cols;
};
__device__ __forceinline__ PaTileWalk pa_tile(const PaArgs& a, uint64_t tile, const uint32_t* frag_tile_prefix) {
  PaTileWalk t;
  uint32_t lo = 0, hi = a.num_fragments;   // last fragment whose first tile is <= tile
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (frag_tile_prefix[mid] <= tile) lo = mid; else hi = mid;
  }
  t.frag = lo;
  t.row0 = (tile - frag_tile_prefix[lo]) * uint64_t(kPaTileRows);
  t.rows = uint64_t(a.num_rows[lo]);
  t.cols = a.col_buffers + size_t(lo) * a.plan.n_cols;
  return t;
}
__device__ __forceinline__ void pa_tile_prefix(const PaArgs& a, uint32_t* frag_tile_prefix) {
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    frag_tile_prefix[0] = 0;
    for (uint32_t f = 0; f < a.num_fragments; ++f) {
      const int64_t rows = a.num_rows[f];
      acc += rows > 0 ? uint32_t((rows + kPaTileRows - 1) / kPaTileRows) : 0;
      frag_tile_prefix[f + 1] = acc;
    }
  }
}

// ---- pass 1 / pass 3: count and scatter --------------------------------------------------------------------------
// NF > 0: "direct" plans — no filters, every key and aggregate argument is a plain outer column: NF field values per row
// live in registers, no interpreter.  NF == 0: the interpreter evaluates the row into vals[] (local memory).
template <bool kScatter, int NF>
__global__ void __launch_bounds__(kPaThreads, 2) pa_pass_kernel(const __grid_constant__ PaArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  uint32_t* frag_tile_prefix = reinterpret_cast<uint32_t*>(dsm);                       // [num_fragments + 1]
  uint32_t* hist = reinterpret_cast<uint32_t*>(dsm) + ((a.num_fragments + 1 + 3) & ~3u);   // count pass: [P]
  const DPlan& p = a.plan;
  const PaLayout& L = a.lay;
  const int tid = threadIdx.x;
  pa_tile_prefix(a, frag_tile_prefix);
  if (!kScatter)
    for (uint32_t i = tid; i < a.P; i += kPaThreads) hist[i] = 0;
  __syncthreads();
  const uint32_t total_tiles = frag_tile_prefix[a.num_fragments];
  int32_t my_err = 0;
  for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const PaTileWalk t = pa_tile(a, tile, frag_tile_prefix);
    if constexpr (NF > 0) {
      const int8_t* fbase[NF];
#pragma unroll
      for (int i = 0; i < NF; ++i) fbase[i] = t.cols[L.f[i].col];
      int64_t fv[kPaRowsPerThread][NF];
      uint32_t part[kPaRowsPerThread];
      bool ok[kPaRowsPerThread];
#pragma unroll
      for (int r = 0; r < kPaRowsPerThread; ++r) {
        const uint64_t pos = t.row0 + uint64_t(r) * kPaThreads + tid;
        ok[r] = pos < t.rows;
        part[r] = 0;
        if (ok[r]) {
          KeyHasher kh(L.n_keys);
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            fv[r][i] = pa_decode_col(fbase[i], pos, L.f[i].colw, L.f[i].is_fp);
            if (i < L.n_keys) { fv[r][i] = pa_key_cast(fv[r][i], L.key_width); kh.add(fv[r][i]); }
          }
          part[r] = pa_partition(kh.finish(), a.P);
        }
      }
      if constexpr (!kScatter) {
#pragma unroll
        for (int r = 0; r < kPaRowsPerThread; ++r)
          if (ok[r]) atomicAdd(&hist[part[r]], 1u);
      } else {
        uint32_t idx[kPaRowsPerThread];
        unsigned long long b[kPaRowsPerThread];
#pragma unroll
        for (int r = 0; r < kPaRowsPerThread; ++r)
          if (ok[r]) { idx[r] = atomicAdd(a.cursor + part[r], 1u); b[r] = a.base[part[r]]; }
#pragma unroll
        for (int r = 0; r < kPaRowsPerThread; ++r) {
          if (!ok[r]) continue;
          uint32_t* dst = a.recs + (b[r] + idx[r]) * uint64_t(L.rec_words);
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            dst[L.f[i].off] = uint32_t(uint64_t(fv[r][i]));
            if (L.f[i].words == 2) dst[L.f[i].off + 1] = uint32_t(uint64_t(fv[r][i]) >> 32);
          }
        }
      }
    } else {
      V vals[HDK_B200_MAX_EXPRS];
      for (int r = 0; r < kPaRowsPerThread; ++r) {
        const uint64_t pos = t.row0 + uint64_t(r) * kPaThreads + tid;
        if (pos >= t.rows) continue;
        int32_t e = 0;
        if (!pa_eval_generic(p, t.cols, pos, vals, e)) { if (kScatter && e) my_err = my_err > 0 ? my_err : e; continue; }
        KeyHasher kh(L.n_keys);
        for (int i = 0; i < L.n_keys; ++i) kh.add(pa_key_cast(vals[L.f[i].expr].i, L.key_width));
        const uint32_t part = pa_partition(kh.finish(), a.P);
        if constexpr (!kScatter) {
          atomicAdd(&hist[part], 1u);
        } else {
          const uint32_t idx = atomicAdd(a.cursor + part, 1u);
          uint32_t* dst = a.recs + (a.base[part] + idx) * uint64_t(L.rec_words);
          for (int i = 0; i < L.n_fields; ++i) {
            int64_t v = vals[L.f[i].expr].i;
            if (i < L.n_keys) v = pa_key_cast(v, L.key_width);
            dst[L.f[i].off] = uint32_t(uint64_t(v));
            if (L.f[i].words == 2) dst[L.f[i].off + 1] = uint32_t(uint64_t(v) >> 32);
          }
        }
      }
    }
  }
  if constexpr (!kScatter) {
    __syncthreads();
    for (uint32_t i = tid; i < a.P; i += kPaThreads)
      if (hist[i]) atomicAdd(a.counts + i, hist[i]);
  } else {
    if (my_err) record_error(a.error_codes, my_err);
  }
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

// ---- pass 4: per-partition aggregation in shared memory ---------------------------------------------------------------
struct PaAggArgs {
  DPlan plan;
  DLayout layout;
  PaLayout lay;
  uint32_t P, T;                    // partitions, shared-table slots
  const uint32_t* counts;
  const unsigned long long* base;
  const uint32_t* recs;
  unsigned int* work_counter;       // next partition to take
  unsigned long long* out_cursor;   // entries of the group-by buffer handed out so far
  int64_t* const* groupby_buf;
  int32_t* error_codes;
  uint32_t acc_off[kMaxAcc];        // byte offset of accumulator a's cells inside dynamic shared memory (ids at 0)
};

__device__ __forceinline__ int64_t pa_field_value(const uint32_t* rec, const PaField& f) {
  if (f.words == 2) return int64_t(uint64_t(rec[f.off]) | (uint64_t(rec[f.off + 1]) << 32));
  return int64_t(int32_t(rec[f.off]));
}

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
  const int RW = F.rec_words;
  // empty table
  for (uint32_t s = tid; s < T; s += kAggThreads) {
    ids[s] = kEmptyId;
    for (int k = 0; k < p.n_acc; ++k) {
      if (p.accs[k].bytes == 4) reinterpret_cast<uint32_t*>(sm + a.acc_off[k])[s] = 0;
      else reinterpret_cast<int64_t*>(sm + a.acc_off[k])[s] = acc_identity(p.accs[k].kind);
    }
  }
  int8_t* const buf = reinterpret_cast<int8_t*>(a.groupby_buf[0]);
  const uint64_t E = L.entry_count;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_part = atomicAdd(a.work_counter, 1u);
    __syncthreads();
    const uint32_t part = s_part;
    if (part >= a.P) break;
    const uint32_t n = a.counts[part];
    if (n == 0) continue;
    const uint32_t* rows = a.recs + a.base[part] * uint64_t(RW);
    const int row_bits = n < (1u << 24) - 1 ? 24 : 32;      // spare bits of the id word hold a fingerprint of the key hash
    const uint32_t row_mask = row_bits == 32 ? 0xffffffffu : (1u << row_bits) - 1u;
    if (tid == 0) { s_stack[0][0] = 1; s_stack[0][1] = 0; s_top = 1; }
    for (;;) {
      __syncthreads();
      if (s_top == 0) break;
      const uint32_t mod = s_stack[s_top - 1][0], rem = s_stack[s_top - 1][1];
      __syncthreads();
      if (tid == 0) { --s_top; s_overflow = 0; s_ngroups = 0; s_emitted = 0; }
      __syncthreads();
      // ---- insert + accumulate
      for (uint32_t i = tid; i < n; i += kAggThreads) {
        const uint32_t* rec = rows + uint64_t(i) * RW;
        KeyHasher kh(F.n_keys);
        for (int k = 0; k < F.n_keys; ++k) kh.add(pa_field_value(rec, F.f[k]));
        const uint64_t h = kh.finish();
        if (mod > 1 && (uint32_t(h >> 13) & (mod - 1)) != rem) continue;
        const uint32_t fp = row_bits == 32 ? 0u : (uint32_t(h >> 45) & 0xffu) << 24;
        const uint32_t mine = fp | i;
        uint32_t slot = __umulhi(uint32_t(h), T);
        uint32_t probes = 0;
        bool found = false;
        while (!found) {
          uint32_t id = *reinterpret_cast<volatile uint32_t*>(ids + slot);
          if (id == kEmptyId) {
            id = atomicCAS(ids + slot, kEmptyId, mine);
            if (id == kEmptyId) break;      // claimed: this record represents the group
          }
          if ((id & ~row_mask) == fp) {     // same fingerprint: compare with the representative's key words
            const uint32_t* rep = rows + uint64_t(id & row_mask) * RW;
            bool eq = true;
            for (int w = 0; w < F.key_words && eq; ++w) eq = rep[w] == rec[w];
            found = eq;
          }
          if (!found) {
            slot = slot + 1 == T ? 0 : slot + 1;
            if (++probes >= T || s_overflow) { s_overflow = 1; break; }
          }
        }
        if (s_overflow) break;
        for (int k = 0; k < p.n_acc; ++k) {
          const DAcc acc = p.accs[k];
          uint8_t* cell = sm + a.acc_off[k] + size_t(slot) * acc.bytes;
          if (acc.kind == ACC_CNT_ALL) { atomicAdd(reinterpret_cast<uint32_t*>(cell), 1u); continue; }
          const PaField& fd = F.f[F.acc_field[k]];
          const int64_t v = pa_field_value(rec, fd);
          if (acc.arg_nullable) {
            const DExpr& te = p.exprs[acc.arg];
            const bool is_null = te.kind == HDK_B200_FP ? __longlong_as_double(v) == fp_null_of(te.width)
                                                        : (v == int_null_of(te.width) || (acc.arg_nullable == 2 && int32_t(v) == INT32_MIN));
            if (is_null) continue;
          }
          switch (acc.kind) {
            case ACC_CNT_NN: atomicAdd(reinterpret_cast<uint32_t*>(cell), 1u); break;
            case ACC_SUM_I: smem_add_i64(cell, v); break;
            case ACC_SUM_F: atomicAdd(reinterpret_cast<double*>(cell), __longlong_as_double(v)); break;
            case ACC_MIN_I: case ACC_MAX_I: bin_update_shared_atomic(acc.kind, cell, v); break;
            default: bin_update_shared_atomic(acc.kind, cell, f64_order_encode(__longlong_as_double(v))); break;   // MIN_F / MAX_F
          }
        }
      }
      __syncthreads();
      if (s_overflow) {
        // the groups of this sub-pass do not fit: forget them and split the sub-pass by two more hash bits
        for (uint32_t s = tid; s < T; s += kAggThreads) {
          ids[s] = kEmptyId;
          for (int k = 0; k < p.n_acc; ++k) {
            if (p.accs[k].bytes == 4) reinterpret_cast<uint32_t*>(sm + a.acc_off[k])[s] = 0;
            else reinterpret_cast<int64_t*>(sm + a.acc_off[k])[s] = acc_identity(p.accs[k].kind);
          }
        }
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
        ids[s] = kEmptyId;
        for (int k = 0; k < p.n_acc; ++k) {
          if (p.accs[k].bytes == 4) reinterpret_cast<uint32_t*>(sm + a.acc_off[k])[s] = 0;
          else reinterpret_cast<int64_t*>(sm + a.acc_off[k])[s] = acc_identity(p.accs[k].kind);
        }
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static size_t pa_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

int partagg_layout(const Lowered& lw, PaLayout* out) {
  const DPlan& p = lw.plan;
  PaLayout& L = *out;
  memset(&L, 0, sizeof(L));
  if (p.hash_type != HDK_B200_BASELINE_HASH || p.n_joins != 0 || p.n_keys < 1) return HDK_B200_E_UNSUPPORTED;
  L.n_keys = p.n_keys;
  L.key_width = lw.layout.key_width;
  int off = 0, nf = 0;
  bool direct = p.n_filters == 0;
  auto add_field = [&](int expr, bool is_key) -> int {
    const DExpr& e = p.exprs[expr];
    PaField& f = L.f[nf];
    f.expr = int16_t(expr);
    f.words = uint8_t((e.kind == HDK_B200_FP || e.width == 8) ? 2 : 1);
    if (is_key && lw.layout.key_width == 4) f.words = 1;
    f.off = uint8_t(off);
    f.is_fp = uint8_t(e.kind == HDK_B200_FP);
    f.col = -1;
    if (e.op == HDK_B200_OP_COL && e.a == 0 && !(e.aux & 1) && !e.guard) { f.col = e.b; f.colw = uint8_t(e.imm.i); }
    else direct = false;
    off += f.words;
    return nf++;
  };
  for (int k = 0; k < p.n_keys; ++k) add_field(p.keys[k].expr, true);
  L.key_words = off;
  for (int a = 0; a < p.n_acc; ++a) {
    L.acc_field[a] = -1;
    if (p.accs[a].arg < 0) continue;
    for (int i = p.n_keys; i < nf; ++i)
      if (L.f[i].expr == p.accs[a].arg) L.acc_field[a] = int8_t(i);
    if (L.acc_field[a] < 0) {
      if (nf >= kPaMaxFields) return HDK_B200_E_UNSUPPORTED;
      L.acc_field[a] = int8_t(add_field(p.accs[a].arg, false));
    }
  }
  if (off > 255) return HDK_B200_E_UNSUPPORTED;
  L.n_fields = nf;
  L.rec_words = off;
  L.direct = direct && nf <= kPaDirectMaxFields;
  return HDK_B200_OK;
}

struct PaGeometry {
  uint32_t P, T;
  size_t agg_smem;
  uint32_t acc_off[kMaxAcc];
  size_t header_bytes, total_bytes;
};

static int partagg_geometry(const Lowered& lw, const PaLayout& L, uint64_t total_rows, PaGeometry* g) {
  const DPlan& p = lw.plan;
  int dev = 0, max_smem = 0;
  HB_CUDA(cudaGetDevice(&dev));
  HB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  size_t per_slot = 4;
  for (int a = 0; a < p.n_acc; ++a) per_slot += p.accs[a].bytes;
  const size_t budget = size_t(max_smem) - 4096;
  uint32_t T = uint32_t(std::min<size_t>((budget - 64 * size_t(p.n_acc + 1)) / per_slot, 1u << 20));
  T &= ~127u;
  if (g_debug.pa_slots > 0) T = std::min<uint32_t>(T, std::max<uint32_t>(128u, uint32_t(g_debug.pa_slots) & ~127u));
  if (T < 128) { set_error("partitioned aggregation: accumulators do not fit in shared memory"); return HDK_B200_E_UNSUPPORTED; }
  // the caller sizes the table at 2 x the estimated number of groups (QE/RelAlgExecutor.cpp:1553-1557): aim at half-full
  // shared tables for that estimate; a partition that still overflows is split by further hash bits
  const uint64_t groups_est = std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(p.entry_count) / 2 + 1, total_rows));
  uint64_t P = (groups_est + T / 2 - 1) / (T / 2);
  // enough partitions to occupy the GPU even with few groups, as long as they keep a few thousand rows each
  P = std::max<uint64_t>(P, std::min<uint64_t>(uint64_t(sm_count()) * 8, total_rows / 4096));
  if (g_debug.pa_partitions > 0) P = uint64_t(g_debug.pa_partitions);
  P = std::max<uint64_t>(1, std::min<uint64_t>(P, kPaMaxPartitions));
  g->P = uint32_t(P);
  g->T = T;
  size_t off = pa_align(size_t(T) * 4, 16);
  for (int a = 0; a < p.n_acc; ++a) { g->acc_off[a] = uint32_t(off); off = pa_align(off + size_t(T) * p.accs[a].bytes, 16); }
  g->agg_smem = off;
  // header: counts u32[P] | cursor u32[P] | base u64[P + 1] | work counter, out cursor
  g->header_bytes = pa_align(size_t(P) * 8 + (size_t(P) + 1) * 8 + 64, 256);
  g->total_bytes = g->header_bytes + pa_align(size_t(total_rows) * L.rec_words * 4, 256) + 256;
  return HDK_B200_OK;
}

int partagg_scratch_bytes(const Lowered& lw, uint64_t total_rows, size_t* bytes) {
  PaLayout L;
  if (int rc = partagg_layout(lw, &L)) { set_error("plan shape not eligible for partitioned aggregation"); return rc; }
  if (total_rows >= 0xfffffffeull) { set_error("partitioned aggregation: too many rows per launch"); return HDK_B200_E_UNSUPPORTED; }
  PaGeometry g;
  if (int rc = partagg_geometry(lw, L, total_rows, &g)) return rc;
  *bytes = g.total_bytes;
  return HDK_B200_OK;
}

template <bool kScatter>
static void (*pa_pick_kernel(int nf))(const PaArgs) {
  switch (nf) {
    case 1: return pa_pass_kernel<kScatter, 1>;
    case 2: return pa_pass_kernel<kScatter, 2>;
    case 3: return pa_pass_kernel<kScatter, 3>;
    case 4: return pa_pass_kernel<kScatter, 4>;
    case 5: return pa_pass_kernel<kScatter, 5>;
    case 6: return pa_pass_kernel<kScatter, 6>;
    case 7: return pa_pass_kernel<kScatter, 7>;
    case 8: return pa_pass_kernel<kScatter, 8>;
    default: return pa_pass_kernel<kScatter, 0>;
  }
}

int launch_partagg(const Lowered& lw, const hdk_b200_kernel_params* params, void* scratch, size_t scratch_bytes, cudaStream_t st,
                   hdk_b200_launch_info* info) {
  PaLayout L;
  if (int rc = partagg_layout(lw, &L)) { set_error("plan shape not eligible for partitioned aggregation"); return rc; }
  const uint64_t total_rows = params->total_rows_hint;
  if (params->num_fragments > kPaMaxFragments) { set_error("more than %u fragments per launch", kPaMaxFragments); return HDK_B200_E_UNSUPPORTED; }
  PaGeometry g;
  if (int rc = partagg_geometry(lw, L, total_rows, &g)) return rc;
  if (scratch_bytes < g.total_bytes || !scratch) { set_error("partitioned aggregation needs %zu scratch bytes, got %zu", g.total_bytes, scratch_bytes); return HDK_B200_E_INVALID; }
  uint8_t* s = static_cast<uint8_t*>(scratch);
  PaArgs a{};
  a.plan = lw.plan;
  a.lay = L;
  a.col_buffers = params->col_buffers;
  a.num_rows = params->num_rows;
  a.num_fragments = uint32_t(params->num_fragments);
  a.P = g.P;
  a.counts = reinterpret_cast<uint32_t*>(s);
  a.cursor = a.counts + g.P;
  a.base = reinterpret_cast<unsigned long long*>(s + size_t(g.P) * 8);
  unsigned int* work_counter = reinterpret_cast<unsigned int*>(a.base + g.P + 1);
  unsigned long long* out_cursor = reinterpret_cast<unsigned long long*>(work_counter + 2);
  a.recs = reinterpret_cast<uint32_t*>(s + g.header_bytes);
  a.error_codes = params->error_codes;
  HB_CUDA(cudaMemsetAsync(s, 0, g.header_bytes, st));
  const int nf = (L.direct && !g_debug.force_generic) ? L.n_fields : 0;   // 0: the interpreter evaluates the rows
  const size_t prefix_bytes = pa_align((size_t(a.num_fragments) + 1) * 4, 16);
  const size_t count_smem = prefix_bytes + size_t(g.P) * 4;
  const int grid = sm_count() * 2;
  auto kc = pa_pick_kernel<false>(nf);
  auto ks = pa_pick_kernel<true>(nf);
  HB_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, int(count_smem)));
  kc<<<count_smem > 100 * 1024 ? sm_count() : grid, kPaThreads, count_smem, st>>>(a);
  HB_LAUNCH_CHECK();
  pa_offsets_kernel<<<1, 1024, 0, st>>>(a.counts, a.base, g.P);
  HB_LAUNCH_CHECK();
  ks<<<grid, kPaThreads, prefix_bytes, st>>>(a);
  HB_LAUNCH_CHECK();
  PaAggArgs ag{};
  ag.plan = lw.plan;
  ag.layout = lw.layout;
  ag.lay = L;
  ag.P = g.P;
  ag.T = g.T;
  ag.counts = a.counts;
  ag.base = a.base;
  ag.recs = a.recs;
  ag.work_counter = work_counter;
  ag.out_cursor = out_cursor;
  ag.groupby_buf = params->groupby_buf;
  ag.error_codes = params->error_codes;
  for (int k = 0; k < lw.plan.n_acc; ++k) ag.acc_off[k] = g.acc_off[k];
  HB_CUDA(cudaFuncSetAttribute(pa_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(g.agg_smem)));
  pa_aggregate_kernel<<<int(std::min<uint32_t>(g.P, uint32_t(sm_count()))), kAggThreads, g.agg_smem, st>>>(ag);
  HB_LAUNCH_CHECK();
  if (info) {
    info->variant = nf > 0 ? 1 : 0;
    info->strategy = HDK_B200_STRATEGY_PARTITIONED;
    info->n_launches = 4;
    info->grid = grid;
    info->block = kAggThreads;
    info->smem_bytes = int(g.agg_smem);
    info->n_accumulators = lw.plan.n_acc;
    info->tile_rows = int(g.P);
  }
  return HDK_B200_OK;
}

}  // namespace hb
