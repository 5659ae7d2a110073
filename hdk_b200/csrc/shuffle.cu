// hdk_b200/csrc/shuffle.cu — hash partitioning of rows by group key: the two passes of partitioned
// aggregation (count, scatter) that feed the NCCL all-to-all of the multi-GPU baseline-hash group-by.
//
// Model: the reference's CPU-only partitioned aggregation (QE/RelAlgExecutor.cpp:691-838):
//   partition function  MurmurHash64A over the 64-bit-widened key components (QE/RowFuncBuilder.cpp:516-577);
//                       the reference masks with (P-1), here multiply-shift range reduction so any GPU count works
//   pass 1 (COUNT)      per-partition histogram, reduced over kernels (QE/Execute.cpp:1365-1435)
//   pass 2 (scatter)    rows copied into pre-sized per-partition columnar buffers (QE/Execute.cpp:1933-2018)
#include <algorithm>
#include <cstdlib>

#include "baseline.cuh"
#include "common.cuh"
#include "eval.cuh"

namespace hb {

constexpr uint32_t kMaxPartitions = 1024;
constexpr uint32_t kMaxFragments = 4096;

// partition of a 64-bit key hash: multiply-shift range reduction of its upper half, floor(h_hi * n / 2^32) — one
// multiply instead of a modulo by a run-time value (~100 instructions for 64 bits).  Any deterministic function of
// the key serves (the reference masks MurmurHash64A with P - 1, QE/RowFuncBuilder.cpp:516-577).
__device__ __forceinline__ uint32_t mod_partitions(uint64_t h, uint32_t n) { return __umulhi(uint32_t(h >> 32), n); }

struct ShuffleArgs {
  DPlan plan;
  const int8_t* const* col_buffers;
  const int64_t* num_rows;
  uint32_t num_fragments;
  uint32_t n_partitions;
  unsigned long long* counts;      // pass 1
  const uint64_t* offsets;         // pass 2
  unsigned long long* cursors;     // pass 2
  int8_t* const* out_cols;         // pass 2
  // fast path: no filters and every group key is a plain outer column → keys are read directly, no interpreter
  int32_t direct;
  int8_t key_col[HDK_B200_MAX_KEYS];
  uint8_t key_w[HDK_B200_MAX_KEYS];
  uint8_t key_days[HDK_B200_MAX_KEYS];   // date column stored as days (decoded to seconds like fixed_width_small_date_decode)
  // region mode (hdk_b200_region_*): partition = region of the baseline group-by table the row's key hashes into
  uint32_t region_mode, entry_count, key_width, region_mul;   // region = min(umulhi(slot, region_mul), n_partitions - 1)
};

// partition id from the (64-bit widened) group keys.  The loops are unrolled over the maximum key count with an early
// exit so that keys[] is only ever indexed by constants and stays in registers.
__device__ __forceinline__ int partition_of_keys(const ShuffleArgs& a, const int64_t* keys, int n_keys) {
  if (a.region_mode) {
    // the slot the baseline probe starts at: key_hash (MurmurHash3 over the key bytes at key_width) % entry_count
    uint32_t h = 0;
#pragma unroll
    for (int k = 0; k < HDK_B200_MAX_KEYS; ++k) {
      if (k >= n_keys) break;
      h = mm3_block(h, uint32_t(uint64_t(keys[k])));
      if (a.key_width == 8) h = mm3_block(h, uint32_t(uint64_t(keys[k]) >> 32));
    }
    const uint32_t slot = mm3_final(h, uint32_t(n_keys) * a.key_width) % a.entry_count;
    return int(min(__umulhi(slot, a.region_mul), a.n_partitions - 1));
  }
  const uint64_t m = 0xc6a4a7935bd1e995ULL;   // MurmurHash64A over the 64-bit widened keys, seed 0 (baseline.cuh)
  uint64_t h = uint64_t(n_keys) * 8 * m;
#pragma unroll
  for (int k = 0; k < HDK_B200_MAX_KEYS; ++k) {
    if (k >= n_keys) break;
    uint64_t x = uint64_t(keys[k]) * m;
    x ^= x >> 47;
    x *= m;
    h ^= x;
    h *= m;
  }
  h ^= h >> 47;
  h *= m;
  h ^= h >> 47;
  return int(mod_partitions(h, a.n_partitions));
}

// kbase[k]: the tile's fragment base of key column k, loaded once per tile (no pointer → data chain per row)
__device__ __forceinline__ int row_partition_direct(const ShuffleArgs& a, const int8_t* const* kbase, uint64_t pos) {
  int64_t keys[HDK_B200_MAX_KEYS];
#pragma unroll
  for (int k = 0; k < HDK_B200_MAX_KEYS; ++k) {
    if (k >= a.plan.n_keys) break;
    const int w = a.key_w[k];
    const int8_t* ptr = kbase[k] + pos * w;
    int64_t v = w == 8 ? *reinterpret_cast<const int64_t*>(ptr) : w == 4 ? int64_t(*reinterpret_cast<const int32_t*>(ptr))
                : w == 2 ? int64_t(*reinterpret_cast<const int16_t*>(ptr)) : int64_t(*ptr);
    if (a.key_days[k]) v = (v == int_null_of(w)) ? INT64_MIN : v * 86400;
    keys[k] = v;
  }
  return partition_of_keys(a, keys, a.plan.n_keys);
}

// partition of one row, or -1 when the row is filtered out
__device__ __forceinline__ int row_partition(const ShuffleArgs& a, const int8_t* const* cols, uint64_t pos, V* vals) {
  const DPlan& p = a.plan;
  auto load_outer = [&](int c, int w) -> uint64_t {
    const int8_t* ptr = cols[c] + pos * w;
    return w == 8 ? *reinterpret_cast<const uint64_t*>(ptr) : w == 4 ? uint64_t(*reinterpret_cast<const uint32_t*>(ptr))
           : w == 2 ? uint64_t(*reinterpret_cast<const uint16_t*>(ptr)) : uint64_t(*reinterpret_cast<const uint8_t*>(ptr));
  };
  auto load_inner = [&](int, int, int) -> uint64_t { return 0; };
  int32_t err = 0;
  for (int n = 0; n < p.n_exprs; ++n) vals[n] = eval_node(p, p.exprs[n], vals, err, load_outer, load_inner);
  for (int f = 0; f < p.n_filters; ++f)
    if (!(vals[p.filters[f]].i > 0)) return -1;
  int64_t keys[HDK_B200_MAX_KEYS];
  for (int k = 0; k < p.n_keys; ++k) keys[k] = vals[p.keys[k].expr].i;
  return partition_of_keys(a, keys, p.n_keys);
}

// Lanes of the warp whose row goes to the same partition as this lane's (0 for a dropped row, part < 0).  MATCH.ANY costs
// over two SM cycles per row (tools/micro/gather_accum_bench.cu: the match_any variants); with the handful of partitions of a
// multi-GPU exchange one ballot per partition is several times cheaper.  All 32 lanes must call it.
__device__ __forceinline__ unsigned warp_peers(int part, uint32_t n_partitions) {
  if (n_partitions <= 16) {
    unsigned mine = 0;
    for (uint32_t q = 0; q < n_partitions; ++q) {
      const unsigned m = __ballot_sync(0xffffffffu, part == int(q));
      if (part == int(q)) mine = m;
    }
    return mine;
  }
  return __match_any_sync(0xffffffffu, part);
}

__global__ void shuffle_scatter_kernel(const __grid_constant__ ShuffleArgs a) {
  V vals[HDK_B200_MAX_EXPRS];
  const DPlan& p = a.plan;
  const uint64_t start = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (uint32_t f = 0; f < a.num_fragments; ++f) {
    const int8_t* const* cols = a.col_buffers + size_t(f) * p.n_cols;
    const uint64_t rows = uint64_t(a.num_rows[f]);
    const uint64_t rows_padded = (rows + step - 1) / step * step;  // keep warps converged for the match
    for (uint64_t pos = start; pos < rows_padded; pos += step) {
      const int part = pos < rows ? row_partition(a, cols, pos, vals) : -1;
      // warp-aggregated reservation: one atomic per distinct partition per warp
      const unsigned peers = warp_peers(part, a.n_partitions);
      if (part < 0) continue;
      const int leader = __ffs(peers) - 1;
      const int rank = __popc(peers & ((1u << lane) - 1));
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(a.cursors + part, (unsigned long long)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      const uint64_t dst = a.offsets[part] + base + rank;
      for (int c = 0; c < p.n_cols; ++c) {
        const int w = p.col_width[c];
        const int8_t* src = cols[c] + pos * w;
        int8_t* out = a.out_cols[c] + dst * w;
        if (w == 8) *reinterpret_cast<uint64_t*>(out) = *reinterpret_cast<const uint64_t*>(src);
        else if (w == 4) *reinterpret_cast<uint32_t*>(out) = *reinterpret_cast<const uint32_t*>(src);
        else if (w == 2) *reinterpret_cast<uint16_t*>(out) = *reinterpret_cast<const uint16_t*>(src);
        else *out = *src;
      }
    }
  }
}

// ---- tile-based passes -------------------------------------------------------------------------
// A CTA takes tiles of kTileRows rows; partition ids of a thread's rows stay in registers.  Counting: a shared
// histogram per CTA (warp-aggregated), flushed once.  Scattering: histogram of the tile → one global reservation per
// partition and tile → every row's position = reservation + its rank inside the tile (warp prefix via match / popc +
// per-warp offsets), so consecutive rows of a partition land next to each other (coalesced, also over NVLink).
constexpr int kShufThreads = 256;
constexpr int kShufRowsPerThread = 8;
constexpr int kTileRows = kShufThreads * kShufRowsPerThread;

struct ScatterToArgs {
  ShuffleArgs base;
  int8_t* const* dest_cols;        // [n_partitions * n_cols]
  const uint64_t* dest_offsets;    // [n_partitions]
  uint32_t off_part_of_pos, off_dest_ptrs;   // staged kernel: byte offsets inside dynamic shared memory (behind the staging area)
};

__device__ __forceinline__ void tile_of(const ShuffleArgs& a, uint64_t tile, const uint32_t* frag_tile_prefix, uint32_t& frag, uint64_t& row0) {
  // fragment holding this tile (few fragments: linear search over the prefix in shared memory)
  frag = 0;
  while (frag + 1 < a.num_fragments && frag_tile_prefix[frag + 1] <= tile) ++frag;
  row0 = (tile - frag_tile_prefix[frag]) * uint64_t(kTileRows);
}

__device__ __forceinline__ void copy_elem(void* o, const void* src, int w) {
  if (w == 8) *reinterpret_cast<uint64_t*>(o) = *reinterpret_cast<const uint64_t*>(src);
  else if (w == 4) *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(src);
  else if (w == 2) *reinterpret_cast<uint16_t*>(o) = *reinterpret_cast<const uint16_t*>(src);
  else *reinterpret_cast<uint8_t*>(o) = *reinterpret_cast<const uint8_t*>(src);
}

// kStaged (few partitions, narrow rows): the tile's rows are first regrouped by partition in shared memory, then every
// partition's run is copied out with consecutive threads writing consecutive elements — whole 128-byte lines per warp
// instead of a handful of elements per destination, which is what NVLink wants when the destinations are peers.
constexpr uint32_t kStagedMaxPartitions = 32;

template <bool kScatter, bool kStaged = false>
__global__ void __launch_bounds__(kShufThreads, 4) shuffle_tile_kernel(const __grid_constant__ ScatterToArgs sa) {
  const ShuffleArgs& a = sa.base;
  const DPlan& p = a.plan;
  extern __shared__ __align__(16) uint8_t staging[];      // kStaged: column-major copy of the tile, rows grouped by partition
  __shared__ unsigned int run_start[kStagedMaxPartitions + 1];
  __shared__ uint32_t stage_col_off[HDK_B200_MAX_COLS];
  __shared__ uint32_t frag_tile_prefix[kMaxFragments + 1];
  __shared__ unsigned int hist[kMaxPartitions];          // rows of the tile (scatter) / of the CTA (count) per partition
  __shared__ unsigned long long base[kMaxPartitions];    // scatter: reserved start per partition
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    uint32_t acc = 0;
    frag_tile_prefix[0] = 0;
    for (uint32_t f = 0; f < a.num_fragments; ++f) {
      const int64_t rows = a.num_rows[f];
      acc += rows > 0 ? uint32_t((rows + kTileRows - 1) / kTileRows) : 0;
      frag_tile_prefix[f + 1] = acc;
    }
  }
  for (uint32_t i = tid; i < a.n_partitions; i += kShufThreads) hist[i] = 0;
  // staged: the tile position → partition map and the destination column pointers live in shared memory, so the copy-out
  // below is one flat loop over the tile's elements without a dependent global load per (partition, column)
  uint8_t* part_of_pos = staging + sa.off_part_of_pos;
  int8_t** dest_s = reinterpret_cast<int8_t**>(staging + sa.off_dest_ptrs);
  if constexpr (kStaged) {
    for (uint32_t i = tid; i < a.n_partitions * uint32_t(p.n_cols); i += kShufThreads) dest_s[i] = sa.dest_cols[i];
  }
  __syncthreads();
  const uint32_t total_tiles = frag_tile_prefix[a.num_fragments];
  V vals[HDK_B200_MAX_EXPRS];
  for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    uint32_t frag;
    uint64_t row0;
    tile_of(a, tile, frag_tile_prefix, frag, row0);
    const int8_t* const* cols = a.col_buffers + size_t(frag) * p.n_cols;
    const uint64_t rows = uint64_t(a.num_rows[frag]);
    const int8_t* kbase[HDK_B200_MAX_KEYS];
    if (a.direct) {
#pragma unroll
      for (int k = 0; k < HDK_B200_MAX_KEYS; ++k) kbase[k] = k < p.n_keys ? cols[a.key_col[k]] : nullptr;
    }
    constexpr int kHoist = 8;                 // column bases kept in registers for the copies (more columns: re-read)
    const int8_t* cbase[kHoist];
    if (kScatter) {
#pragma unroll
      for (int c = 0; c < kHoist; ++c) cbase[c] = c < p.n_cols ? cols[c] : nullptr;
    }
    int part[kShufRowsPerThread];
#pragma unroll
    for (int r = 0; r < kShufRowsPerThread; ++r) {
      const uint64_t pos = row0 + uint64_t(r) * kShufThreads + tid;      // coalesced: consecutive threads, consecutive rows
      part[r] = pos < rows ? (a.direct ? row_partition_direct(a, kbase, pos) : row_partition(a, cols, pos, vals)) : -1;
    }
    // (a loop of its own: the key loads of all the thread's rows are in flight together, no warp vote between them)
#pragma unroll
    for (int r = 0; r < kShufRowsPerThread; ++r) {
      if (a.n_partitions > 32) {   // many partitions: lanes rarely meet, MATCH.ANY would cost one round per distinct value
        if (part[r] >= 0) atomicAdd(&hist[part[r]], 1u);
      } else {                     // warp-aggregated histogram update
        const unsigned peers = warp_peers(part[r], a.n_partitions);
        if (part[r] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&hist[part[r]], (unsigned)__popc(peers));
      }
    }
    if (!kScatter) continue;
    __syncthreads();
    // one reservation per partition and tile
    if (kStaged && tid == 0) {   // where each partition's run starts inside the tile (<= 32 partitions)
      unsigned int acc = 0;
      for (uint32_t i = 0; i < a.n_partitions; ++i) { run_start[i] = acc; acc += hist[i]; }
      run_start[a.n_partitions] = acc;
      uint32_t off = 0;
      for (int c = 0; c < p.n_cols; ++c) { stage_col_off[c] = off; off += uint32_t(kTileRows) * p.col_width[c]; }
    }
    for (uint32_t i = tid; i < a.n_partitions; i += kShufThreads) {
      base[i] = hist[i] ? sa.dest_offsets[i] + atomicAdd(a.cursors + i, (unsigned long long)hist[i]) : 0;
      hist[i] = 0;   // becomes the running position inside the reservation
    }
    __syncthreads();
    if constexpr (kStaged) {
      // position of every row of the thread inside the staged tile (rows grouped by partition)
      uint32_t lp[kShufRowsPerThread];
#pragma unroll
      for (int r = 0; r < kShufRowsPerThread; ++r) {
        const unsigned peers = warp_peers(part[r], a.n_partitions);
        lp[r] = 0xffffffffu;
        if (part[r] < 0) continue;
        const int leader = __ffs(peers) - 1;
        unsigned int start = 0;
        if (lane == leader) start = atomicAdd(&hist[part[r]], (unsigned)__popc(peers));
        start = __shfl_sync(peers, start, leader);
        lp[r] = run_start[part[r]] + start + __popc(peers & ((1u << lane) - 1u));
        part_of_pos[lp[r]] = uint8_t(part[r]);
      }
      // column by column: the loads of all the thread's rows are in flight together, then their stores
#pragma unroll
      for (int c = 0; c < kHoist; ++c) {
        if (c >= p.n_cols) break;
        const int w = p.col_width[c];
        uint64_t v[kShufRowsPerThread];
#pragma unroll
        for (int r = 0; r < kShufRowsPerThread; ++r) {
          const int8_t* src = cbase[c] + (row0 + uint64_t(r) * kShufThreads + tid) * w;
          v[r] = 0;
          if (lp[r] != 0xffffffffu)
            v[r] = w == 8 ? *reinterpret_cast<const uint64_t*>(src) : w == 4 ? uint64_t(*reinterpret_cast<const uint32_t*>(src))
                   : w == 2 ? uint64_t(*reinterpret_cast<const uint16_t*>(src)) : uint64_t(*reinterpret_cast<const uint8_t*>(src));
        }
        uint8_t* sc = staging + stage_col_off[c];
#pragma unroll
        for (int r = 0; r < kShufRowsPerThread; ++r) {
          if (lp[r] == 0xffffffffu) continue;
          uint8_t* dst = sc + size_t(lp[r]) * w;
          if (w == 8) *reinterpret_cast<uint64_t*>(dst) = v[r];
          else if (w == 4) *reinterpret_cast<uint32_t*>(dst) = uint32_t(v[r]);
          else if (w == 2) *reinterpret_cast<uint16_t*>(dst) = uint16_t(v[r]);
          else *dst = uint8_t(v[r]);
        }
      }
      for (int c = kHoist; c < p.n_cols; ++c) {
        const int w = p.col_width[c];
#pragma unroll
        for (int r = 0; r < kShufRowsPerThread; ++r)
          if (lp[r] != 0xffffffffu) copy_elem(staging + stage_col_off[c] + size_t(lp[r]) * w, cols[c] + (row0 + uint64_t(r) * kShufThreads + tid) * w, w);
      }
      __syncthreads();
      // copy the runs out: the staged tile is ordered by partition, so consecutive threads write consecutive elements of a
      // partition's run (whole 128-byte lines per warp except where a run ends)
      const uint32_t n_tile = run_start[a.n_partitions];
      for (int c = 0; c < p.n_cols; ++c) {
        const int w = p.col_width[c];
        const uint8_t* sc = staging + stage_col_off[c];
#pragma unroll 4
        for (uint32_t i = tid; i < n_tile; i += kShufThreads) {
          const uint32_t pr = part_of_pos[i];
          int8_t* out = dest_s[pr * uint32_t(p.n_cols) + uint32_t(c)] + (base[pr] + (i - run_start[pr])) * w;
          copy_elem(out, sc + size_t(i) * w, w);
        }
      }
      __syncthreads();
    } else {
#pragma unroll
    for (int r = 0; r < kShufRowsPerThread; ++r) {
      const unsigned peers = warp_peers(part[r], a.n_partitions);
      if (part[r] < 0) continue;
      const int leader = __ffs(peers) - 1;
      unsigned int start = 0;
      if (lane == leader) start = atomicAdd(&hist[part[r]], (unsigned)__popc(peers));   // shared-memory atomic, per warp and partition
      start = __shfl_sync(peers, start, leader);
      const uint64_t dst = base[part[r]] + start + __popc(peers & ((1u << lane) - 1u));
      const uint64_t pos = row0 + uint64_t(r) * kShufThreads + tid;
      int8_t* const* out = sa.dest_cols + size_t(part[r]) * p.n_cols;
#pragma unroll
      for (int c = 0; c < kHoist; ++c) {
        if (c >= p.n_cols) break;
        const int w = p.col_width[c];
        copy_elem(out[c] + dst * w, cbase[c] + pos * w, w);
      }
      for (int c = kHoist; c < p.n_cols; ++c) {
        const int w = p.col_width[c];
        copy_elem(out[c] + dst * w, cols[c] + pos * w, w);
      }
    }
    }
    __syncthreads();
    for (uint32_t i = tid; i < a.n_partitions; i += kShufThreads) hist[i] = 0;
    __syncthreads();
  }
  if (!kScatter) {
    __syncthreads();
    for (uint32_t i = tid; i < a.n_partitions; i += kShufThreads)
      if (hist[i]) atomicAdd(a.counts + i, (unsigned long long)hist[i]);
  } else {
    __threadfence_system();   // rows written to peer memory are visible before the host-side barrier that follows
  }
}

static void detect_direct_keys(const DPlan& p, ShuffleArgs* a) {
  a->direct = 0;
  if (p.n_filters) return;
  for (int k = 0; k < p.n_keys; ++k) {
    const DExpr& e = p.exprs[p.keys[k].expr];
    if (e.op != HDK_B200_OP_COL || e.a != 0 || e.kind != HDK_B200_INT) return;
    a->key_col[k] = e.b;
    a->key_w[k] = uint8_t(e.imm.i);
    a->key_days[k] = uint8_t(e.aux & 1);
  }
  a->direct = 1;
}

// ---- large-tile scatter ----------------------------------------------------------------------------------------------
// Up to 1024 partitions: a CTA regroups a tile of up to 8192 rows by partition in shared memory (column-major staging
// + the partition of every staged position), reserves one run per partition and tile with a single global atomic, and
// copies the staged tile out position by position — consecutive threads write consecutive elements of a partition's run
// (hundreds of bytes even with > 100 partitions), where the direct scatter writes a handful of elements per destination.
constexpr int kBigThreads = 512;
constexpr int kBigMaxRpt = 16;

struct BigScatterArgs {
  ScatterToArgs s;
  uint32_t tile_rows, rpt;                       // rows per tile (multiple of kBigThreads), rows per thread
  uint32_t off_part, off_hist, off_run, off_base, off_prefix;   // dynamic shared-memory map (staging at 0)
  uint32_t stage_off[HDK_B200_MAX_COLS];
};

__global__ void __launch_bounds__(kBigThreads, 1) scatter_big_tile_kernel(const __grid_constant__ BigScatterArgs ba) {
  extern __shared__ __align__(16) uint8_t sm[];
  const ShuffleArgs& a = ba.s.base;
  const DPlan& p = a.plan;
  uint16_t* part_of_pos = reinterpret_cast<uint16_t*>(sm + ba.off_part);
  unsigned int* hist = reinterpret_cast<unsigned int*>(sm + ba.off_hist);
  unsigned int* run_start = reinterpret_cast<unsigned int*>(sm + ba.off_run);
  unsigned long long* base = reinterpret_cast<unsigned long long*>(sm + ba.off_base);
  uint32_t* frag_tile_prefix = reinterpret_cast<uint32_t*>(sm + ba.off_prefix);
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t P = a.n_partitions;
  const bool many = P > 32;
  if (tid == 0) {
    uint32_t acc = 0;
    frag_tile_prefix[0] = 0;
    for (uint32_t f = 0; f < a.num_fragments; ++f) {
      const int64_t rows = a.num_rows[f];
      acc += rows > 0 ? uint32_t((rows + ba.tile_rows - 1) / ba.tile_rows) : 0;
      frag_tile_prefix[f + 1] = acc;
    }
  }
  __syncthreads();
  const uint32_t total_tiles = frag_tile_prefix[a.num_fragments];
  V vals[HDK_B200_MAX_EXPRS];
  for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    uint32_t frag = 0;
    while (frag + 1 < a.num_fragments && frag_tile_prefix[frag + 1] <= tile) ++frag;
    const uint64_t row0 = (tile - frag_tile_prefix[frag]) * uint64_t(ba.tile_rows);
    const int8_t* const* cols = a.col_buffers + size_t(frag) * p.n_cols;
    const uint64_t rows = uint64_t(a.num_rows[frag]);
    const int8_t* kbase[HDK_B200_MAX_KEYS];
    if (a.direct) {
#pragma unroll
      for (int k = 0; k < HDK_B200_MAX_KEYS; ++k) kbase[k] = k < p.n_keys ? cols[a.key_col[k]] : nullptr;
    }
    for (uint32_t i = tid; i < P; i += kBigThreads) hist[i] = 0;
    __syncthreads();
    int part[kBigMaxRpt];
#pragma unroll
    for (int r = 0; r < kBigMaxRpt; ++r) {
      part[r] = -1;
      if (uint32_t(r) < ba.rpt) {
        const uint64_t pos = row0 + uint64_t(r) * kBigThreads + tid;
        if (pos < rows) part[r] = a.direct ? row_partition_direct(a, kbase, pos) : row_partition(a, cols, pos, vals);
      }
    }
#pragma unroll
    for (int r = 0; r < kBigMaxRpt; ++r) {
      if (uint32_t(r) < ba.rpt) {
        if (many) {   // many partitions: lanes rarely meet, MATCH.ANY would cost one round per distinct value
          if (part[r] >= 0) atomicAdd(&hist[part[r]], 1u);
        } else {
          const unsigned peers = warp_peers(part[r], a.n_partitions);
          if (part[r] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&hist[part[r]], (unsigned)__popc(peers));
        }
      }
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scan of the histogram: lane l owns a contiguous chunk of ceil(P / 32) partitions
      const uint32_t chunk = (P + 31) / 32, lo = min(uint32_t(lane) * chunk, P), hi = min(lo + chunk, P);
      unsigned int sum = 0;
      for (uint32_t i = lo; i < hi; ++i) sum += hist[i];
      unsigned int incl = sum;
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      unsigned int run = incl - sum;
      for (uint32_t i = lo; i < hi; ++i) { run_start[i] = run; run += hist[i]; }
      if (lane == 31) run_start[P] = incl;
    }
    __syncthreads();
    for (uint32_t i = tid; i < P; i += kBigThreads) {
      base[i] = hist[i] ? ba.s.dest_offsets[i] + atomicAdd(a.cursors + i, (unsigned long long)hist[i]) : 0;
      hist[i] = 0;   // becomes the running position inside the run
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kBigMaxRpt; ++r) {
      if (uint32_t(r) < ba.rpt) {
        uint32_t lp = 0;
        if (many) {
          if (part[r] >= 0) lp = run_start[part[r]] + atomicAdd(&hist[part[r]], 1u);
        } else {
          const unsigned peers = warp_peers(part[r], a.n_partitions);
          if (part[r] >= 0) {
            const int leader = __ffs(peers) - 1;
            unsigned int start = 0;
            if (lane == leader) start = atomicAdd(&hist[part[r]], (unsigned)__popc(peers));
            start = __shfl_sync(peers, start, leader);
            lp = run_start[part[r]] + start + __popc(peers & ((1u << lane) - 1u));
          }
        }
        if (part[r] >= 0) {
          const uint64_t pos = row0 + uint64_t(r) * kBigThreads + tid;
          part_of_pos[lp] = uint16_t(part[r]);
          for (int c = 0; c < p.n_cols; ++c) {
            const int w = p.col_width[c];
            copy_elem(sm + ba.stage_off[c] + size_t(lp) * w, cols[c] + pos * w, w);
          }
        }
      }
    }
    __syncthreads();
    const uint32_t n_tile = run_start[P];
    for (int c = 0; c < p.n_cols; ++c) {
      const int w = p.col_width[c];
      const uint8_t* sc = sm + ba.stage_off[c];
      for (uint32_t i = tid; i < n_tile; i += kBigThreads) {
        const uint32_t pr = part_of_pos[i];
        int8_t* out = ba.s.dest_cols[size_t(pr) * p.n_cols + c] + (base[pr] + (i - run_start[pr])) * w;
        copy_elem(out, sc + size_t(i) * w, w);
      }
    }
    __syncthreads();
  }
  __threadfence_system();
}

// launch the large-tile scatter when its shared-memory plan fits; false = use the small-tile kernels
static bool launch_big_scatter(const ScatterToArgs& sa, const Lowered& lw, uint32_t n_partitions, cudaStream_t st, int* rc) {
  *rc = HDK_B200_OK;
  const size_t row_bytes = std::max<size_t>(lw.stage_row_bytes, 1);
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
  const size_t fixed = size_t(n_partitions) * (4 + 4 + 8) + 64 + (size_t(sa.base.num_fragments) + 1) * 4 + 64;
  uint32_t rpt = kBigMaxRpt;
  while (rpt >= 4 && size_t(rpt) * kBigThreads * (row_bytes + 2) + fixed + 256 > size_t(max_smem)) --rpt;
  if (rpt < 4 || lw.plan.n_cols > HDK_B200_MAX_COLS) return false;
  BigScatterArgs ba{};
  ba.s = sa;
  ba.rpt = rpt;
  ba.tile_rows = rpt * kBigThreads;
  size_t off = 0;
  for (int c = 0; c < lw.plan.n_cols; ++c) { ba.stage_off[c] = uint32_t(off); off += size_t(ba.tile_rows) * lw.plan.col_width[c]; off = (off + 15) & ~size_t(15); }
  ba.off_part = uint32_t(off); off += size_t(ba.tile_rows) * 2; off = (off + 15) & ~size_t(15);
  ba.off_hist = uint32_t(off); off += size_t(n_partitions) * 4; off = (off + 15) & ~size_t(15);
  ba.off_run = uint32_t(off); off += (size_t(n_partitions) + 1) * 4; off = (off + 15) & ~size_t(15);
  ba.off_base = uint32_t(off); off += size_t(n_partitions) * 8; off = (off + 15) & ~size_t(15);
  ba.off_prefix = uint32_t(off); off += (size_t(sa.base.num_fragments) + 1) * 4;
  if (off > size_t(max_smem)) return false;
  if (cudaFuncSetAttribute(scatter_big_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(off)) != cudaSuccess) { cudaGetLastError(); return false; }
  scatter_big_tile_kernel<<<sm_count(), kBigThreads, off, st>>>(ba);
  ++g_launch_count;
  if (cudaGetLastError() != cudaSuccess) { set_error("large-tile scatter launch failed"); *rc = HDK_B200_E_CUDA; }
  return true;
}

// a plan without layout information: lower only what the shuffle needs
static int lower_for_shuffle(const hdk_b200_plan* plan, Lowered* lw) {
  hdk_b200_qmd q{};
  q.hash_type = HDK_B200_BASELINE_HASH;
  q.entry_count = 1;
  q.key_count = plan ? plan->n_keys : 0;
  q.key_width = 8;
  // slots: fabricate 8-byte slots so that lower_plan accepts every target
  hdk_b200_plan p2 = *plan;
  int slot = 0;
  for (int t = 0; t < p2.n_targets; ++t) {
    if (p2.targets[t].agg == HDK_B200_AGG_NONE) { p2.targets[t].slot = -1; continue; }
    p2.targets[t].slot = slot;
    const int n = p2.targets[t].agg == HDK_B200_AGG_AVG ? 2 : 1;
    for (int i = 0; i < n; ++i) q.slot_padded[slot++] = 8;
  }
  q.slot_count = slot;
  return lower_plan(&p2, &q, lw);
}

}  // namespace hb

using namespace hb;

// the staged small-tile scatter: dynamic shared memory = staging area | tile position → partition bytes | destination pointers
static int launch_staged_scatter(ScatterToArgs& sa, const Lowered& lw, uint32_t n_partitions, size_t staging, cudaStream_t st) {
  sa.off_part_of_pos = uint32_t((staging + 15) & ~size_t(15));
  sa.off_dest_ptrs = uint32_t((sa.off_part_of_pos + size_t(kTileRows) + 15) & ~size_t(15));
  const size_t dyn = sa.off_dest_ptrs + size_t(n_partitions) * size_t(lw.plan.n_cols) * sizeof(void*);
  HB_CUDA(cudaFuncSetAttribute(shuffle_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn)));
  shuffle_tile_kernel<true, true><<<sm_count() * 3, kShufThreads, dyn, st>>>(sa);
  return HDK_B200_OK;
}

extern "C" {

int hdk_b200_shuffle_count(const hdk_b200_plan* plan, const hdk_b200_kernel_params* params, uint32_t n_partitions,
                           uint64_t* counts, void* stream) {
  if (!plan || !params || !counts || n_partitions == 0 || n_partitions > kMaxPartitions) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  if (plan->n_joins) { set_error("shuffle of joined plans is not supported"); return HDK_B200_E_UNSUPPORTED; }
  Lowered lw;
  if (int rc = lower_for_shuffle(plan, &lw)) return rc;
  ShuffleArgs a{};
  a.plan = lw.plan;
  a.col_buffers = params->col_buffers;
  a.num_rows = params->num_rows;
  a.num_fragments = uint32_t(params->num_fragments);
  a.n_partitions = n_partitions;
  a.counts = reinterpret_cast<unsigned long long*>(counts);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * n_partitions, st));
  if (a.num_fragments > kMaxFragments) { set_error("more than %u fragments", kMaxFragments); return HDK_B200_E_UNSUPPORTED; }
  ScatterToArgs sa{};
  sa.base = a;
  detect_direct_keys(lw.plan, &sa.base);
  shuffle_tile_kernel<false><<<sm_count() * 4, kShufThreads, 0, st>>>(sa);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

static int region_setup(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_params* params, uint32_t n_regions,
                        Lowered* lw, ScatterToArgs* sa) {
  if (!plan || !qmd || !params || n_regions == 0 || n_regions > kMaxPartitions) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  if (qmd->hash_type != HDK_B200_BASELINE_HASH) { set_error("region partitioning is for baseline-hash plans"); return HDK_B200_E_UNSUPPORTED; }
  if (plan->n_joins) { set_error("region partitioning of joined plans is not supported"); return HDK_B200_E_UNSUPPORTED; }
  if (params->num_fragments > kMaxFragments) { set_error("more than %u fragments", kMaxFragments); return HDK_B200_E_UNSUPPORTED; }
  if (int rc = lower_plan(plan, qmd, lw)) return rc;
  sa->base.plan = lw->plan;
  sa->base.col_buffers = params->col_buffers;
  sa->base.num_rows = params->num_rows;
  sa->base.num_fragments = uint32_t(params->num_fragments);
  sa->base.n_partitions = n_regions;
  sa->base.region_mode = 1;
  sa->base.entry_count = qmd->entry_count;
  sa->base.key_width = uint32_t(qmd->key_width);
  // region = floor(slot * n_regions / E) up to rounding: monotone in the slot, which is all that matters
  sa->base.region_mul = uint32_t(std::min<uint64_t>(((uint64_t(n_regions) << 32) + qmd->entry_count - 1) / qmd->entry_count, 0xffffffffull));
  detect_direct_keys(lw->plan, &sa->base);
  return HDK_B200_OK;
}

int hdk_b200_region_count(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_params* params, uint32_t n_regions,
                          uint64_t* counts, void* stream) {
  Lowered lw;
  ScatterToArgs sa{};
  if (int rc = region_setup(plan, qmd, params, n_regions, &lw, &sa)) return rc;
  if (!counts) { set_error("null counts"); return HDK_B200_E_INVALID; }
  sa.base.counts = reinterpret_cast<unsigned long long*>(counts);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * n_regions, st));
  shuffle_tile_kernel<false><<<sm_count() * 4, kShufThreads, 0, st>>>(sa);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_region_scatter_to(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_params* params,
                               uint32_t n_regions, int8_t* const* dest_cols, const uint64_t* dest_offsets, uint64_t* cursors, void* stream) {
  Lowered lw;
  ScatterToArgs sa{};
  if (int rc = region_setup(plan, qmd, params, n_regions, &lw, &sa)) return rc;
  if (!dest_cols || !dest_offsets || !cursors) { set_error("null argument"); return HDK_B200_E_INVALID; }
  sa.base.cursors = reinterpret_cast<unsigned long long*>(cursors);
  sa.dest_cols = dest_cols;
  sa.dest_offsets = dest_offsets;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CUDA(cudaMemsetAsync(cursors, 0, sizeof(uint64_t) * n_regions, st));
  const char* big_env = getenv("HDK_B200_SCATTER_BIG");   // tuning hook: 0 = never use the large-tile kernel
  if (n_regions >= 4 && !(big_env && big_env[0] == '0')) {
    int rc = HDK_B200_OK;
    if (launch_big_scatter(sa, lw, n_regions, st, &rc)) return rc;
  }
  const size_t staging = size_t(kTileRows) * lw.stage_row_bytes;
  if (n_regions <= kStagedMaxPartitions && staging <= 96 * 1024) {
    if (int rc = launch_staged_scatter(sa, lw, n_regions, staging, st)) return rc;
  } else {
    shuffle_tile_kernel<true><<<sm_count() * 4, kShufThreads, 0, st>>>(sa);
  }
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_shuffle_scatter_to(const hdk_b200_plan* plan, const hdk_b200_kernel_params* params, uint32_t n_partitions,
                                int8_t* const* dest_cols, const uint64_t* dest_offsets, uint64_t* cursors, void* stream) {
  if (!plan || !params || !dest_cols || !dest_offsets || !cursors || n_partitions == 0 || n_partitions > kMaxPartitions) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  if (plan->n_joins) { set_error("shuffle of joined plans is not supported"); return HDK_B200_E_UNSUPPORTED; }
  if (params->num_fragments > kMaxFragments) { set_error("more than %u fragments", kMaxFragments); return HDK_B200_E_UNSUPPORTED; }
  Lowered lw;
  if (int rc = lower_for_shuffle(plan, &lw)) return rc;
  ScatterToArgs sa{};
  sa.base.plan = lw.plan;
  sa.base.col_buffers = params->col_buffers;
  sa.base.num_rows = params->num_rows;
  sa.base.num_fragments = uint32_t(params->num_fragments);
  sa.base.n_partitions = n_partitions;
  sa.base.cursors = reinterpret_cast<unsigned long long*>(cursors);
  sa.dest_cols = dest_cols;
  sa.dest_offsets = dest_offsets;
  detect_direct_keys(lw.plan, &sa.base);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CUDA(cudaMemsetAsync(cursors, 0, sizeof(uint64_t) * n_partitions, st));
  const char* big_env = getenv("HDK_B200_SCATTER_BIG");   // tuning hook: 0 = never, 1 = also for few partitions
  if ((n_partitions > kStagedMaxPartitions || (big_env && big_env[0] == '1')) && !(big_env && big_env[0] == '0')) {
    int rc = HDK_B200_OK;
    if (launch_big_scatter(sa, lw, n_partitions, st, &rc)) return rc;
  }
  const size_t staging = size_t(kTileRows) * lw.stage_row_bytes;
  bool staged = n_partitions >= 4 && n_partitions <= kStagedMaxPartitions && staging <= 96 * 1024;
  if (const char* env = getenv("HDK_B200_SCATTER_STAGED")) staged = env[0] == '1' && n_partitions <= kStagedMaxPartitions && staging <= 96 * 1024;   // tuning hook
  if (staged) {
    if (int rc = launch_staged_scatter(sa, lw, n_partitions, staging, st)) return rc;
  } else {
    shuffle_tile_kernel<true><<<sm_count() * 4, kShufThreads, 0, st>>>(sa);
  }
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

int hdk_b200_shuffle_scatter(const hdk_b200_plan* plan, const hdk_b200_kernel_params* params, uint32_t n_partitions,
                             const uint64_t* offsets, uint64_t* cursors, int8_t* const* out_cols, void* stream) {
  if (!plan || !params || !offsets || !cursors || !out_cols || n_partitions == 0 || n_partitions > kMaxPartitions) { set_error("bad argument"); return HDK_B200_E_INVALID; }
  if (plan->n_joins) { set_error("shuffle of joined plans is not supported"); return HDK_B200_E_UNSUPPORTED; }
  Lowered lw;
  if (int rc = lower_for_shuffle(plan, &lw)) return rc;
  ShuffleArgs a{};
  a.plan = lw.plan;
  a.col_buffers = params->col_buffers;
  a.num_rows = params->num_rows;
  a.num_fragments = uint32_t(params->num_fragments);
  a.n_partitions = n_partitions;
  a.offsets = offsets;
  a.cursors = reinterpret_cast<unsigned long long*>(cursors);
  a.out_cols = out_cols;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HB_CUDA(cudaMemsetAsync(cursors, 0, sizeof(uint64_t) * n_partitions, st));
  shuffle_scatter_kernel<<<sm_count() * 4, 256, 0, st>>>(a);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

}  // extern "C"
