// hdk_b200/csrc/sort.cu — ORDER BY / LIMIT over an aggregated result, on the device.
//
// The reference sorts a ResultSet by building a permutation of its non-empty entries with a comparator over the
// ORDER BY targets (sortResultSet, QE/ResultSetSort.cpp:752-851; ResultSetComparator::operator(), :333-470: per order
// entry NULL placement by `nulls_first`, then `(lhs < rhs) != is_desc`, dictionary targets by string), with
// std::partial_sort for LIMIT (topPermutation, :504-520) — host work proportional to the number of groups, which
// for a baseline-hash result (1e8 groups) dwarfs the aggregation itself.
//
// Here every ORDER BY target becomes one 64-bit key whose unsigned order IS the comparator's order, and the
// permutation comes from a stable LSD radix sort (8-bit digits) of (key, row id) pairs, last ORDER BY target first.
// Input: the dense 8-byte result columns hdk_b200_compact_result produces.  HBM-bound: per pass one read of the keys
// for the histogram, one read + one write of the pairs for the scatter; digits on which all keys agree (known from
// the OR / AND of the keys, reduced while they are generated) cost nothing, so a COUNT column below 2^24 takes three
// passes, not eight.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace hb {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
#ifndef HB_SORT_ITEMS
#define HB_SORT_ITEMS 12   // 4: 12.8 ms, 8: 10.5 ms, 12: 10.0 ms per 1e8 41-bit keys (longer runs per digit in the staged write)
#endif
constexpr int kSortItems = HB_SORT_ITEMS;                      // elements per thread per chunk
constexpr int kSortChunk = kSortThreads * kSortItems;          // 3072
constexpr int kSortMaxCtas = 148 * 4;

struct SortState {
  unsigned long long key_or, key_and;   // over all keys of the current ORDER BY target
  unsigned int totals[8][256];          // per radix pass: rows per digit over all CTAs
};

struct SortBufs {
  uint64_t* keys[2];
  uint32_t* idx[2];      // idx[0] = the caller's permutation buffer
  uint32_t* hist;        // [n_ctas][256]
  SortState* state;
  uint64_t n;
  uint32_t n_ctas;
  uint64_t per_cta;      // elements per CTA range, a multiple of kSortChunk
};

// digit p is constant over all keys ⇔ OR and AND agree on its bits
__device__ __forceinline__ bool digit_varies(const SortState* s, int p) { return (((s->key_or ^ s->key_and) >> (8 * p)) & 0xff) != 0; }
// which of the two buffers holds the data before pass p: every pass that ran flipped it
__device__ __forceinline__ int cur_before(const SortState* s, int p) {
  int c = 0;
  for (int q = 0; q < p; ++q) c ^= int(digit_varies(s, q));
  return c;
}

// --------------------------------------------------------------------------------------------- keys
struct KeyArgs {
  const int64_t* col;
  const int32_t* dict_rank;
  int64_t dict_size;
  int32_t is_fp, type_width, nullable, is_desc, nulls_first;
};

// unsigned key whose ascending order is the comparator's order (QE/ResultSetSort.cpp:410-470)
__device__ __forceinline__ uint64_t order_key(const KeyArgs& k, int64_t cell) {
  bool is_null;
  uint64_t u;
  if (k.is_fp) {
    const double d = __longlong_as_double(cell);
    is_null = k.nullable && d == (k.type_width == 4 ? double(1.17549435e-38f) : 2.2250738585072014e-308);
    const uint64_t b = d == 0.0 ? 0ull : uint64_t(cell);   // -0.0 and +0.0 compare equal
    u = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
  } else {
    is_null = k.nullable && cell == int_null_of(k.type_width);
    if (k.dict_rank && !is_null) cell = (cell >= 0 && cell < k.dict_size) ? int64_t(k.dict_rank[cell]) : cell;
    u = uint64_t(cell) ^ 0x8000000000000000ull;
  }
  if (!k.nullable) return k.is_desc ? ~u : u;
  if (is_null) return k.nulls_first ? 0ull : ~0ull;
  // one code at the NULL end is kept free: an integer's sentinel is its type minimum, so non-NULL u >= 1
  // (fp: the extreme codes are NaN payloads no arithmetic produces)
  uint64_t r = k.is_desc ? ~u : (k.is_fp ? u : u - 1);            // ints: [0, 2^64 - 2] either way
  if (k.is_fp) r = min(max(r, uint64_t(1)), ~uint64_t(1)) - 1;    // clamp into [0, 2^64 - 3]
  return k.nulls_first ? r + 1 : r;
}

__global__ void __launch_bounds__(256) sort_keys_kernel(const KeyArgs k, SortBufs b, int first) {
  unsigned long long o = 0, a = ~0ull;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < b.n; i += step) {
    uint32_t row;
    if (first) { row = uint32_t(i); b.idx[0][i] = row; } else row = b.idx[0][i];
    const uint64_t key = order_key(k, k.col[row]);
    b.keys[0][i] = key;
    o |= key;
    a &= key;
  }
  for (int d = 16; d; d >>= 1) {
    o |= __shfl_xor_sync(0xffffffffu, o, d);
    a &= __shfl_xor_sync(0xffffffffu, a, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicOr(&b.state->key_or, o);
    atomicAnd(&b.state->key_and, a);
  }
}

// --------------------------------------------------------------------------------------------- one radix pass
// rows per digit of this CTA's range → hist[cta][digit]; and, summed over the CTAs, → state->totals[pass][digit]
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(SortBufs b, int pass) {
  if (!digit_varies(b.state, pass)) return;
  __shared__ uint32_t bins[256];
  bins[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t* __restrict__ src = cur_before(b.state, pass) ? b.keys[1] : b.keys[0];
  const uint64_t lo = blockIdx.x * b.per_cta, hi = min(b.n, lo + b.per_cta);
  const int sh = 8 * pass;
  uint64_t i = lo + threadIdx.x;
  for (; i + 3 * kSortThreads < hi; i += 4 * kSortThreads) {      // four independent loads in flight per thread
    const uint64_t k0 = src[i], k1 = src[i + kSortThreads], k2 = src[i + 2 * kSortThreads], k3 = src[i + 3 * kSortThreads];
    atomicAdd(&bins[(k0 >> sh) & 0xff], 1u);
    atomicAdd(&bins[(k1 >> sh) & 0xff], 1u);
    atomicAdd(&bins[(k2 >> sh) & 0xff], 1u);
    atomicAdd(&bins[(k3 >> sh) & 0xff], 1u);
  }
  for (; i < hi; i += kSortThreads) atomicAdd(&bins[(src[i] >> sh) & 0xff], 1u);
  __syncthreads();
  const uint32_t c = bins[threadIdx.x];
  b.hist[uint64_t(blockIdx.x) * 256 + threadIdx.x] = c;
  if (c) atomicAdd(&b.state->totals[pass][threadIdx.x], c);
}

// Stable scatter.  A CTA owns a contiguous range and walks it chunk by chunk; inside a chunk warp w owns the
// elements [w*32*kSortItems, (w+1)*32*kSortItems) in (item, lane) order.  Output position of an element = rows of smaller digits (all
// CTAs) + rows of its digit in earlier CTAs + in earlier chunks of this CTA + in earlier warps of the chunk + earlier
// elements of the same digit in its own warp (lanes with equal digits found with one ballot per digit bit).
// kStaged: the chunk is first regrouped by digit in shared memory and written out run by run, so that neighbouring
// threads write neighbouring addresses instead of 8-byte pieces of 256 different streams.
template <bool kStaged>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(SortBufs b, int pass) {
  if (!digit_varies(b.state, pass)) return;
  __shared__ uint32_t cnt[kSortWarps][257];     // [.][256] collects the out-of-range lanes of the last chunk
  __shared__ uint32_t off[256];                 // next output position of each digit for this CTA
  __shared__ uint32_t scan[256];
  __shared__ uint64_t skey[kStaged ? kSortChunk : 1];
  __shared__ uint32_t srow[kStaged ? kSortChunk : 1];
  const int cur = cur_before(b.state, pass);
  const uint64_t* __restrict__ ksrc = cur ? b.keys[1] : b.keys[0];
  const uint32_t* __restrict__ isrc = cur ? b.idx[1] : b.idx[0];
  uint64_t* __restrict__ kdst = cur ? b.keys[0] : b.keys[1];
  uint32_t* __restrict__ idst = cur ? b.idx[0] : b.idx[1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int sh = 8 * pass;
  {
    uint32_t before = 0;                        // rows of digit `tid` in the CTAs ahead of this one (coalesced rows of 1 KB)
    uint32_t c = 0;
    for (; c + 4 <= blockIdx.x; c += 4)
      before += b.hist[uint64_t(c) * 256 + tid] + b.hist[uint64_t(c + 1) * 256 + tid] + b.hist[uint64_t(c + 2) * 256 + tid] +
                b.hist[uint64_t(c + 3) * 256 + tid];
    for (; c < blockIdx.x; ++c) before += b.hist[uint64_t(c) * 256 + tid];
    const uint32_t total = b.state->totals[pass][tid];
    scan[tid] = total;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {         // inclusive scan over the digits
      const uint32_t v = tid >= d ? scan[tid - d] : 0;
      __syncthreads();
      scan[tid] += v;
      __syncthreads();
    }
    off[tid] = scan[tid] - total + before;
  }
  const uint64_t lo = blockIdx.x * b.per_cta, hi = min(b.n, lo + b.per_cta);
  for (uint64_t base = lo; base < hi; base += kSortChunk) {
    for (int i = tid; i < kSortWarps * 257; i += kSortThreads) (&cnt[0][0])[i] = 0;
    __syncthreads();
    uint64_t key[kSortItems];
    uint32_t row[kSortItems], rank[kSortItems];
    int dig[kSortItems];
    const uint64_t wbase = base + uint64_t(warp) * (32 * kSortItems) + lane;
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const uint64_t i = wbase + uint64_t(j) * 32;
      const bool ok = i < hi;
      key[j] = ok ? ksrc[i] : 0;
      row[j] = ok ? isrc[i] : 0;
      dig[j] = ok ? int((key[j] >> sh) & 0xff) : 256;
    }
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
      const int d = dig[j];
      unsigned peers = __ballot_sync(0xffffffffu, d < 256);
      if (d >= 256) peers = ~peers;
#pragma unroll
      for (int bit = 0; bit < 8; ++bit) {
        const unsigned m = __ballot_sync(0xffffffffu, (d >> bit) & 1);
        peers &= ((d >> bit) & 1) ? m : ~m;
      }
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = cnt[warp][d];
        cnt[warp][d] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[j] = old + __popc(peers & lt);
      __syncwarp();
    }
    __syncthreads();
    if constexpr (!kStaged) {
      uint32_t run = off[tid];                  // 256 threads = 256 digits
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t c = cnt[w][tid];
        cnt[w][tid] = run;
        run += c;
      }
      off[tid] = run;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kSortItems; ++j) {
        if (dig[j] < 256) {
          const uint32_t pos = cnt[warp][dig[j]] + rank[j];
          kdst[pos] = key[j];
          idst[pos] = row[j];
        }
      }
      __syncthreads();
    } else {
      // positions inside the chunk: digits in order, inside a digit warps in order
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t c = cnt[w][tid];
        cnt[w][tid] = run;                      // rows of digit tid in earlier warps of the chunk
        run += c;
      }
      scan[tid] = run;                          // rows of digit tid in the chunk
      __syncthreads();
      for (int d = 1; d < 256; d <<= 1) {
        const uint32_t v = tid >= d ? scan[tid - d] : 0;
        __syncthreads();
        scan[tid] += v;
        __syncthreads();
      }
      const uint32_t first = scan[tid] - run;   // chunk slot of digit tid's first row
      __syncthreads();
      scan[tid] = first;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kSortItems; ++j) {
        if (dig[j] < 256) {
          const uint32_t slot = scan[dig[j]] + cnt[warp][dig[j]] + rank[j];
          skey[slot] = key[j];
          srow[slot] = row[j];
        }
      }
      __syncthreads();
      const uint32_t in_chunk = uint32_t(min(uint64_t(kSortChunk), hi - base));
      for (uint32_t slot = tid; slot < in_chunk; slot += kSortThreads) {
        const uint64_t k = skey[slot];
        const int d = int((k >> sh) & 0xff);
        const uint32_t pos = off[d] + (slot - scan[d]);
        kdst[pos] = k;
        idst[pos] = srow[slot];
      }
      __syncthreads();
      off[tid] += run;
    }
  }
}

// after the last pass of a target: bring the permutation back to idx[0] if an odd number of passes ran
__global__ void __launch_bounds__(256) sort_settle_kernel(SortBufs b) {
  if (cur_before(b.state, 8) == 0) return;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < b.n; i += step) b.idx[0][i] = b.idx[1][i];
}

// --------------------------------------------------------------------------------------------- LIMIT prefilter
// ORDER BY … LIMIT k over n >> k rows: only rows whose PRIMARY key is among the k smallest can be in the answer
// (topPermutation keeps the k least rows, QE/ResultSetSort.cpp:504-520).  An MSD radix select finds the bucket that
// holds the k-th smallest primary key — one histogram pass over the keys per digit, stopping as soon as the rows up to and
// including that bucket are few — and the rows with key <= the bucket's upper bound (every tie included, the later
// ORDER BY targets still decide among them) are compacted into the candidate list the LSD sort then runs on.
struct SelectState {
  unsigned long long prefix, mask;   // the digits fixed so far
  unsigned long long k;              // rank still to find inside the current bucket (1-based)
  unsigned long long threshold;      // result: candidates are the rows with key <= threshold
  unsigned long long n_candidates;
  unsigned int done;
  unsigned int hist[256];
};

__global__ void __launch_bounds__(256) select_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, const SortState* ss,
                                                          SelectState* s, int pass) {
  if (s->done || !digit_varies(ss, pass)) return;
  __shared__ uint32_t bins[256];
  bins[threadIdx.x] = 0;
  __syncthreads();
  const unsigned long long prefix = s->prefix, mask = s->mask;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += step) {
    const uint64_t key = keys[i];
    if ((key & mask) == prefix) atomicAdd(&bins[(key >> (8 * pass)) & 0xff], 1u);
  }
  __syncthreads();
  if (bins[threadIdx.x]) atomicAdd(&s->hist[threadIdx.x], bins[threadIdx.x]);
}

__global__ void select_pick_kernel(const SortState* ss, SelectState* s, int pass, unsigned long long top_n, unsigned long long stop_at) {
  if (threadIdx.x || s->done) return;
  const unsigned long long digit_mask = 0xffull << (8 * pass);
  if (!digit_varies(ss, pass)) {             // every key has the same digit here: it joins the prefix
    s->prefix |= ss->key_or & digit_mask;
    s->mask |= digit_mask;
  } else {
    unsigned long long cum = 0;
    int d = 0;
    for (; d < 255; ++d) {
      if (cum + s->hist[d] >= s->k) break;
      cum += s->hist[d];
    }
    const unsigned long long in_bucket = s->hist[d];
    s->k -= cum;
    s->prefix |= (unsigned long long)d << (8 * pass);
    s->mask |= digit_mask;
    for (int i = 0; i < 256; ++i) s->hist[i] = 0;
    // rows before this bucket = top_n - k; stop refining once [everything up to and including the bucket] is small
    if ((top_n - s->k) + in_bucket <= stop_at) s->done = 1;
  }
  if (pass == 0) s->done = 1;
  if (s->done) s->threshold = s->prefix | ~s->mask;    // upper bound of the bucket
}

__global__ void __launch_bounds__(256) select_compact_kernel(const uint64_t* __restrict__ keys, uint64_t n, SelectState* s,
                                                             uint32_t* __restrict__ rows_out) {
  const unsigned long long thr = s->threshold;
  const int lane = threadIdx.x & 31;
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t base = blockIdx.x * uint64_t(blockDim.x); base < n; base += step) {   // whole warps iterate together
    const uint64_t i = base + threadIdx.x;
    const bool keep = i < n && keys[i] <= thr;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    unsigned long long wbase = 0;
    if (lane == 0 && m) wbase = atomicAdd(&s->n_candidates, (unsigned long long)__popc(m));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (keep) rows_out[wbase + __popc(m & ((1u << lane) - 1u))] = uint32_t(i);
  }
}

struct GatherArgs {
  const int64_t* in[HDK_B200_MAX_TARGETS];
  int64_t* out[HDK_B200_MAX_TARGETS];
  int32_t n_cols;
  const uint32_t* perm;
  uint64_t n_out;
};

__global__ void __launch_bounds__(256) gather_rows_kernel(const __grid_constant__ GatherArgs a) {
  const uint64_t step = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < a.n_out; i += step) {
    const uint32_t r = a.perm[i];
    for (int c = 0; c < a.n_cols; ++c) a.out[c][i] = a.in[c][r];
  }
}

static inline uint32_t sort_ctas(uint64_t n) { return uint32_t(std::min<uint64_t>(kSortMaxCtas, (n + kSortChunk - 1) / kSortChunk)); }
static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace hb

using namespace hb;

extern "C" {

size_t hdk_b200_sort_scratch_bytes(uint64_t n_rows) {
  if (n_rows == 0) return 256;
  return align256(n_rows * 8) * 2 + align256(n_rows * 4) + align256(size_t(256) * sort_ctas(n_rows) * 4) +
         align256(sizeof(SortState)) + align256(sizeof(SelectState));
}

// the LSD sort of the rows listed in b.idx[0] (or of all rows, identity, when `identity`)
static int sort_rows_lsd(SortBufs b, const int64_t* const* cols, const hdk_b200_order_entry* order, int n_order, bool identity,
                         cudaStream_t st) {
  const uint64_t chunks = (b.n + kSortChunk - 1) / kSortChunk;
  b.n_ctas = sort_ctas(b.n);
  b.per_cta = (chunks + b.n_ctas - 1) / b.n_ctas * kSortChunk;
  const int flat_grid = int(std::min<uint64_t>(148 * 8, (b.n + 255) / 256));
  static const bool staged = [] { const char* e = getenv("HDK_B200_SORT_STAGED"); return e ? atoi(e) != 0 : true; }();   // tuning hook
  for (int o = n_order - 1; o >= 0; --o) {       // LSD over the ORDER BY list: least significant target first
    const hdk_b200_order_entry& oe = order[o];
    KeyArgs k{cols[oe.column], oe.dict_rank, oe.dict_size, oe.is_fp, oe.type_width, oe.nullable, oe.is_desc, oe.nulls_first};
    HB_CUDA(cudaMemsetAsync(b.state, 0x00, sizeof(SortState), st));
    HB_CUDA(cudaMemsetAsync(&b.state->key_and, 0xff, 8, st));
    sort_keys_kernel<<<flat_grid, 256, 0, st>>>(k, b, identity && o == n_order - 1);
    HB_LAUNCH_CHECK();
    for (int pass = 0; pass < 8; ++pass) {
      sort_hist_kernel<<<b.n_ctas, kSortThreads, 0, st>>>(b, pass);
      HB_LAUNCH_CHECK();
      if (staged) sort_scatter_kernel<true><<<b.n_ctas, kSortThreads, 0, st>>>(b, pass);
      else sort_scatter_kernel<false><<<b.n_ctas, kSortThreads, 0, st>>>(b, pass);
      HB_LAUNCH_CHECK();
    }
    sort_settle_kernel<<<flat_grid, 256, 0, st>>>(b);
    HB_LAUNCH_CHECK();
  }
  return HDK_B200_OK;
}

int hdk_b200_sort_permutation(const int64_t* const* cols, const hdk_b200_order_entry* order, int n_order, uint64_t n_rows,
                              uint64_t top_n, uint32_t* permutation, uint64_t* n_sorted, void* scratch, size_t scratch_bytes,
                              void* stream) {
  if (n_sorted) *n_sorted = n_rows;
  if (!order || n_order < 1 || n_order > HDK_B200_MAX_TARGETS) { set_error("sort: bad order entry count"); return HDK_B200_E_INVALID; }
  if (n_rows >> 32) { set_error("sort: more than 2^32 rows"); return HDK_B200_E_UNSUPPORTED; }   // RowSortException, ResultSetSort.cpp:785-787
  if (n_rows == 0) return HDK_B200_OK;
  if (!cols || !permutation || !scratch) { set_error("sort: null argument"); return HDK_B200_E_INVALID; }
  if (scratch_bytes < hdk_b200_sort_scratch_bytes(n_rows)) { set_error("sort: scratch too small"); return HDK_B200_E_NOMEM; }
  for (int i = 0; i < n_order; ++i) {
    const hdk_b200_order_entry& oe = order[i];
    if (oe.column < 0 || oe.column >= HDK_B200_MAX_TARGETS || !cols[oe.column]) { set_error("sort: bad order column"); return HDK_B200_E_INVALID; }
    if (oe.type_width != 1 && oe.type_width != 2 && oe.type_width != 4 && oe.type_width != 8) { set_error("sort: bad type width"); return HDK_B200_E_INVALID; }
    if (oe.is_fp && oe.dict_rank) { set_error("sort: dictionary rank on a floating-point target"); return HDK_B200_E_INVALID; }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SortBufs b{};
  uint8_t* p = static_cast<uint8_t*>(scratch);
  b.keys[0] = reinterpret_cast<uint64_t*>(p); p += align256(n_rows * 8);
  b.keys[1] = reinterpret_cast<uint64_t*>(p); p += align256(n_rows * 8);
  b.idx[0] = permutation;
  b.idx[1] = reinterpret_cast<uint32_t*>(p); p += align256(n_rows * 4);
  b.hist = reinterpret_cast<uint32_t*>(p); p += align256(size_t(256) * sort_ctas(n_rows) * 4);
  b.state = reinterpret_cast<SortState*>(p); p += align256(sizeof(SortState));
  SelectState* sel = reinterpret_cast<SelectState*>(p);
  b.n = n_rows;
  bool identity = true;
  if (top_n && top_n <= n_rows / 8 && n_rows >= (1u << 16)) {
    // LIMIT prefilter on the primary target: keys → radix select → candidate rows in idx[0]
    const hdk_b200_order_entry& oe = order[0];
    KeyArgs k{cols[oe.column], oe.dict_rank, oe.dict_size, oe.is_fp, oe.type_width, oe.nullable, oe.is_desc, oe.nulls_first};
    const int flat_grid = int(std::min<uint64_t>(148 * 8, (n_rows + 255) / 256));
    HB_CUDA(cudaMemsetAsync(&b.state->key_or, 0x00, 8, st));
    HB_CUDA(cudaMemsetAsync(&b.state->key_and, 0xff, 8, st));
    HB_CUDA(cudaMemsetAsync(sel, 0, sizeof(SelectState), st));
    HB_CUDA(cudaMemcpyAsync(&sel->k, &top_n, 8, cudaMemcpyHostToDevice, st));
    sort_keys_kernel<<<flat_grid, 256, 0, st>>>(k, b, 1);
    HB_LAUNCH_CHECK();
    const unsigned long long stop_at = std::max<unsigned long long>(4 * top_n, 1u << 15);
    for (int pass = 7; pass >= 0; --pass) {
      select_hist_kernel<<<flat_grid, 256, 0, st>>>(b.keys[0], n_rows, b.state, sel, pass);
      HB_LAUNCH_CHECK();
      select_pick_kernel<<<1, 32, 0, st>>>(b.state, sel, pass, top_n, stop_at);
      HB_LAUNCH_CHECK();
    }
    select_compact_kernel<<<flat_grid, 256, 0, st>>>(b.keys[0], n_rows, sel, b.idx[0]);
    HB_LAUNCH_CHECK();
    unsigned long long m = 0;   // the candidate count sizes the sort's grids: the one synchronisation of this call
    HB_CUDA(cudaMemcpyAsync(&m, &sel->n_candidates, 8, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaStreamSynchronize(st));
    if (m < top_n || m > n_rows) { set_error("sort: LIMIT prefilter kept %llu of %llu rows for LIMIT %llu", m, (unsigned long long)n_rows, (unsigned long long)top_n); return HDK_B200_E_CUDA; }
    b.n = m;
    identity = false;
    if (n_sorted) *n_sorted = m;
  }
  return sort_rows_lsd(b, cols, order, n_order, identity, st);
}

int hdk_b200_gather_rows(const int64_t* const* cols_in, int64_t* const* cols_out, int n_cols, const uint32_t* permutation,
                         uint64_t n_out, void* stream) {
  if (n_cols < 1 || n_cols > HDK_B200_MAX_TARGETS) { set_error("gather: bad column count"); return HDK_B200_E_INVALID; }
  if (n_out == 0) return HDK_B200_OK;
  if (!cols_in || !cols_out || !permutation) { set_error("gather: null argument"); return HDK_B200_E_INVALID; }
  GatherArgs a{};
  for (int c = 0; c < n_cols; ++c) {
    if (!cols_in[c] || !cols_out[c]) { set_error("gather: null column"); return HDK_B200_E_INVALID; }
    a.in[c] = cols_in[c];
    a.out[c] = cols_out[c];
  }
  a.n_cols = n_cols;
  a.perm = permutation;
  a.n_out = n_out;
  gather_rows_kernel<<<int(std::min<uint64_t>(148 * 8, (n_out + 255) / 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

}  // extern "C"
