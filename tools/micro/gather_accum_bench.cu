// microbenchmark (round 2): what bounds config 5 (star join + group-by SUM(fp64)) on sm_100a —
//  (1) random 4-byte / 2-byte gathers from an L2-resident dimension payload (LDG variants, 16-byte bulk copies),
//  (2) fp64 SUM into ~1000 shared-memory bins: CTA table + atomicAdd(double) (CAS loop) vs per-warp tables with
//      plain read-modify-write and in-warp duplicate handling (match_any ranks / tag write + read back / none).
// Reports cycles per row per SM; the HBM roof of config 5 (12 B/row at 6537 GB/s) is 0.53 cycles per row per SM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// ---------------- gathers ----------------
template <int MODE, int U>
__global__ void __launch_bounds__(256) gather_k(const void* table, uint32_t n_entries, uint32_t iters, unsigned long long* out) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned long long acc = 0;
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { s += 0x9e3779b9u; idx[u] = uint32_t((uint64_t(mix(s)) * n_entries) >> 32); }
    uint32_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (MODE == 0) v[u] = __ldg(reinterpret_cast<const uint32_t*>(table) + idx[u]);
      else if (MODE == 1) v[u] = __ldg(reinterpret_cast<const uint16_t*>(table) + idx[u]);
      else if (MODE == 2) v[u] = __ldcg(reinterpret_cast<const uint32_t*>(table) + idx[u]);
      else if (MODE == 3) { asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v[u]) : "l"(reinterpret_cast<const uint32_t*>(table) + idx[u])); }
      else if (MODE == 4) v[u] = __ldg(reinterpret_cast<const uint8_t*>(table) + idx[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  if (acc == 0x123456789ull) out[0] = acc;
}

// 16-byte bulk copies (UBLKCP) as the gather: each lane fetches the 16-byte piece holding its element into its own
// shared-memory slot; one mbarrier per warp counts the bytes
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
template <int U>
__global__ void __launch_bounds__(256) gather_bulk_k(const uint32_t* table, uint32_t n_entries, uint32_t iters, unsigned long long* out) {
  __shared__ __align__(16) uint32_t slot[8][U][32][4];
  __shared__ __align__(8) uint64_t bar[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[warp]))); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, phase = 0;
  unsigned long long acc = 0;
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { s += 0x9e3779b9u; idx[u] = uint32_t((uint64_t(mix(s)) * n_entries) >> 32); }
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[warp])), "r"(uint32_t(U * 32 * 16)) : "memory");
    __syncwarp();
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(smem_u32(&slot[warp][u][lane][0])),
                   "l"(table + (idx[u] & ~3u)), "r"(smem_u32(&bar[warp])) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar[warp])), "r"(phase) : "memory");
    phase ^= 1;
#pragma unroll
    for (int u = 0; u < U; ++u) acc += slot[warp][u][lane][idx[u] & 3u];
    __syncwarp();
  }
  if (acc == 0x123456789ull) out[0] = acc;
}

// ---------------- fp64 SUM into G bins ----------------
// MODE 0: one table per CTA, atomicAdd(double) in shared memory (CAS loop)
// MODE 1: one table per warp, match_any ranks, plain RMW
// MODE 2: one table per warp, tag write + read back picks one winner per bin per round
// MODE 3: one table per warp, duplicates ignored (WRONG results: the ceiling of the per-warp form)
// MODE 4: one table per CTA, match_any leaders pre-add their group's values (shuffles), one atomicAdd per distinct bin
// MODE 5: one table per warp, tags as in 2 but tags are 16-bit and keyed by (round-stamp | lane) so they never need clearing
template <int MODE>
__global__ void __launch_bounds__(256) accum_k(uint32_t G, uint32_t iters, double* out) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bool per_warp = MODE == 1 || MODE == 2 || MODE == 3 || MODE == 5;
  double* bins = reinterpret_cast<double*>(sm) + (per_warp ? size_t(warp) * G : 0);
  uint32_t* tags = reinterpret_cast<uint32_t*>(sm + size_t(nw) * G * 8) + size_t(warp) * G;
  if (per_warp) { for (uint32_t i = lane; i < G; i += 32) { bins[i] = 0.0; if (MODE == 2 || MODE == 5) tags[i] = 0xffffffffu; } }
  else { for (uint32_t i = threadIdx.x; i < G; i += blockDim.x) bins[i] = 0.0; }
  __syncthreads();
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
  for (uint32_t it = 0; it < iters; ++it) {
    s += 0x9e3779b9u;
    const uint32_t h = mix(s);
    const uint32_t idx = uint32_t((uint64_t(h) * G) >> 32);
    const double x = double(int32_t(h & 0xffff) - 32768) * 0.25;
    if (MODE == 0) {
      atomicAdd(&bins[idx], x);
    } else if (MODE == 1) {
      const uint32_t peers = __match_any_sync(0xffffffffu, idx);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      const uint32_t rounds = __reduce_max_sync(0xffffffffu, rank);
      for (uint32_t r = 0; r <= rounds; ++r) {
        if (rank == r) bins[idx] += x;
        __syncwarp();
      }
    } else if (MODE == 2) {
      bool pending = true;
      while (__any_sync(0xffffffffu, pending)) {
        if (pending) tags[idx] = lane;
        __syncwarp();
        if (pending && tags[idx] == uint32_t(lane)) { bins[idx] += x; pending = false; }
        __syncwarp();
      }
    } else if (MODE == 3) {
      bins[idx] += x;
      __syncwarp();
    } else if (MODE == 4) {
      const uint32_t peers = __match_any_sync(0xffffffffu, idx);
      double sum = x;
      // (uniform part) number of rounds = max popc(peers)
      const uint32_t maxp = __reduce_max_sync(0xffffffffu, __popc(peers));
      if (maxp > 1) {
        uint32_t m = peers;
        sum = 0.0;
        for (uint32_t r = 0; r < maxp; ++r) {
          const int src = m ? __ffs(m) - 1 : lane;
          const double v = __shfl_sync(0xffffffffu, x, src);
          if (m) { sum += v; m &= m - 1; }
        }
      }
      const bool leader = (peers & ((1u << lane) - 1u)) == 0;
      if (leader) atomicAdd(&bins[idx], sum);
    } else if (MODE == 5) {
      bool pending = true;
      uint32_t round = it << 8;
      while (__any_sync(0xffffffffu, pending)) {
        const uint32_t tag = round | uint32_t(lane);
        if (pending) tags[idx] = tag;
        __syncwarp();
        if (pending && tags[idx] == tag) { bins[idx] += x; pending = false; }
        __syncwarp();
        round += 32;
      }
    }
  }
  __syncthreads();
  // fold into the global result (for the correctness check)
  if (per_warp) { for (uint32_t i = lane; i < G; i += 32) atomicAdd(&out[i], bins[i]); }
  else { for (uint32_t i = threadIdx.x; i < G; i += blockDim.x) atomicAdd(&out[i], bins[i]); }
}

static double sm_clock_ghz = 1.965;

template <int MODE, int U>
void run_gather(const char* name, const void* table, uint32_t n_entries, int ctas_per_sm) {
  unsigned long long* out; CK(cudaMalloc(&out, 64));
  const uint32_t iters = 2000 / U;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather_k<MODE, U><<<148 * ctas_per_sm, 256>>>(table, n_entries, 50, out);
  CK(cudaEventRecord(e0));
  gather_k<MODE, U><<<148 * ctas_per_sm, 256>>>(table, n_entries, iters, out);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double rows = 148.0 * ctas_per_sm * 256 * iters * U;
  printf("gather %-44s U=%d ctas/SM=%d: %7.3f ms %7.1f G rows/s  %.3f cyc/row/SM\n", name, U, ctas_per_sm, ms, rows / ms / 1e6,
         ms * 1e-3 * sm_clock_ghz * 1e9 * 148 / rows);
  cudaFree(out);
}
template <int U>
void run_gather_bulk(const uint32_t* table, uint32_t n_entries, int ctas_per_sm) {
  unsigned long long* out; CK(cudaMalloc(&out, 64));
  const uint32_t iters = 1000 / U;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather_bulk_k<U><<<148 * ctas_per_sm, 256>>>(table, n_entries, 20, out);
  CK(cudaEventRecord(e0));
  gather_bulk_k<U><<<148 * ctas_per_sm, 256>>>(table, n_entries, iters, out);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double rows = 148.0 * ctas_per_sm * 256 * iters * U;
  printf("gather %-44s U=%d ctas/SM=%d: %7.3f ms %7.1f G rows/s  %.3f cyc/row/SM\n", "16-byte bulk copy per row (UBLKCP)", U, ctas_per_sm, ms,
         rows / ms / 1e6, ms * 1e-3 * sm_clock_ghz * 1e9 * 148 / rows);
  cudaFree(out);
}

template <int MODE>
void run_accum(const char* name, uint32_t G, int ctas_per_sm, const std::vector<double>* expect, std::vector<double>* result) {
  double* out; CK(cudaMalloc(&out, G * 8));
  const bool per_warp = MODE == 1 || MODE == 2 || MODE == 3 || MODE == 5;
  const size_t smem = per_warp ? size_t(8) * G * 8 + ((MODE == 2 || MODE == 5) ? size_t(8) * G * 4 : 0) : size_t(G) * 8;
  CK(cudaFuncSetAttribute(accum_k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const uint32_t iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  accum_k<MODE><<<148 * ctas_per_sm, 256, smem>>>(G, 50, out);
  CK(cudaMemset(out, 0, G * 8));
  CK(cudaEventRecord(e0));
  accum_k<MODE><<<148 * ctas_per_sm, 256, smem>>>(G, iters, out);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<double> h(G);
  CK(cudaMemcpy(h.data(), out, G * 8, cudaMemcpyDeviceToHost));
  const char* verdict = "";
  if (expect) {
    bool same = true;
    for (uint32_t i = 0; i < G; ++i) same = same && h[i] == (*expect)[i];   // quarter-integers: sums are exact
    verdict = same ? "== CTA-atomic sums" : "DIFFERENT sums";
  }
  if (result) *result = h;
  const double rows = 148.0 * ctas_per_sm * 256 * iters;
  printf("accum  %-44s G=%5u ctas/SM=%d smem=%6zu: %7.3f ms %7.1f G rows/s  %.3f cyc/row/SM  %s\n", name, G, ctas_per_sm, smem, ms,
         rows / ms / 1e6, ms * 1e-3 * sm_clock_ghz * 1e9 * 148 / rows, verdict);
  cudaFree(out);
}

int main() {
  const uint32_t n = 10'000'000;
  std::vector<uint32_t> h(n);
  for (uint32_t i = 0; i < n; ++i) h[i] = (i * 2654435761u) % 1000u;
  void* t32; CK(cudaMalloc(&t32, size_t(n) * 4 + 64));
  CK(cudaMemcpy(t32, h.data(), size_t(n) * 4, cudaMemcpyHostToDevice));
  for (int c : {2, 4, 8}) {
    run_gather<0, 4>("u32 __ldg (40 MB table)", t32, n, c);
    run_gather<0, 8>("u32 __ldg (40 MB table)", t32, n, c);
    run_gather<1, 8>("u16 __ldg (20 MB table)", t32, n, c);
    run_gather<4, 8>("u8 __ldg (10 MB table)", t32, n, c);
    run_gather<2, 8>("u32 __ldcg", t32, n, c);
    run_gather<3, 8>("u32 ld.nc.L1::no_allocate", t32, n, c);
  }
  run_gather<0, 8>("u32 __ldg (4 MB table)", t32, 1'000'000, 4);
  run_gather<0, 8>("u32 __ldg (400 KB table)", t32, 100'000, 4);
  run_gather<0, 8>("u32 __ldg (40 KB table: L1 hits)", t32, 10'000, 4);
  for (int c : {2, 4}) { run_gather_bulk<4>(reinterpret_cast<const uint32_t*>(t32), n, c); run_gather_bulk<8>(reinterpret_cast<const uint32_t*>(t32), n, c); }
  for (uint32_t G : {1000u, 100u}) {
    for (int c : {1, 2}) {
      std::vector<double> expect;
      run_accum<0>("CTA table, atomicAdd(double)", G, c, nullptr, &expect);
      run_accum<4>("CTA table, match_any leaders + atomicAdd", G, c, &expect, nullptr);
      run_accum<1>("per-warp tables, match_any ranks", G, c, &expect, nullptr);
      run_accum<2>("per-warp tables, tag + read back", G, c, &expect, nullptr);
      run_accum<5>("per-warp tables, stamped tags", G, c, &expect, nullptr);
      run_accum<3>("per-warp tables, duplicates ignored (wrong)", G, c, &expect, nullptr);
    }
  }
  return 0;
}
