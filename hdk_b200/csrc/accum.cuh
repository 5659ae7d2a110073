// hdk_b200/csrc/accum.cuh — accumulator primitives shared by the scan kernel (scan.cu) and the partitioned
// baseline-hash aggregation (partagg.cu): NULL test of an aggregate's argument, the NEUTRAL identities, the value a
// row contributes, and the update of a private / shared-memory / global cell.
// Reference semantics: agg_* and their _skip_val forms (QE/RuntimeFunctions.cpp:388-880), GPU shared-memory forms
// (QE/cuda_mapd_rt.cu:423-1083).
#pragma once
#include "common.cuh"
#include "eval.cuh"

namespace hb {

// Is the accumulator's argument NULL for this row?  mode 1: the argument's own sentinel; mode 2: the
// reference's COUNT(int64) quirk (see lower.cu)
__device__ __forceinline__ bool acc_arg_is_null(const DPlan& p, const DAcc& a, const V* vals) {
  if (!a.arg_nullable) return false;
  const DExpr& t = p.exprs[a.arg];
  const V v = vals[a.arg];
  if (t.kind == HDK_B200_FP) return v.f == fp_null_of(t.width);
  if (v.i == int_null_of(t.width)) return true;
  return a.arg_nullable == 2 && int32_t(v.i) == INT32_MIN;
}

__device__ __forceinline__ int64_t acc_identity(uint8_t kind) {
  return (kind == ACC_MIN_I || kind == ACC_MIN_F) ? INT64_MAX : (kind == ACC_MAX_I || kind == ACC_MAX_F) ? INT64_MIN : 0;
}

// value contributed by this row to accumulator `a` (as an int64 cell / double bits)
__device__ __forceinline__ int64_t acc_input(const DPlan& p, const DAcc& a, const V* vals) {
  switch (a.kind) {
    case ACC_CNT_ALL: case ACC_CNT_NN: return 1;
    case ACC_SUM_I: case ACC_MIN_I: case ACC_MAX_I: return vals[a.arg].i;
    case ACC_SUM_F: return vals[a.arg].i;  // bits of the double
    default: return f64_order_encode(vals[a.arg].f);  // MIN_F / MAX_F
  }
}

// ---------------------------------------------------------------------------------------------
// accumulator updates
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bin_update_private(uint8_t kind, uint8_t* bin, int64_t x) {
  switch (kind) {
    case ACC_CNT_ALL: case ACC_CNT_NN: *reinterpret_cast<uint32_t*>(bin) += 1u; break;
    case ACC_SUM_I: *reinterpret_cast<int64_t*>(bin) += x; break;
    case ACC_SUM_F: *reinterpret_cast<double*>(bin) += __longlong_as_double(x); break;
    case ACC_MIN_I: case ACC_MIN_F: { int64_t* b = reinterpret_cast<int64_t*>(bin); *b = min(*b, x); break; }
    default: { int64_t* b = reinterpret_cast<int64_t*>(bin); *b = max(*b, x); break; }
  }
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

__device__ __forceinline__ void bin_update_shared_atomic(uint8_t kind, uint8_t* bin, int64_t x) {
  switch (kind) {
    case ACC_CNT_ALL: case ACC_CNT_NN: atomicAdd(reinterpret_cast<uint32_t*>(bin), 1u); break;
    case ACC_SUM_I: smem_add_i64(bin, x); break;
    case ACC_SUM_F: atomicAdd(reinterpret_cast<double*>(bin), __longlong_as_double(x)); break;
    // 64-bit shared atomics are CAS loops (SASS ATOMS.CAST.SPIN.64): look first, a bin only ever moves towards x
    case ACC_MIN_I: case ACC_MIN_F:
      if (x < *reinterpret_cast<volatile int64_t*>(bin)) atomicMin(reinterpret_cast<long long*>(bin), static_cast<long long>(x));
      break;
    default:
      if (x > *reinterpret_cast<volatile int64_t*>(bin)) atomicMax(reinterpret_cast<long long*>(bin), static_cast<long long>(x));
      break;
  }
}
__device__ __forceinline__ void cell_update_global(uint8_t kind, int64_t* cell, int64_t x) {
  switch (kind) {
    case ACC_CNT_ALL: case ACC_CNT_NN: case ACC_SUM_I:
      atomicAdd(reinterpret_cast<unsigned long long*>(cell), static_cast<unsigned long long>(x)); break;
    case ACC_SUM_F: atomicAdd(reinterpret_cast<double*>(cell), __longlong_as_double(x)); break;
    case ACC_MIN_I: case ACC_MIN_F: atomicMin(reinterpret_cast<long long*>(cell), static_cast<long long>(x)); break;
    default: atomicMax(reinterpret_cast<long long*>(cell), static_cast<long long>(x)); break;
  }
}

}  // namespace hb
