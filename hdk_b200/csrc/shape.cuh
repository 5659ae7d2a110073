// hdk_b200/csrc/shape.cuh — plan shapes: the compile-time structure of a lowered plan.
#pragma once
#ifndef __CUDACC_RTC__
#include <type_traits>
#include <utility>
#endif

#include "common.cuh"

namespace hb {

// ---- plan shapes -----------------------------------------------------------------------------
// The generic kernel interprets the device plan at run time.  For the plan shapes listed in
// static_shapes.inc (generated at build time from the named configs by tools/gen_static_shapes.py)
// the SAME row code is instantiated with the plan's structure as a compile-time constant, so the
// expression switch, type checks and accumulator dispatch fold away and `vals[]` lives in registers.
// Literal values, key ranges and entry counts stay run-time parameters in both cases.
struct GenericShape {
  static constexpr bool is_static = false;
  static constexpr int rows_per_iter = 1;
  __host__ __device__ static constexpr DPlan get() { return DPlan{}; }
};
template <int ID>
struct StaticShape;

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---- pre-compiled plan shapes ------------------------------------------------------------------
#define HB_STATIC_SHAPE(ID, SIG, NAME, RPI, ...)                                 \
  template <>                                                                    \
  struct StaticShape<ID> {                                                       \
    static constexpr bool is_static = true;                                      \
    static constexpr int rows_per_iter = RPI;                                    \
    __host__ __device__ static constexpr DPlan get() { return DPlan __VA_ARGS__; } \
  };
#ifndef HB_JIT   // (a run-time compiled translation unit defines its own single shape)
#include "static_shapes.inc"
#endif
#undef HB_STATIC_SHAPE

}  // namespace hb
