// hdk_b200/csrc/common.cuh — shared device/host definitions of the sm_100a hot-path library.
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation (jit.cu): no host headers; fixed-width types and limits from libcu++
#include <cuda/std/cstdint>
#include <cuda/std/climits>
#include <cuda/std/type_traits>
#include <cuda/std/utility>
using cuda::std::int8_t; using cuda::std::int16_t; using cuda::std::int32_t; using cuda::std::int64_t;
using cuda::std::uint8_t; using cuda::std::uint16_t; using cuda::std::uint32_t; using cuda::std::uint64_t; using cuda::std::uintptr_t;
namespace std { using cuda::std::integral_constant; using cuda::std::conditional; }
#define HDK_B200_NO_STD_HEADERS 1
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "../../include/hdk_b200.h"

namespace hb {

// ---------------------------------------------------------------------------------------
// error handling (host)
// ---------------------------------------------------------------------------------------
#ifndef __CUDACC_RTC__
void set_error(const char* fmt, ...);
extern unsigned long long g_launch_count;
#define HB_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::hb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return HDK_B200_E_CUDA;                                                            \
    }                                                                                    \
  } while (0)
#define HB_LAUNCH_CHECK()                                \
  do {                                                   \
    ++::hb::g_launch_count;                              \
    HB_CUDA(cudaGetLastError());                         \
  } while (0)

int sm_count();
#endif

// ---------------------------------------------------------------------------------------
// device plan: the C-ABI plan lowered to a compact, kernel-parameter-sized form
// ---------------------------------------------------------------------------------------
constexpr uint8_t kAuxInQual = 0x80;   // errors of such a node are raised whether or not the row passes the filter

struct DExpr {          // 16 bytes
  uint8_t op;           // hdk_b200_op
  int8_t a, b;          // operand nodes (OP_COL: table, column)
  uint8_t aux;          // bit 0: the ABI's aux (date-in-days column / checked arithmetic); bit 7 (kAuxInQual): part of a filter qual
  uint8_t kind, width, nullable;  // result type
  uint8_t guard;        // 0, or 1 + the node that must be true for this one to be evaluated (inside a CASE arm)
  union {
    int64_t i;
    double f;
  } imm;
};

enum AccKind : uint8_t {
  ACC_CNT_ALL = 0,  // rows of the group
  ACC_CNT_NN = 1,   // rows whose argument is not NULL
  ACC_SUM_I = 2,
  ACC_SUM_F = 3,
  ACC_MIN_I = 4,    // MIN over int64 (also order-encoded doubles)
  ACC_MAX_I = 5,
  ACC_MIN_F = 6,    // MIN over doubles, stored order-encoded → merged with integer MIN
  ACC_MAX_F = 7
};

struct DAcc {           // 4 bytes
  uint8_t kind;         // AccKind
  int8_t arg;           // expr node, -1 for CNT_ALL
  uint8_t arg_nullable; // skip rows whose argument is NULL
  uint8_t bytes;        // bin width inside shared memory: 4 (counters) or 8
};

struct DKey {           // 48 bytes
  int64_t min_val;
  int64_t null_translated;  // max + (bucket ? bucket : 1)
  int64_t mult;             // Π cardinality of the previous keys
  int64_t card;             // this key's bucketed cardinality (incl. the NULL bin)
  int32_t expr;
  uint8_t has_nulls, width, pad0, pad1;
  int64_t bucket;           // > 1: the key is (value - min) / bucket (ExpressionRange buckets: a DATE key counts days, 86400 s each)
};

struct DJoin {          // 48 bytes
  int64_t min_key, max_key, null_val;
  int32_t key_expr;     // perfect: the key node; baseline: the last component in node order (probe point)
  uint8_t key_nullable, one_to_many, by_slot, pad1;   // by_slot: presence bitmap + slot-ordered inner columns (run-time, not structure)
  // baseline join table (composite / wide-range keys): n_key_exprs > 0
  uint8_t n_key_exprs, key_width, pad2, pad3;
  int8_t key_exprs[HDK_B200_MAX_KEYS];
  int32_t pad4;
};

constexpr int kMaxAcc = 28;

struct DPlan {
  int32_t n_exprs, n_filters, n_keys, n_joins, n_acc, n_cols;
  uint32_t entry_count;
  int32_t hash_type;
  DExpr exprs[HDK_B200_MAX_EXPRS];      // 768 B
  int8_t filters[HDK_B200_MAX_FILTERS];
  DKey keys[HDK_B200_MAX_KEYS];         // 256 B
  DJoin joins[HDK_B200_MAX_JOINS];      // 128 B
  DAcc accs[kMaxAcc];                   // 112 B
  uint8_t col_width[HDK_B200_MAX_COLS]; // physical width of outer column c
  int64_t join_entry_count[HDK_B200_MAX_JOINS];
};

// Per-slot recipe used by finalize / baseline kernels: how to produce the reference's slot
// encoding from accumulators (perfect hash) or how to update it in place (baseline).
enum SlotOp : uint8_t {
  SLOT_KEY = 0,       // agg_id of group key `key_index`
  SLOT_COUNT = 1,     // COUNT(*) / COUNT(arg) / AVG count
  SLOT_SUM = 2,       // SUM / AVG sum
  SLOT_MIN = 3,
  SLOT_MAX = 4
};

struct DSlot {          // 32 bytes
  uint8_t op;           // SlotOp
  uint8_t bytes;        // agg_chosen_bytes actually written: 4 or 8 (float aggregate in an 8-byte slot: 4)
  uint8_t padded;       // padded slot width (stride contribution)
  uint8_t is_fp;        // value is floating point (double when bytes==8, float when 4)
  uint8_t skip_null;    // TargetInfo.skip_null_val
  uint8_t is_avg_sum;   // AVG's sum slot: init 0 even when nullable
  int8_t acc;           // accumulator feeding this slot (perfect hash), -1 none
  int8_t acc_cnt;       // accumulator telling "any non-null row" (nullable SUM/MIN/MAX), -1 none
  int32_t off;          // row-wise: byte offset inside the slot area; columnar: unused
  int32_t key_index;    // SLOT_KEY
  int16_t arg;          // expr node feeding the slot (baseline path), -1 = none (COUNT(*))
  uint8_t arg_kind, arg_width, arg_nullable;   // type of that node
  uint8_t key_width, key_nullable;             // SLOT_KEY: logical type of the key
  uint8_t count_mode;   // SLOT_COUNT over a nullable arg: 1 plain, 2 = int32-truncation quirk (see lower.cu)
  int64_t init_val;     // qmd.init_vals[slot]: doubles as the skip value (null sentinel)
  uint64_t col_off;     // columnar: byte offset of the slot column
};

struct DLayout {
  uint32_t entry_count;
  int32_t key_count, key_width, keyless, columnar, slot_count, target_idx_for_key;
  uint32_t row_bytes;     // row-wise
  uint32_t key_bytes;     // row-wise aligned key part
  DSlot slots[HDK_B200_MAX_SLOTS];  // 1 KB
};

struct Lowered {
  DPlan plan;
  DLayout layout;
  // accumulator classes for the multi-GPU merge (cells = n * entry_count)
  int n_sum_i, n_sum_f, n_min, n_max;   // accs are ordered: [sum_i | sum_f | min | max]
  size_t work_table_bytes;
  size_t stage_row_bytes;  // Σ physical widths of the outer columns
};

#ifndef __CUDACC_RTC__
int lower_plan(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, Lowered* out);
uint64_t plan_signature(const DPlan& p);
#endif

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int64_t int_null_of(int w) {
  return w == 1 ? int64_t(INT8_MIN) : w == 2 ? int64_t(INT16_MIN) : w == 4 ? int64_t(INT32_MIN) : INT64_MIN;
}
__host__ __device__ __forceinline__ double fp_null_of(int w) {
  return w == 4 ? double(1.17549435e-38f) : 2.2250738585072014e-308;
}
__host__ __device__ __forceinline__ int64_t resize_int(int64_t v, int w) {
  return w == 1 ? int64_t(int8_t(v)) : w == 2 ? int64_t(int16_t(v)) : w == 4 ? int64_t(int32_t(v)) : v;
}
// order-preserving map double → int64 (so fp MIN/MAX merge with integer min/max everywhere)
__host__ __device__ __forceinline__ int64_t f64_order_encode(double d) {
  int64_t b;
#ifdef __CUDA_ARCH__
  b = __double_as_longlong(d);
#else
  memcpy(&b, &d, 8);
#endif
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__host__ __device__ __forceinline__ double f64_order_decode(int64_t e) {
  int64_t b = e ^ ((e >> 63) & 0x7fffffffffffffffLL);
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}

// record_error_code semantics (QE/RuntimeFunctions.cpp:1123-1135): a positive (persistent)
// code is never overwritten; a negative (out of slots) code only replaces 0 / negative.
__device__ __forceinline__ void record_error(int32_t* err, int32_t code) {
  if (!code) return;
  int32_t old = *reinterpret_cast<volatile int32_t*>(err);
  while (old <= 0) {
    const int32_t prev = atomicCAS(err, old, code);
    if (prev == old) break;
    old = prev;
  }
}

}  // namespace hb
