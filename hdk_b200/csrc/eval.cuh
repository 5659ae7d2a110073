// hdk_b200/csrc/eval.cuh — per-row evaluation of the plan's expression DAG on the device.
// Semantics follow the reference's generated row function and the helpers it calls
// (QE/DecodersImpl.h, QE/RuntimeFunctions.cpp:45-384, QE/CastIR.cpp, QE/DateTimeIR.cpp,
// omniscidb/Utils/ExtractFromTime.cpp); every case cites its source.
#pragma once
#include "common.cuh"

namespace hb {

// ---------------------------------------------------------------------------------------------
// per-row expression evaluation
// ---------------------------------------------------------------------------------------------
union V {
  int64_t i;
  double f;
};

__device__ __forceinline__ bool v_is_null(const DExpr& t, const V& v) {
  if (!t.nullable) return false;
  return t.kind == HDK_B200_FP ? (v.f == fp_null_of(t.width)) : (v.i == int_null_of(t.width));
}
__device__ __forceinline__ V v_null(const DExpr& t) {
  V v;
  if (t.kind == HDK_B200_FP) v.f = fp_null_of(t.width); else v.i = int_null_of(t.width);
  return v;
}

// omniscidb/Utils/ExtractFromTime.cpp:156-163 (fast path) and :260-271
__device__ __forceinline__ int64_t dev_extract_year(int64_t t) {
  const uint32_t kEpochOffsetYear1900 = 2208988800u, kSecsJanToMar1900 = 5097600u;
  if (t >= 0 && t <= int64_t(0xFFFFFFFFu - kEpochOffsetYear1900)) {
    const uint32_t s1900 = uint32_t(t) + kEpochOffsetYear1900;
    const uint32_t leap = (s1900 - kSecsJanToMar1900) / 126230400u;
    return (s1900 - leap * 86400u) / 31536000u + 1900;
  }
  const int64_t day = (t < 0 ? t - 86399 : t) / 86400;
  const int64_t d2 = day - 11017;
  const int64_t era = (d2 < 0 ? d2 - 146096 : d2) / 146097;
  const unsigned doe = unsigned(d2 - era * 146097);
  const unsigned yoe = (doe - doe / 1460 + doe / 36524 - (doe == 146096)) / 365;
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  return 2000 + era * 400 + yoe + (306u <= doy);
}

template <class LoadOuter, class LoadInner>
__device__ __forceinline__ V eval_node(const DPlan& p, const DExpr& e, const V* vals, int32_t& err,
                                       LoadOuter&& load_outer, LoadInner&& load_inner) {
  V r;
  r.i = 0;
  if (e.guard && !(vals[e.guard - 1].i > 0)) return r;  // inside a CASE arm this row does not take (QE/CaseIR.cpp:66-93)
  switch (e.op) {
    case HDK_B200_OP_COL: {
      // fixed_width_{int,float,double,small_date}_decode (QE/DecodersImpl.h:31-161)
      uint64_t raw = e.a == 0 ? load_outer(e.b, int(e.imm.i)) : load_inner(e.a - 1, e.b, int(e.imm.i));
      const int pw = int(e.imm.i);
      if (e.kind == HDK_B200_FP) {
        r.f = pw == 4 ? double(__uint_as_float(uint32_t(raw))) : __longlong_as_double(int64_t(raw));
      } else {
        int64_t v = pw == 1 ? int64_t(int8_t(raw)) : pw == 2 ? int64_t(int16_t(raw)) : pw == 4 ? int64_t(int32_t(raw)) : int64_t(raw);
        if (e.aux & 1) v = (v == int_null_of(pw)) ? int_null_of(e.width) : v * 86400;
        r.i = v;
      }
      break;
    }
    case HDK_B200_OP_CONST: r.i = e.imm.i; break;
    case HDK_B200_OP_ADD:
    case HDK_B200_OP_SUB:
    case HDK_B200_OP_MUL:
    case HDK_B200_OP_DIV: {
      const DExpr &ta = p.exprs[e.a], &tb = p.exprs[e.b];
      const V a = vals[e.a], b = vals[e.b];
      if (v_is_null(ta, a) || v_is_null(tb, b)) { r = v_null(e); break; }
      if (e.kind == HDK_B200_FP) {
        if (e.op == HDK_B200_OP_DIV && b.f == 0.0) { if (e.aux & 2) r = v_null(e); else err = HDK_B200_ERR_DIV_BY_ZERO; break; }
        if (e.width == 4) {
          const float x = float(a.f), y = float(b.f);
          r.f = double(e.op == HDK_B200_OP_ADD ? x + y : e.op == HDK_B200_OP_SUB ? x - y : e.op == HDK_B200_OP_MUL ? x * y : x / y);
        } else {
          r.f = e.op == HDK_B200_OP_ADD ? a.f + b.f : e.op == HDK_B200_OP_SUB ? a.f - b.f : e.op == HDK_B200_OP_MUL ? a.f * b.f : a.f / b.f;
        }
      } else {
        int64_t lo;
        bool ovf = false;
        if (e.op == HDK_B200_OP_DIV) {
          if (b.i == 0) { if (e.aux & 2) r = v_null(e); else err = HDK_B200_ERR_DIV_BY_ZERO; break; }
          lo = (a.i == INT64_MIN && b.i == -1) ? INT64_MIN : a.i / b.i;
          ovf = (a.i == INT64_MIN && b.i == -1);
        } else if (e.op == HDK_B200_OP_ADD) {
          lo = int64_t(uint64_t(a.i) + uint64_t(b.i));
          ovf = ((a.i ^ lo) & (b.i ^ lo)) < 0;
        } else if (e.op == HDK_B200_OP_SUB) {
          lo = int64_t(uint64_t(a.i) - uint64_t(b.i));
          ovf = ((a.i ^ b.i) & (a.i ^ lo)) < 0;
        } else {
          lo = int64_t(uint64_t(a.i) * uint64_t(b.i));
          const int64_t hi = __mul64hi(a.i, b.i);
          ovf = hi != (lo >> 63);
        }
        if ((e.aux & 1) && (ovf || resize_int(lo, e.width) != lo)) { err = HDK_B200_ERR_OVERFLOW_OR_UNDERFLOW; break; }
        r.i = resize_int(lo, e.width);
      }
      break;
    }
    case HDK_B200_OP_UMINUS: {
      const DExpr& ta = p.exprs[e.a];
      if (v_is_null(ta, vals[e.a])) { r = v_null(e); break; }
      if (e.kind == HDK_B200_FP) { r.f = -vals[e.a].f; break; }
      // operand == type minimum raises (codegenUMinus, QE/ArithmeticIR.cpp:782-811); a nullable minimum was NULL above
      if (vals[e.a].i == int_null_of(e.width)) { err = HDK_B200_ERR_OVERFLOW_OR_UNDERFLOW; break; }
      r.i = resize_int(-vals[e.a].i, e.width);
      break;
    }
    case HDK_B200_OP_CAST: {
      const DExpr& ta = p.exprs[e.a];
      const V a = vals[e.a];
      if (v_is_null(ta, a)) { r = v_null(e); break; }
      if (ta.kind == HDK_B200_FP && e.kind == HDK_B200_INT) {
        // round half away from zero then truncate (QE/CastIR.cpp:529-541, RuntimeFunctions.cpp:309-345)
        r.i = resize_int(int64_t(a.f + (a.f < 0 ? -0.5 : 0.5)), e.width);
      } else if (ta.kind == HDK_B200_INT && e.kind == HDK_B200_FP) {
        r.f = e.width == 4 ? double(float(a.i)) : double(a.i);
      } else if (e.kind == HDK_B200_FP) {
        r.f = e.width == 4 ? double(float(a.f)) : a.f;
      } else {
        // narrowing integer cast: v > max or v <= min (= the NULL sentinel) of the target raises (QE/CastIR.cpp:405-462)
        if (ta.width > e.width) {
          const int64_t mx = (int64_t(1) << (8 * e.width - 1)) - 1;
          if (a.i > mx || a.i <= -mx - 1) { err = HDK_B200_ERR_OVERFLOW_OR_UNDERFLOW; break; }
        }
        r.i = resize_int(a.i, e.width);
      }
      break;
    }
    case HDK_B200_OP_EXTRACT_YEAR: {
      const DExpr& ta = p.exprs[e.a];
      int64_t t = vals[e.a].i;
      const int64_t units = e.imm.i;
      // Fast path for times inside extract_year's own fast range [0, 2^32 - 2208988800) s (ExtractFromTime.cpp:262),
      // which also excludes NULL (every sentinel is negative).  Days since the epoch come from one fp64 fma:
      // t < 2^41 is exact, (t + 1/2) / D is at least 1/(2 D) > 5e-9 away from an integer and the rounding error is
      // < 1e-11, so truncation is exact; between 1901 and 2099 every fourth year is leap, so
      // year = 1970 + (4 days + 2) / 1461.  Checked against the reference formula for every second of the range
      // and every day boundary in ms (tests/test_gpu_parity.py::test_extract_year_boundaries).
      if (units <= 1000 && uint64_t(t) < uint64_t(2085978496) * uint64_t(units)) {
        const double inv = 1.0 / (86400.0 * double(units));
        const uint32_t days = __double2uint_rz(fma(double(t), inv, 0.5 * inv));
        r.i = 1970 + (4u * days + 2u) / 1461u;
        break;
      }
      if (ta.nullable && t == int_null_of(ta.width)) { r = v_null(e); break; }
      if (units > 1) t = ta.nullable ? ((t < 0 ? t - (units - 1) : t) / units) : t / units;  // QE/DateTimeIR.cpp:314-320
      r.i = dev_extract_year(t);
      break;
    }
    case HDK_B200_OP_LT: case HDK_B200_OP_LE: case HDK_B200_OP_GT:
    case HDK_B200_OP_GE: case HDK_B200_OP_EQ: case HDK_B200_OP_NE: {
      const DExpr &ta = p.exprs[e.a], &tb = p.exprs[e.b];
      const V a = vals[e.a], b = vals[e.b];
      if (v_is_null(ta, a) || v_is_null(tb, b)) { r.i = INT8_MIN; break; }
      bool t;
      if (ta.kind == HDK_B200_FP) {
        t = e.op == HDK_B200_OP_LT ? a.f < b.f : e.op == HDK_B200_OP_LE ? a.f <= b.f : e.op == HDK_B200_OP_GT ? a.f > b.f
            : e.op == HDK_B200_OP_GE ? a.f >= b.f : e.op == HDK_B200_OP_EQ ? a.f == b.f : a.f != b.f;
      } else {
        t = e.op == HDK_B200_OP_LT ? a.i < b.i : e.op == HDK_B200_OP_LE ? a.i <= b.i : e.op == HDK_B200_OP_GT ? a.i > b.i
            : e.op == HDK_B200_OP_GE ? a.i >= b.i : e.op == HDK_B200_OP_EQ ? a.i == b.i : a.i != b.i;
      }
      r.i = t;
      break;
    }
    case HDK_B200_OP_AND: {  // logical_and, QE/RuntimeFunctions.cpp:362-372
      const int64_t l = vals[e.a].i, rr = vals[e.b].i, nul = INT8_MIN;
      r.i = l == nul ? (rr == 0 ? rr : nul) : rr == nul ? (l == 0 ? l : nul) : ((l && rr) ? 1 : 0);
      break;
    }
    case HDK_B200_OP_OR: {  // logical_or, :374-384
      const int64_t l = vals[e.a].i, rr = vals[e.b].i, nul = INT8_MIN;
      r.i = l == nul ? (rr == 0 ? nul : rr) : rr == nul ? (l == 0 ? nul : l) : ((l || rr) ? 1 : 0);
      break;
    }
    case HDK_B200_OP_NOT: {
      const int64_t o = vals[e.a].i;
      r.i = o == INT8_MIN ? o : (o ? 0 : 1);
      break;
    }
    case HDK_B200_OP_IS_NULL: {
      DExpr t = p.exprs[e.a];
      t.nullable = 1;
      r.i = v_is_null(t, vals[e.a]);
      break;
    }
    case HDK_B200_OP_CASE:  // toBool(when) ? then : else, QE/CaseIR.cpp:68-111; both arms already have the node's type
      r = vals[e.a].i > 0 ? vals[e.b] : vals[int(e.imm.i)];
      break;
    default: err = 1000; break;
  }
  return r;
}


}  // namespace hb
