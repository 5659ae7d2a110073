/*
 * oracle/oracle_ref.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Same glue (driver.inc), but every per-row runtime function is the REFERENCE's own code:
 * the reference's RuntimeFunctions.cpp (which textually includes GroupByRuntime.cpp,
 * JoinHashTableQueryRuntime.cpp and DecodersImpl.h) is compiled as part of this
 * translation unit from where it lies under $(REF) — nothing is copied into the repo —
 * so the always_inline runtime inlines into the row loop much like the JIT links the
 * runtime bitcode into the generated row function.
 * Built only where the reference tree exists (see Makefile); output in oracle/_ref/.
 */
#include "QueryEngine/RuntimeFunctions.cpp"
#include "QueryEngine/MurmurHash.h"
#include "Utils/ExtractFromTime.h"
extern "C" int64_t extract_year(const int64_t timeval); /* Utils/ExtractFromTime.cpp:260; the JIT binds it by name */

/* The TBB-backed quantile collectors are not on this path; the runtime only references them. */
extern "C" {
void agg_quantile_impl_int8(int64_t*, int8_t) { abort(); }
void agg_quantile_impl_int16(int64_t*, int16_t) { abort(); }
void agg_quantile_impl_int32(int64_t*, int32_t) { abort(); }
void agg_quantile_impl_int64(int64_t*, int64_t) { abort(); }
void agg_quantile_impl_float(int64_t*, float) { abort(); }
void agg_quantile_impl_double(int64_t*, double) { abort(); }
}
#define ORACLE_KIND "reference"
#define ORACLE_KIND_FN oracle_kind
#include "driver.inc"
