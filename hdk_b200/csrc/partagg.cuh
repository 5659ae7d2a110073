// hdk_b200/csrc/partagg.cuh — interface of the radix-partitioned baseline-hash aggregation (partagg.cu).
#pragma once
#include "common.cuh"

namespace hb {

constexpr int kPaMaxFields = HDK_B200_MAX_KEYS + kMaxAcc;

// One field of a packed record: the value of an expression node (a group key after the cast to the key width, or an
// aggregate argument), 1 or 2 four-byte words.
struct PaField {
  int16_t expr;     // plan node whose value the field holds
  uint8_t words;    // 1: the value fits 32 bits (sign-extended on read); 2: 64 bits (int64, or the bits of a double)
  uint8_t off;      // first word inside the record
};

struct PaLayout {
  int32_t ok;                     // the plan can run partitioned (baseline hash, no joins, record fits)
  int32_t n_keys, n_fields, rec_words, key_words;
  PaField f[kPaMaxFields];        // keys first (the first key_words words of a record are its key)
  int8_t acc_field[kMaxAcc];      // field holding accumulator a's argument, -1 for COUNT(*)
};

// Record layout of a plan: a function of the plan's STRUCTURE only, so that pre-compiled shapes know it at compile time.
__host__ __device__ constexpr PaLayout pa_layout_of(const DPlan& p) {
  PaLayout L{};
  if (p.hash_type != HDK_B200_BASELINE_HASH || p.n_joins != 0 || p.n_keys < 1) return L;
  L.n_keys = p.n_keys;
  int off = 0, nf = 0;
  for (int k = 0; k < p.n_keys; ++k) {
    const DExpr& e = p.exprs[p.keys[k].expr];
    L.f[nf].expr = int16_t(p.keys[k].expr);
    L.f[nf].words = uint8_t(e.width == 8 ? 2 : 1);
    L.f[nf].off = uint8_t(off);
    off += L.f[nf].words;
    ++nf;
  }
  L.key_words = off;
  for (int a = 0; a < p.n_acc; ++a) {
    L.acc_field[a] = -1;
    if (p.accs[a].arg < 0) continue;
    for (int i = p.n_keys; i < nf; ++i)
      if (L.f[i].expr == p.accs[a].arg) L.acc_field[a] = int8_t(i);
    if (L.acc_field[a] >= 0) continue;
    if (nf >= kPaMaxFields) return L;
    const DExpr& e = p.exprs[p.accs[a].arg];
    L.f[nf].expr = int16_t(p.accs[a].arg);
    L.f[nf].words = uint8_t((e.kind == HDK_B200_FP || e.width == 8) ? 2 : 1);
    L.f[nf].off = uint8_t(off);
    off += L.f[nf].words;
    L.acc_field[a] = int8_t(nf++);
  }
  if (off > 64) return L;
  L.n_fields = nf;
  L.rec_words = off;
  L.ok = 1;
  return L;
}

int partagg_scratch_bytes(const Lowered& lw, uint64_t total_rows, size_t* bytes);
// Enqueues the partitioned path.  *fallback_flag (device int) is raised by it when a partition is too heavy for one CTA
// (hot keys): its remaining kernels then do nothing and the caller's global-table path, enqueued behind with
// run_if = *fallback_flag and its work table at *fallback_work (inside `scratch`), does the launch's work instead.
int launch_partagg(const Lowered& lw, const hdk_b200_kernel_params* params, void* scratch, size_t scratch_bytes, cudaStream_t st,
                   hdk_b200_launch_info* info, const int** fallback_flag, int64_t** fallback_work);

// process-wide debug / tuning overrides (hdk_b200_debug_set)
struct DebugKnobs {
  int force_generic;        // 1: never use a pre-compiled plan shape
  int force_strategy;       // -1, or an HDK_B200_STRATEGY_* the perfect-hash launch must use
  int partitioned;          // -1 auto, 0 never, 1 whenever the plan is eligible and the scratch area suffices
  int pa_slots;             // 0 auto, else slots of the per-CTA shared table of the partitioned aggregation (tests: force splits)
  int pa_partitions;        // 0 auto, else the number of partitions
  int pa_heavy_rows;        // 0 auto, else the partition size beyond which the launch falls back to the global-table probe
  int jit;                  // run-time specialisation of plan shapes without pre-compiled kernels: 0 off, 1 in the background, 2 before the launch
  int geo_env_refresh;      // 1: read HDK_B200_GEO on every launch instead of once per process (tools/sweep_geo.py)
};
extern DebugKnobs g_debug;

}  // namespace hb
