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
  int8_t col;       // plain outer column feeding the field, or -1 (needs the interpreter)
  uint8_t colw;     // physical width of that column
  uint8_t is_fp;
  uint8_t pad;
};

struct PaLayout {
  int32_t n_keys, n_fields, rec_words, key_words, key_width;
  int32_t direct;                 // no filters, every field is a plain outer column, few fields: no interpreter
  PaField f[kPaMaxFields];        // keys first (the first key_words words of a record are its key)
  int8_t acc_field[kMaxAcc];      // field holding accumulator a's argument, -1 for COUNT(*)
};

int partagg_layout(const Lowered& lw, PaLayout* out);
int partagg_scratch_bytes(const Lowered& lw, uint64_t total_rows, size_t* bytes);
int launch_partagg(const Lowered& lw, const hdk_b200_kernel_params* params, void* scratch, size_t scratch_bytes, cudaStream_t st,
                   hdk_b200_launch_info* info);

// process-wide debug / tuning overrides (hdk_b200_debug_set)
struct DebugKnobs {
  int force_generic;        // 1: never use a pre-compiled plan shape
  int force_strategy;       // -1, or an HDK_B200_STRATEGY_* the perfect-hash launch must use
  int partitioned;          // -1 auto, 0 never, 1 whenever the plan is eligible and the scratch area suffices
  int pa_slots;             // 0 auto, else slots of the per-CTA shared table of the partitioned aggregation (tests: force splits)
  int pa_partitions;        // 0 auto, else the number of partitions
};
extern DebugKnobs g_debug;

}  // namespace hb
