/*
 * hdk_b200.h — C ABI of the B200-native replacement for HDK's fused
 * scan → filter → (join probe) → group-by → aggregate hot path.
 *
 * Everything here is plain C: pointers, sizes, PODs.  No torch / C++ types cross
 * this boundary.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference tree, QE = omniscidb/QueryEngine,
 * JHT = QE/JoinHashTable).
 *
 * All `*_on_device` / `launch` / `reduce` entry points take DEVICE pointers and a
 * `cudaStream_t` passed as `void*` (NULL = legacy default stream, which is what
 * the reference launches on: QE/DeviceKernel.cpp:54-85).  They are asynchronous
 * with respect to the host unless stated otherwise.  Return value: 0 on success,
 * a negative HDK_B200_E_* code on a host-side failure (bad plan, CUDA error).
 * Query-time errors are reported in-band through the `error_codes` device array,
 * with the reference's convention (QE/QueryExecutionContext.cpp:221-236,
 * QE/RuntimeFunctions.cpp:1123-1135): >0 persistent error, <0 out of slots.
 */
#ifndef HDK_B200_H
#define HDK_B200_H

#ifndef HDK_B200_NO_STD_HEADERS /* (the library's own run-time compiled kernels get these types from libcu++) */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define HDK_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define HDK_B200_API __attribute__((visibility("default")))
#else
#define HDK_B200_API
#endif

/* ---- limits of the plan POD ------------------------------------------------ */
#define HDK_B200_MAX_EXPRS 48
#define HDK_B200_MAX_KEYS 8
#define HDK_B200_MAX_TARGETS 24
#define HDK_B200_MAX_SLOTS 32
#define HDK_B200_MAX_FILTERS 8
#define HDK_B200_MAX_JOINS 4
#define HDK_B200_MAX_COLS 32

/* ---- host-side error codes ------------------------------------------------- */
#define HDK_B200_OK 0
#define HDK_B200_E_INVALID (-1)     /* malformed plan / descriptor                    */
#define HDK_B200_E_UNSUPPORTED (-2) /* plan shape outside the supported set: caller   */
                                    /* must fail the query (never a CPU fallback)     */
#define HDK_B200_E_CUDA (-3)        /* CUDA runtime error; see hdk_b200_last_error()  */
#define HDK_B200_E_NOMEM (-4)

/* ---- in-band (device) error codes: QE/Execute.h:1019-1031 ------------------ */
#define HDK_B200_ERR_DIV_BY_ZERO 1
#define HDK_B200_ERR_OUT_OF_SLOTS 3 /* stored negated (<0), see record_error_code */
#define HDK_B200_ERR_OVERFLOW_OR_UNDERFLOW 7
#define HDK_B200_ERR_SINGLE_VALUE_FOUND_MULTIPLE_VALUES 15

/* ---- sentinels: Shared/InlineNullValues.h:33-39, QE/GpuRtConstants.h:29-32 -- */
#define HDK_B200_EMPTY_KEY_64 INT64_C(9223372036854775807)
#define HDK_B200_EMPTY_KEY_32 2147483647

/* ============================================================================
 * Value types.  A value travelling through a plan is either a 64-bit integer or
 * a double in a 64-bit register, like the reference's row function widens every
 * decoded column (QE/DecodersImpl.h:31-60, :135).  `width` is the LOGICAL byte
 * width of the SQL type; it selects the null sentinel:
 *   int  1/2/4/8 → INT8_MIN / INT16_MIN / INT32_MIN / INT64_MIN
 *   fp   4/8     → FLT_MIN / DBL_MIN
 * ==========================================================================*/
enum hdk_b200_kind { HDK_B200_INT = 0, HDK_B200_FP = 1 };

typedef struct hdk_b200_type {
  int8_t kind;     /* hdk_b200_kind */
  int8_t width;    /* logical bytes: 1,2,4,8 */
  int8_t nullable; /* 0/1 */
  int8_t pad;
} hdk_b200_type;

/* ============================================================================
 * Expression DAG, topologically ordered: node i may only reference nodes < i.
 * This is the subset of hdk::ir::Expr that the named plan shapes need
 * (QE/ColumnIR.cpp, ArithmeticIR.cpp, CompareIR.cpp, LogicalIR.cpp, CastIR.cpp,
 * DateTimeIR.cpp).  Null semantics follow the reference's *_nullable helpers
 * (QE/RuntimeFunctions.cpp:45-212): an arithmetic result is NULL if any nullable
 * operand is NULL; a comparison yields NULL (treated as false by a filter).
 * ==========================================================================*/
enum hdk_b200_op {
  HDK_B200_OP_COL = 0,     /* a = table (0 = outer/fact, j>0 = inner table of join j-1), b = column index; */
                           /* ival = physical byte width in the chunk; type = logical type;                */
                           /* aux: 0 plain fixed-width, 1 = date-in-days encoding → seconds                */
                           /* (fixed_width_small_date_decode, QE/DecodersImpl.h:153-161)                   */
  HDK_B200_OP_CONST = 1,   /* ival (INT) or fval (FP) */
  HDK_B200_OP_ADD = 2,
  HDK_B200_OP_SUB = 3,
  HDK_B200_OP_MUL = 4,
  HDK_B200_OP_DIV = 5,     /* ERR_DIV_BY_ZERO on a zero divisor; with aux bit 1 (value 2) NULL instead:       */
                           /* safe_div_* under Config null_div_by_zero (QE/ArithmeticIR.cpp:587-597)        */
  HDK_B200_OP_CAST = 6,    /* a → type.  fp→int rounds half away from zero (QE/CastIR.cpp:529-541,        */
                           /* RuntimeFunctions.cpp:309-345); int→fp exact convert; int→int re-sentinels;    */
                           /* narrowing int→int raises ERR_OVERFLOW_OR_UNDERFLOW outside (min, max] of the  */
                           /* target (codegenCastBetweenIntTypesOverflowChecks, CastIR.cpp:405-462)         */
  HDK_B200_OP_EXTRACT_YEAR = 7, /* a = timestamp in seconds after ival-division: ival = units per second   */
                           /* (1, 1000, 1e6, 1e9) (QE/DateTimeIR.cpp:281-320, Utils/ExtractFromTime.cpp:260-271) */
  HDK_B200_OP_LT = 8,
  HDK_B200_OP_LE = 9,
  HDK_B200_OP_GT = 10,
  HDK_B200_OP_GE = 11,
  HDK_B200_OP_EQ = 12,
  HDK_B200_OP_NE = 13,
  HDK_B200_OP_AND = 14,    /* 3-valued, QE/RuntimeFunctions.cpp:362-384 */
  HDK_B200_OP_OR = 15,
  HDK_B200_OP_NOT = 16,
  HDK_B200_OP_IS_NULL = 17,
  HDK_B200_OP_UMINUS = 18,
  HDK_B200_OP_CASE = 19    /* a = WHEN condition, b = THEN value, ival = ELSE value node: b if a is true (> 0, a NULL */
                           /* condition is not), else the ELSE node — CodeGenerator::codegenCase, QE/CaseIR.cpp:51-113. */
                           /* Both values already carry the node's type; further WHEN arms nest in the ELSE node.     */
};

typedef struct hdk_b200_expr {
  int32_t op;   /* hdk_b200_op */
  int32_t a;    /* operand node index, or table for OP_COL */
  int32_t b;    /* operand node index, or column index for OP_COL */
  int32_t aux;
  hdk_b200_type type; /* result type */
  int32_t guard; /* 0 = always evaluated.  g > 0: the node sits inside a CASE arm and is evaluated only for rows where
                  * node g-1 is true; elsewhere it raises no error and its value is never selected.  The reference gets
                  * this from the basic blocks codegenCase emits (QE/CaseIR.cpp:66-93): a division in a THEN arm cannot
                  * fail for rows that take another arm. */
  int64_t ival;
  double fval;
} hdk_b200_expr;

/* ============================================================================
 * Targets: QE/TargetExprBuilder.cpp:42-80 (agg_fn_base_names), Shared/TargetInfo.h
 * ==========================================================================*/
enum hdk_b200_agg {
  HDK_B200_AGG_NONE = 0, /* projected group key → agg_id */
  HDK_B200_AGG_COUNT = 1,
  HDK_B200_AGG_SUM = 2,
  HDK_B200_AGG_MIN = 3,
  HDK_B200_AGG_MAX = 4,
  HDK_B200_AGG_AVG = 5  /* two slots: {agg_sum, agg_count} */
};

typedef struct hdk_b200_target {
  int32_t agg;           /* hdk_b200_agg */
  int32_t arg;           /* expr node index, -1 for COUNT(*) */
  hdk_b200_type type;    /* TargetInfo.type  (result type; AVG over ints: int64) */
  hdk_b200_type arg_type;/* TargetInfo.agg_arg_type (unused for COUNT(*) / AGG_NONE) */
  int32_t skip_null_val; /* TargetInfo.skip_null_val */
  int32_t key_index;     /* AGG_NONE: index of the group key this target projects; else -1 */
  int32_t slot;          /* first slot index in the ColSlotContext; -1 = no slot (baseline: key target) */
  int32_t pad;
} hdk_b200_target;

/* ============================================================================
 * Query memory descriptor subset: omniscidb/ResultSet/QueryMemoryDescriptor.{h,cpp},
 * ColSlotContext.cpp, QE/MemoryLayoutBuilder.cpp:795-994.  It fully determines the
 * byte layout of the group-by buffer (QueryMemoryDescriptor.cpp:240-256, 290-342):
 *   row-wise : E × [ key_count × key_width, padded to 8 | slot0 | slot1 … ] padded to 8
 *              (8-byte slots are 8-aligned, ColSlotContext.cpp:143-158)
 *   keyless  : no key part (perfect hash only)
 *   columnar : [key col 8B × E]…  then per slot [padded × E, 8-aligned]
 * ==========================================================================*/
enum hdk_b200_hash_type { HDK_B200_PERFECT_HASH = 0, HDK_B200_BASELINE_HASH = 1 };

typedef struct hdk_b200_qmd {
  int32_t hash_type;        /* hdk_b200_hash_type */
  int32_t keyless;          /* hasKeylessHash() */
  int32_t target_idx_for_key; /* getTargetIdxForKey(): slot index telling an empty entry when keyless */
  int32_t output_columnar;  /* didOutputColumnar() */
  uint32_t entry_count;
  int32_t key_count;        /* getGroupbyColCount() */
  int32_t key_width;        /* getEffectiveKeyWidth(): 8 for perfect hash, 4/8 baseline */
  int32_t slot_count;
  int64_t min_val;          /* single-column perfect hash only */
  int64_t max_val;
  int64_t bucket;
  int32_t has_nulls;
  int32_t pad;
  int8_t slot_padded[HDK_B200_MAX_SLOTS];  /* getPaddedSlotWidthBytes(i): 0,4,8 */
  int8_t slot_logical[HDK_B200_MAX_SLOTS];
  int64_t init_vals[HDK_B200_MAX_SLOTS];   /* init_agg_val_vec, QE/OutputBufferInitialization.cpp:30-68 */
} hdk_b200_qmd;

/* Per group key: QE/RowFuncBuilder.cpp:447-478, :748-801 */
typedef struct hdk_b200_key {
  int32_t expr;          /* node index */
  int32_t has_nulls;     /* translate NULL → max+bucket (perfect hash only) */
  int64_t min_val;
  int64_t max_val;
  int64_t bucket;        /* 0 = none */
  int64_t cardinality;   /* getBucketedCardinality(): (max-min)/(bucket?bucket:1) + 1 + has_nulls */
} hdk_b200_key;

/* Equi-join against a perfect join hash table built by hdk_b200_fill_hash_join_buff_on_device
 * (probe = hash_join_idx[_nullable], QE/GroupByRuntime.cpp:298-329).  Inner join: a
 * row without a match is dropped.  The matched inner row id addresses the inner
 * table's columns (OP_COL with a = join index + 1). */
typedef struct hdk_b200_join {
  int32_t key_expr;      /* outer-side key node */
  int32_t one_to_many;   /* 0: OneToOne int32[entries]; 1: offsets|counts|payload (JHT/PerfectJoinHashTable.cpp:861-886) */
  int64_t min_key;
  int64_t max_key;
  int64_t null_val;      /* outer key null sentinel; used when key nullable */
  int32_t key_nullable;
  /* extension (OneToOne only).  1: JOIN_HASH_TABLES[j] is the presence bitmap and inner_col_buffers[j][*] are the
   * slot-ordered copies made by hdk_b200_gather_join_payload_on_device, so a probe costs one random access per
   * referenced column instead of table + column; results are identical.  2: as 1, and the caller knows that every
   * slot of [min_key, max_key] is occupied (as many distinct non-NULL keys as entries), so the bitmap is not read. */
  int32_t payload_by_slot;
  int64_t entry_count;   /* hash entries (for one_to_many buffer offsets; baseline: entries of the table) */
  /* Composite / wide-range equi-joins probe a BASELINE join table (JHT/BaselineJoinHashTable.cpp) in its one-to-one
   * layout E x (key components ‖ payload row id), every cell key_width (4 or 8) bytes, MurmurHash1 + linear probing
   * (write_baseline_hash_slot, JHT/Runtime/HashJoinRuntime.cpp:438-471; probe baseline_hash_join_idx_{32,64},
   * JHT/Runtime/JoinHashTableQueryRuntime.cpp:43-98).  n_key_exprs = 0: perfect table on `key_expr` (above).
   * n_key_exprs >= 1: `key_exprs` are the outer-side component nodes in key order and `key_expr` is the LAST of them
   * in node order (the probe happens once it has been evaluated); min_key / max_key / null_val are unused — rows with
   * a NULL component were never inserted by the build, so they simply miss. */
  int32_t n_key_exprs;
  int32_t key_width;
  int32_t key_exprs[HDK_B200_MAX_KEYS];
} hdk_b200_join;

typedef struct hdk_b200_plan {
  int32_t abi_version;
  int32_t n_exprs;
  int32_t n_filters;
  int32_t n_keys;
  int32_t n_targets;
  int32_t n_joins;
  int32_t n_cols;        /* outer table physical columns passed per fragment */
  int32_t pad;
  hdk_b200_expr exprs[HDK_B200_MAX_EXPRS];
  int32_t filters[HDK_B200_MAX_FILTERS];  /* node indices; row passes iff all are TRUE (not NULL) */
  hdk_b200_key keys[HDK_B200_MAX_KEYS];
  hdk_b200_target targets[HDK_B200_MAX_TARGETS];
  hdk_b200_join joins[HDK_B200_MAX_JOINS];
} hdk_b200_plan;

/* ============================================================================
 * Kernel launch.  Replaces the JIT'd multifrag_query_hoisted_literals
 * (QE/RuntimeFunctions.cpp:1692-1726) launched by NvidiaKernel::launch
 * (QE/DeviceKernel.cpp:54-85) from QueryExecutionContext::launchGpuCode
 * (QE/QueryExecutionContext.cpp:238-550).  The parameter block is the
 * reference's 12-slot kernel-param vector (QE/QueryExecutionContext.h:112-126)
 * as a POD; every pointer is a DEVICE pointer except where noted.
 * ==========================================================================*/
typedef struct hdk_b200_kernel_params {
  const int8_t* const* col_buffers; /* COL_BUFFERS: device array [num_fragments * n_cols] of chunk pointers */
  uint64_t num_fragments;           /* NUM_FRAGMENTS (by value) */
  const int8_t* literals;           /* LITERALS: unused (constants live in the plan) */
  const int64_t* num_rows;          /* NUM_ROWS: device array [num_fragments] (outer table) */
  const uint64_t* frag_row_offsets; /* FRAG_ROW_OFFSETS: device array [num_fragments] */
  int32_t max_matched;              /* MAX_MATCHED (by value) */
  int32_t* total_matched;           /* TOTAL_MATCHED: unused for group-by */
  const int64_t* init_agg_vals;     /* INIT_AGG_VALS: unused (qmd.init_vals) */
  int64_t* const* groupby_buf;      /* GROUPBY_BUF: device array of buffer pointers; [0] is used */
  int32_t* error_codes;             /* ERROR_CODE: device int32[>=1]; slot 0 receives the aggregate code */
  uint32_t num_tables;              /* NUM_TABLES */
  const int64_t* join_hash_tables;  /* JOIN_HASH_TABLES: device array [n_joins] of table addresses */
  /* -- extension: inner-table columns for joins, device array [n_joins * HDK_B200_MAX_COLS] */
  const int8_t* const* inner_col_buffers;
  /* -- extension: sum of NUM_ROWS when the host knows it (the reference fills NUM_ROWS from a host vector,
   *    QE/QueryExecutionContext.cpp:789-964), 0 = unknown.  Only sizes tiles so that small inputs still give
   *    every resident CTA several tiles, and the record area of the partitioned baseline-hash path (a launch that finds
   *    more rows than the hint falls back to the per-row probe); results never depend on it. */
  uint64_t total_rows_hint;
} hdk_b200_kernel_params;

/* KernelOptions, QE/DeviceKernel.h:33-43.  grid/block of 0 = library picks
 * (persistent grid: a multiple of the SM count). */
typedef struct hdk_b200_kernel_options {
  unsigned int gridDimX, gridDimY, gridDimZ;
  unsigned int blockDimX, blockDimY, blockDimZ;
  unsigned int sharedMemBytes;
  unsigned int literalsOffset;
  int hoistLiterals;
} hdk_b200_kernel_options;

/* Filled by hdk_b200_launch when non-NULL (host struct, written synchronously at
 * enqueue time): which pre-compiled kernel variant ran and how many launches. */
typedef struct hdk_b200_launch_info {
  int32_t variant;       /* HDK_B200_VARIANT_* */
  int32_t strategy;      /* HDK_B200_STRATEGY_* */
  int32_t n_launches;    /* kernels enqueued by this call */
  int32_t grid, block, smem_bytes;
  int32_t n_accumulators;
  int32_t tile_rows;     /* rows per staged tile */
} hdk_b200_launch_info;

#define HDK_B200_VARIANT_JIT 1000 /* hdk_b200_launch_info.variant: kernels specialised for the plan's shape at run time */

enum hdk_b200_strategy {
  HDK_B200_STRATEGY_THREAD_PRIVATE = 0, /* per-thread bins in shared memory, no atomics */
  HDK_B200_STRATEGY_CTA_SHARED = 1,     /* per-CTA table in shared memory, shared atomics */
  HDK_B200_STRATEGY_GLOBAL = 2,         /* perfect hash straight into the global work table */
  HDK_B200_STRATEGY_BASELINE = 3,       /* open addressing in the global group-by buffer */
  HDK_B200_STRATEGY_REGISTER = 4,       /* <= 8 groups, pre-compiled shapes: every thread keeps all groups' accumulators in registers */
  HDK_B200_STRATEGY_PARTITIONED = 5     /* baseline hash, large tables: rows radix-partitioned by key hash, every partition      */
                                        /* aggregated in shared memory, groups written straight into the reference layout        */
};

/* Validate a plan/descriptor pair and report the scratch (device) bytes the
 * launch needs.  Host-only, no CUDA calls. */
HDK_B200_API int hdk_b200_plan_check(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, size_t* scratch_bytes);

/* Scratch bytes hdk_b200_launch wants for `total_rows` input rows (sum of NUM_ROWS; 0 = unknown): at least what
 * hdk_b200_plan_check reports.  Large baseline-hash group-bys (QE/RelAlgExecutor.cpp:691-838 is the reference's own,
 * CPU-only, partitioned aggregation) run radix-partitioned when the caller provides this much scratch and states
 * params->total_rows_hint: one packed record per row (keys + aggregate arguments) is staged in the scratch area.
 * With less scratch the launch falls back to probing the global table per row — same results, slower. */
HDK_B200_API int hdk_b200_launch_scratch_bytes(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, uint64_t total_rows,
                                               size_t* scratch_bytes);

/* Bytes of the group-by buffer for this descriptor:
 * QueryMemoryDescriptor::getBufferSizeBytes (QueryMemoryDescriptor.cpp:457-481). */
HDK_B200_API size_t hdk_b200_buffer_size_bytes(const hdk_b200_qmd* qmd);

/* Fill a group-by buffer with EMPTY_KEY / init values for this descriptor.
 * Replaces QueryMemoryInitializer::initRowGroups / initColumnarGroups
 * (QE/QueryMemoryInitializer.cpp:502-687) and init_group_by_buffer_gpu /
 * init_columnar_group_by_buffer_gpu (QE/GpuInitGroups.cu:120-188). */
HDK_B200_API int hdk_b200_init_group_by_buffer(const hdk_b200_qmd* qmd, int64_t* groups_buffer, void* stream);

/* 1:1 mirrors of the reference's host-callable initialisers, same argument meaning
 * (QE/GpuInitGroups.h:29-52); block/grid of 0 = library picks. */
HDK_B200_API int hdk_b200_init_group_by_buffer_on_device(int64_t* groups_buffer,
                                            const int64_t* init_vals, /* DEVICE */
                                            uint32_t groups_buffer_entry_count,
                                            uint32_t key_count,
                                            uint32_t key_width,
                                            uint32_t row_size_quad,
                                            int keyless,
                                            int8_t warp_size,
                                            size_t block_size_x,
                                            size_t grid_size_x,
                                            void* stream);
HDK_B200_API int hdk_b200_init_columnar_group_by_buffer_on_device(int64_t* groups_buffer,
                                                     const int64_t* init_vals, /* DEVICE */
                                                     uint32_t groups_buffer_entry_count,
                                                     uint32_t key_count,
                                                     uint32_t agg_col_count,
                                                     const int8_t* col_sizes, /* DEVICE */
                                                     int need_padding,
                                                     int keyless,
                                                     int8_t key_size,
                                                     size_t block_size_x,
                                                     size_t grid_size_x,
                                                     void* stream);

/* The fused scan → filter → join probe → group-by → aggregate launch over all
 * fragments of one device.  `groupby_buf[0]` must have been initialised (above).
 * `scratch` is a device scratch area of at least the size hdk_b200_plan_check
 * reported (may be NULL when that is 0).  The result buffer is byte-compatible
 * with `qmd` so the reference's ResultSet can iterate it. */
HDK_B200_API int hdk_b200_launch(const hdk_b200_plan* plan,
                    const hdk_b200_qmd* qmd,
                    const hdk_b200_kernel_options* ko, /* may be NULL */
                    const hdk_b200_kernel_params* params,
                    void* scratch,
                    size_t scratch_bytes,
                    void* stream,
                    hdk_b200_launch_info* info /* may be NULL */);

/* Multi-GPU split of the same launch for perfect-hash plans:
 *   1. hdk_b200_launch_partial: scan this device's fragments into a NEUTRAL
 *      columnar work table (int64/f64 cells: SUM→0, MIN→+max, MAX→−max, counts)
 *      of hdk_b200_work_table_cells(plan,qmd) 8-byte cells, laid out so that
 *      cells [0, n_sum) reduce with SUM, [n_sum, n_sum+n_min) with MIN and the
 *      rest with MAX  → three NCCL all-reduces (or one per op) merge ranks;
 *   2. hdk_b200_finalize: convert the merged work table into the reference
 *      encoding (keys, null sentinels, compact widths) inside groupby_buf. */
typedef struct hdk_b200_work_table_layout {
  uint64_t n_cells;     /* total 8-byte cells */
  uint64_t sum_cells;   /* [0, sum_cells): merge with SUM; int64 cells first, then f64 */
  uint64_t sum_i64_cells;
  uint64_t min_cells;   /* next min_cells: merge with MIN (int64 order; f64 stored order-preserving) */
  uint64_t max_cells;   /* last max_cells: merge with MAX */
} hdk_b200_work_table_layout;

HDK_B200_API int hdk_b200_work_table_layout_get(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd,
                                   hdk_b200_work_table_layout* out);
HDK_B200_API int hdk_b200_init_work_table(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd,
                             int64_t* work_table, void* stream);
HDK_B200_API int hdk_b200_launch_partial(const hdk_b200_plan* plan,
                            const hdk_b200_qmd* qmd,
                            const hdk_b200_kernel_options* ko,
                            const hdk_b200_kernel_params* params, /* groupby_buf unused */
                            int64_t* work_table,
                            void* stream,
                            hdk_b200_launch_info* info);
HDK_B200_API int hdk_b200_finalize(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd,
                      const int64_t* work_table, int64_t* groups_buffer, void* stream);

/* ResultSetReduction::reduce (QE/ResultSetReduction.cpp:174-330): merge `that`
 * into `this` on the device, both in `qmd`'s layout (`that_entry_count` may be
 * smaller for baseline hash, :196-201).  Perfect hash: slot-wise reduceOneSlot
 * (:1234-1320).  Baseline: re-insert every non-empty entry (:696-760); running out
 * of slots sets *error_codes to a negative code (ReductionRanOutOfSlots). */
HDK_B200_API int hdk_b200_reduce(const hdk_b200_plan* plan,
                    const hdk_b200_qmd* qmd,
                    int64_t* this_buffer,
                    const int64_t* that_buffer,
                    uint32_t that_entry_count,
                    int32_t* error_codes,
                    void* stream);

/* ============================================================================
 * Join hash tables.  1:1 mirrors of the reference's host-callable device builders
 * (JHT/Runtime/HashJoinRuntime.h:66-300, JHT/Runtime/HashJoinRuntimeGpu.cu) with an
 * added stream.  JoinColumn / JoinColumnTypeInfo keep the reference's fields
 * (HashJoinRuntime.h:99-131).
 * ==========================================================================*/
enum hdk_b200_column_type { HDK_B200_SMALL_DATE = 0, HDK_B200_SIGNED = 1, HDK_B200_UNSIGNED = 2, HDK_B200_DOUBLE = 3 };

typedef struct hdk_b200_join_chunk {
  const int8_t* col_buff; /* DEVICE */
  size_t num_elems;
  size_t row_id;          /* row id of the chunk's first element */
} hdk_b200_join_chunk;

typedef struct hdk_b200_join_column {
  const int8_t* col_chunks_buff; /* DEVICE array of hdk_b200_join_chunk */
  size_t col_chunks_buff_sz;
  size_t num_chunks;
  size_t num_elems;
  size_t elem_sz;
} hdk_b200_join_column;

typedef struct hdk_b200_join_column_type_info {
  size_t elem_sz;
  int64_t min_val;
  int64_t max_val;
  int64_t null_val;
  int uses_bw_eq;
  int64_t translated_null_val;
  int column_type; /* hdk_b200_column_type */
} hdk_b200_join_column_type_info;

/* init_hash_join_buff_on_device (HashJoinRuntime.h:72-74) */
HDK_B200_API int hdk_b200_init_hash_join_buff_on_device(int32_t* buff, int64_t entry_count,
                                           int32_t invalid_slot_val, void* stream);
/* fill_hash_join_buff_on_device[_bucketized] (HashJoinRuntime.h:163-183; body
 * HashJoinRuntime.cpp:198-296).  `dev_err_buff`: device int, set to -1 on a
 * duplicate key (→ NeedsOneToManyHash, JHT/Builders/PerfectHashTableBuilder.h:96-141).
 * bucket_normalization = 1 for the non-bucketized form. */
HDK_B200_API int hdk_b200_fill_hash_join_buff_on_device(int32_t* buff, int32_t invalid_slot_val, int for_semi_join,
                                           int* dev_err_buff,
                                           const hdk_b200_join_column* join_column,
                                           const hdk_b200_join_column_type_info* type_info,
                                           int64_t bucket_normalization, void* stream);
/* fill_one_to_many_hash_table_on_device[_bucketized] (HashJoinRuntime.h:236-254;
 * HashJoinRuntimeGpu.cu:134-236): buff = offsets[E] | counts[E] | payload[num_elems]. */
HDK_B200_API int hdk_b200_fill_one_to_many_hash_table_on_device(int32_t* buff, int64_t hash_entry_count,
                                                   int32_t invalid_slot_val,
                                                   const hdk_b200_join_column* join_column,
                                                   const hdk_b200_join_column_type_info* type_info,
                                                   int64_t bucket_normalization, void* stream);
/* init/fill_baseline_hash_join_buff_on_device_{32,64} (HashJoinRuntime.h:76-97, 185-234;
 * body HashJoinRuntime.cpp:298-576): composite-key open addressing with MurmurHash1,
 * entry = key_component_count keys (+ one payload slot when with_val_slot).
 * key_width = 4 or 8 selects the _32 / _64 form. */
HDK_B200_API int hdk_b200_init_baseline_hash_join_buff_on_device(int8_t* hash_join_buff, int64_t entry_count,
                                                    size_t key_component_count, int with_val_slot,
                                                    int32_t invalid_slot_val, int key_width, void* stream);
HDK_B200_API int hdk_b200_fill_baseline_hash_join_buff_on_device(int8_t* hash_buff, int64_t entry_count,
                                                    int32_t invalid_slot_val, int for_semi_join,
                                                    size_t key_component_count, int with_val_slot,
                                                    int* dev_err_buff,
                                                    const hdk_b200_join_column* join_columns,       /* HOST array [key_component_count] */
                                                    const hdk_b200_join_column_type_info* type_infos, /* HOST array */
                                                    int key_width, void* stream);
/* fill_one_to_many_baseline_hash_table_on_device_{32,64} (HashJoinRuntime.h:256-300):
 * buff = offsets[E] | counts[E] | payload[num_elems] over the composite-key dictionary. */
HDK_B200_API int hdk_b200_fill_one_to_many_baseline_hash_table_on_device(int32_t* buff, const int8_t* composite_key_dict,
                                                            int64_t hash_entry_count, int32_t invalid_slot_val,
                                                            size_t key_component_count,
                                                            const hdk_b200_join_column* join_columns,
                                                            const hdk_b200_join_column_type_info* type_infos,
                                                            int key_width, void* stream);
/* Stand-alone probes, vectorised over a key column (used by tests and by callers
 * that materialise join results): hash_join_idx (QE/GroupByRuntime.cpp:298-308) and
 * baseline_hash_join_idx_{32,64} (JHT/Runtime/JoinHashTableQueryRuntime.cpp:43-98).
 * out[i] = matching slot value / entry index or -1. */
HDK_B200_API int hdk_b200_probe_hash_join_on_device(const int32_t* buff, const int64_t* keys, int64_t n,
                                       int64_t min_key, int64_t max_key, int64_t* out, void* stream);

/* Extension: re-order one inner-table column by hash slot for a OneToOne perfect table (the layout the reference's
 * fill_hash_join_buff_on_device produces, JHT/Runtime/HashJoinRuntime.cpp fill_hash_join_buff):
 *   out_by_slot[slot] = table[slot] >= 0 ? inner_col[table[slot]] : 0      (elem_width 1, 2, 4 or 8 bytes)
 *   present_bitmap bit slot = table[slot] >= 0                            (uint32 words, may be NULL)
 * The probe of a plan whose join has payload_by_slot = 1 then reads the bitmap (entry_count / 8 bytes, cache
 * resident) and one element of each referenced column per row. */
HDK_B200_API int hdk_b200_gather_join_payload_on_device(const int32_t* hash_table, int64_t entry_count, const int8_t* inner_col,
                                                        int elem_width, int8_t* out_by_slot, uint32_t* present_bitmap,
                                                        void* stream);
HDK_B200_API int hdk_b200_probe_baseline_hash_join_on_device(const int8_t* hash_buff, const int8_t* keys /* n × key_component_count × key_width */,
                                                int64_t n, size_t key_component_count, int key_width,
                                                int64_t entry_count, int with_val_slot, int64_t* out, void* stream);

/* ============================================================================
 * Partitioned aggregation shuffle (model: QE/RelAlgExecutor.cpp:691-838,
 * partition function MurmurHash64A over 64-bit-widened keys & (P-1)…
 * QE/RowFuncBuilder.cpp:516-577; here floor(upper 32 bits * n_partitions / 2^32) so that
 * any GPU count works without a modulo).  Two passes over the fragments of one device:
 *   pass 1: hdk_b200_shuffle_count  → counts[n_partitions] (device uint64)
 *   pass 2: hdk_b200_shuffle_scatter → rows written partition-contiguous into
 *           `out_cols[c]` (one device array per plan column, outer table only)
 *           at offsets[p] + running index.
 * The NCCL all-to-all between the passes' outputs is the caller's
 * (torch.distributed) job.
 * ==========================================================================*/
HDK_B200_API int hdk_b200_shuffle_count(const hdk_b200_plan* plan, const hdk_b200_kernel_params* params,
                           uint32_t n_partitions, uint64_t* counts /* DEVICE, zeroed by callee */, void* stream);
HDK_B200_API int hdk_b200_shuffle_scatter(const hdk_b200_plan* plan, const hdk_b200_kernel_params* params,
                             uint32_t n_partitions,
                             const uint64_t* offsets /* DEVICE [n_partitions] exclusive prefix */,
                             uint64_t* cursors /* DEVICE [n_partitions], zeroed by callee */,
                             int8_t* const* out_cols /* DEVICE array [n_cols] */, void* stream);
/* Locality pass for large baseline-hash group-bys on ONE GPU (same two passes, different partition function): rows are
 * regrouped by the REGION of the group-by table their key hashes into (region = floor(slot * n_regions / entry_count),
 * slot = key_hash % entry_count as in get_group_value, QE/GroupByRuntime.cpp:31-54).  Aggregating the regrouped rows in
 * order then touches one L2-sized region of the table at a time instead of a random 32-byte sector of a multi-GB table
 * per row.  Results are unchanged (same hash, same probing, same table).  Filters are applied by the passes. */
HDK_B200_API int hdk_b200_region_count(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_params* params,
                                       uint32_t n_regions, uint64_t* counts /* DEVICE [n_regions] */, void* stream);
HDK_B200_API int hdk_b200_region_scatter_to(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_params* params,
                                            uint32_t n_regions, int8_t* const* dest_cols /* DEVICE [n_regions * n_cols] */,
                                            const uint64_t* dest_offsets /* DEVICE [n_regions] */,
                                            uint64_t* cursors /* DEVICE [n_regions], zeroed by callee */, void* stream);

/* pass 2 with one destination per partition: partition p's rows go to dest_cols[p * n_cols + c] (DEVICE array of DEVICE
 * pointers) starting at row dest_offsets[p].  The pointers may address PEER memory (hdk_b200_peer_alloc / _open): the
 * scatter kernel then IS the all-to-all — each GPU writes its rows straight into the owners' receive buffers through
 * NVLink, no staging copy and no collective for the payload (counts are exchanged beforehand to derive the offsets;
 * a barrier after the kernels makes the rows visible to the owners). */
HDK_B200_API int hdk_b200_shuffle_scatter_to(const hdk_b200_plan* plan, const hdk_b200_kernel_params* params, uint32_t n_partitions,
                                             int8_t* const* dest_cols /* DEVICE [n_partitions * n_cols] */,
                                             const uint64_t* dest_offsets /* DEVICE [n_partitions] */,
                                             uint64_t* cursors /* DEVICE [n_partitions], zeroed by callee */, void* stream);


/* ============================================================================
 * Multi-GPU merge of perfect-hash partials over peer memory (NVLink / NVSwitch), one process per GPU.
 * The reference merges per-device result sets on the host (Executor::reduceMultiDeviceResultSets,
 * QE/Execute.cpp:1224-1336); here the scan kernel itself publishes this GPU's neutral partial table:
 * the last CTA to finish copies it into slot `my_rank` of EVERY peer's exchange buffer with plain
 * 16-byte stores over NVLink and raises a flag there; the finalize kernel of each GPU waits for the
 * n_peers flags of the current epoch, merges the slots (SUM / MIN / MAX per accumulator) and writes
 * the reference-encoded buffer.  No NCCL call and no host synchronisation on the data path.
 *
 *   exchange buffer (one per rank and plan, peer-visible):
 *     flags[2][HDK_B200_MAX_PEERS] uint64 | parity 0: n_peers slots x cells int64 | parity 1: the same
 *   `epoch` (1, 2, 3, ...) is the caller's launch counter for this plan, equal on all ranks; its parity
 *   selects the half, so a fast rank can publish epoch e+1 while a slow one still merges epoch e.
 *
 * hdk_b200_peer_* wrap cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle so that the buffers can be
 * shared between the per-GPU processes (handles travel through the host-side process group).
 * ==========================================================================*/
#define HDK_B200_MAX_PEERS 16
#define HDK_B200_IPC_HANDLE_BYTES 64
#define HDK_B200_ERR_PEER_TIMEOUT 1004 /* a peer's flag did not arrive (in-band error code, positive = persistent) */
#define HDK_B200_ERR_CLAIM_TIMEOUT 1005 /* baseline hash: a claimed entry's key was never published (bounded wait of the claim protocol) */
HDK_B200_API int hdk_b200_peer_alloc(size_t bytes, void** dev_ptr, uint8_t handle[HDK_B200_IPC_HANDLE_BYTES]);
HDK_B200_API int hdk_b200_peer_open(const uint8_t handle[HDK_B200_IPC_HANDLE_BYTES], void** dev_ptr);
HDK_B200_API int hdk_b200_peer_close(void* dev_ptr);
HDK_B200_API int hdk_b200_peer_free(void* dev_ptr);
HDK_B200_API int hdk_b200_exchange_bytes(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, int n_peers, size_t* bytes);
/* zero the flags of a local exchange buffer (once per buffer; barrier across ranks before the first launch) */
HDK_B200_API int hdk_b200_exchange_init(void* local_exchange, void* stream);
/* init work table + scan (+ publish to peers) + wait / merge / finalize.  scratch: hdk_b200_plan_check bytes + 64.
 * peer_exchange: HOST array [n_peers] of DEVICE pointers, entry r = rank r's exchange buffer as mapped here. */
HDK_B200_API int hdk_b200_launch_exchange(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, const hdk_b200_kernel_options* ko,
                                          const hdk_b200_kernel_params* params, void* scratch, size_t scratch_bytes,
                                          void* const* peer_exchange, int n_peers, int my_rank, uint64_t epoch, void* stream,
                                          hdk_b200_launch_info* info);

/* ============================================================================
 * Result-set side helpers on the device ("next" row: ResultSet → Arrow).
 * Compact the non-empty entries (ResultSetStorage::isEmptyEntry,
 * omniscidb/ResultSet/ResultSetStorage.cpp:439-525) of a group-by buffer into
 * dense output columns, one 8-byte cell per target per row, AVG finalised with
 * pair_to_double (ResultSetBufferAccessors.h:168-195).  `row_count` is a device
 * uint64 the kernel increments.  Columns are int64 or double per target type.
 * ==========================================================================*/
HDK_B200_API int hdk_b200_compact_result(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd,
                            const int64_t* groups_buffer,
                            int64_t* const* out_cols /* DEVICE array [n_targets] of int64[entry_count] */,
                            uint64_t* row_count /* DEVICE */, void* stream);

/* Arrow buffers of one result column from its compacted 8-byte cells, on the device (ArrowResultSetConverter::
 * convertToArrowTable, omniscidb/ResultSet/ArrowResultSetConverter.cpp — there a host loop over ResultSet rows): the value
 * buffer of the Arrow type (int8/16/32/64, float, double: out_width bytes per row, NULL slots zeroed) and the validity
 * bitmap (bit i = row i is not NULL, LSB first, ceil(n_rows / 32) words; may be NULL when the column is not nullable).
 * A cell is NULL when it equals null_int (integers) / null_fp (cells holding doubles) — the sentinels
 * hdk_b200_compact_result normalises to.  *null_count (DEVICE, optional) receives the number of NULLs. */
HDK_B200_API int hdk_b200_arrow_column_on_device(const int64_t* cells, uint64_t n_rows, int cells_are_fp, int out_is_fp, int out_width,
                                                 int nullable, int64_t null_int, double null_fp, int8_t* values, uint32_t* validity,
                                                 uint64_t* null_count, void* stream);

/* ============================================================================
 * ORDER BY / LIMIT over the compacted result, on the device ("next" row: top-k / ORDER BY over aggregated results).
 * Replaces sortResultSet (QE/ResultSetSort.cpp:752-851): the permutation it leaves in the ResultSet
 * (ResultSet::setPermutationBuffer) is computed here, over the dense rows hdk_b200_compact_result wrote.
 * Order of one entry = ResultSetComparator::operator() (:333-470): NULLs before or after every value according to
 * `nulls_first` whatever the direction, then (lhs < rhs) != is_desc; rows equal on every entry come in unspecified
 * order (the reference uses std::sort / std::partial_sort).  Dictionary-encoded targets compare by string: pass the
 * rank of each dictionary id in string order.  LIMIT n = the first n entries of the permutation
 * (topPermutation, :504-520); hdk_b200_gather_rows then materialises those rows only.
 * ==========================================================================*/
typedef struct hdk_b200_order_entry {
  int32_t column;       /* index into cols[] = tle_no - 1 (hdk::ir::OrderEntry) */
  int32_t is_fp;        /* the cells are doubles (bits), else int64 */
  int32_t type_width;   /* logical width of the target's type: selects the NULL sentinel compact_result wrote */
  int32_t nullable;
  int32_t is_desc;
  int32_t nulls_first;
  const int32_t* dict_rank; /* DEVICE int32[dict_size] or NULL */
  int64_t dict_size;
} hdk_b200_order_entry;

HDK_B200_API size_t hdk_b200_sort_scratch_bytes(uint64_t n_rows);
/* cols: HOST array [HDK_B200_MAX_TARGETS or fewer, indexed by order[i].column] of DEVICE int64[n_rows];
 * permutation: DEVICE uint32[n_rows], receives the row ids in result order.
 * top_n: 0 = order all rows.  > 0 = the caller will use the first top_n entries only (LIMIT): when top_n << n_rows the rows
 * that cannot be among them are dropped before sorting (radix select on the first ORDER BY target, every tie kept), and
 * *n_sorted (HOST, optional) receives how many leading entries of `permutation` are valid (top_n <= *n_sorted <= n_rows).
 * That path reads one counter back and therefore synchronises `stream`; without it the call is asynchronous. */
HDK_B200_API int hdk_b200_sort_permutation(const int64_t* const* cols, const hdk_b200_order_entry* order, int n_order,
                                           uint64_t n_rows, uint64_t top_n, uint32_t* permutation, uint64_t* n_sorted,
                                           void* scratch, size_t scratch_bytes, void* stream);
/* cols_out[c][i] = cols_in[c][permutation[i]] for i < n_out; cols_in / cols_out: HOST arrays [n_cols] of DEVICE pointers */
HDK_B200_API int hdk_b200_gather_rows(const int64_t* const* cols_in, int64_t* const* cols_out, int n_cols,
                                      const uint32_t* permutation, uint64_t n_out, void* stream);

/* ============================================================================
 * Storage-side helper on the device ("next" row: ArrowStorage -> device residency).
 * Arrow fixed-width column data + validity bitmap, already copied to the device,
 * become the chunk format of the hot path in place: slots whose validity bit is 0
 * receive the in-band NULL sentinel (omniscidb/ArrowStorage/ArrowStorageUtils.cpp:100-170:
 * INTn_MIN for integers / timestamps / date32 days, FLT_MIN / DBL_MIN for floating point), and the
 * chunk statistics the planner needs (min, max, has_nulls — ArrowStorage.cpp:1000-1040
 * ChunkStats; they decide perfect vs baseline hash) are reduced in the same pass.
 * `stats` lives on the device and ACCUMULATES over calls (one call per Arrow chunk of a
 * fragment's column); initialise it with hdk_b200_init_chunk_stats_on_device.
 * `validity` may be NULL (no NULLs); bit i of the column is bit (bit_offset + i), LSB first.
 * ==========================================================================*/
typedef struct hdk_b200_chunk_stats {
  int64_t min_i, max_i;         /* integer columns: over the non-NULL values (INT64_MAX / INT64_MIN when none) */
  int64_t min_f_enc, max_f_enc; /* floating point: order-preserving int64 image of the double, b ^ ((b >> 63) & INT64_MAX) */
  uint64_t null_count;          /* values equal to the sentinel after materialisation */
  uint64_t row_count;
} hdk_b200_chunk_stats;
HDK_B200_API int hdk_b200_init_chunk_stats_on_device(hdk_b200_chunk_stats* stats, void* stream);
HDK_B200_API int hdk_b200_materialize_nulls_on_device(int8_t* values, int elem_width /* 1, 2, 4, 8 */, int is_fp,
                                                      const uint8_t* validity, int64_t bit_offset, int64_t num_elems,
                                                      hdk_b200_chunk_stats* stats, void* stream);

/* ============================================================================
 * Host-buffer convenience wrapper = what the reference-facing plugin call does
 * end to end (H2D of the chunks, init, launch, D2H of the group-by buffer and the
 * error code; QE/QueryExecutionContext.cpp:238-550).  All pointers are HOST
 * pointers; `col_buffers[f * n_cols + c]`; `out_buffer` receives
 * hdk_b200_buffer_size_bytes(qmd) bytes.  Synchronous.  Returns the aggregate
 * in-band error code (>0 / <0) or 0, or HDK_B200_E_* (≤ -1000 offset) on host errors.
 * ==========================================================================*/
HDK_B200_API int hdk_b200_query_host(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd,
                        const int8_t* const* col_buffers, const int64_t* num_rows,
                        uint64_t num_fragments,
                        const int64_t* join_hash_tables_host /* HOST addresses of HOST int32 tables, [n_joins] or NULL */,
                        const size_t* join_table_bytes,
                        const int8_t* const* inner_col_buffers /* HOST, [n_joins*MAX_COLS] or NULL */,
                        const size_t* inner_col_bytes,
                        int8_t* out_buffer, int device, hdk_b200_launch_info* info);

/* ============================================================================
 * Run-time specialisation.  The reference compiles EVERY work unit (Executor::compileWorkUnit, QE/NativeCodegen.cpp:
 * 1403-1560) and caches the native code by plan; here the fused kernel is a template over the plan's structure
 * (expression DAG, key / aggregate kinds, column widths — never literals, key ranges or entry counts): the named configs'
 * shapes are instantiated at build time, any other supported plan is instantiated by NVRTC (libnvrtc, loaded on demand)
 * from the same source the first time its shape is launched and cached by structural signature for the life of the
 * process.  While a shape compiles in the background its launches run the interpreting kernel — same results.
 * Without libnvrtc everything runs on the interpreting kernel.
 * ==========================================================================*/
typedef struct hdk_b200_jit_stats {
  uint64_t shapes_compiled, shapes_failed, shapes_pending;
  uint64_t launches;                 /* launches that ran run-time compiled kernels */
  double last_compile_ms, total_compile_ms;
  int available;                     /* libnvrtc found and usable */
} hdk_b200_jit_stats;
HDK_B200_API int hdk_b200_jit_get_stats(hdk_b200_jit_stats* out);
/* Stop run-time specialisation: drops the shapes still queued and returns once no compile is in flight (a few seconds at most).
 * Call it before the process exits (the compile runs on a worker thread that must not outlive the libraries it uses); later
 * launches of shapes without kernels keep running interpreted. */
HDK_B200_API int hdk_b200_jit_shutdown(void);
/* block until no shape is being compiled (tests, benchmarks) */
HDK_B200_API int hdk_b200_jit_wait(void);

/* ---- misc ------------------------------------------------------------------ */
/* Process-wide debug / tuning knobs (tests, tools/): never needed for correct results.
 *   "force_generic"            1 = never dispatch to a pre-compiled plan shape (run the interpreting kernel)
 *   "force_strategy"           -1 = library picks; HDK_B200_STRATEGY_* = accumulation strategy of perfect-hash launches
 *   "partitioned_aggregation"  -1 = library picks; 0 = never; 1 = whenever the plan is eligible and the scratch suffices
 *   "partitioned_table_slots"  0 = library picks; else the slots of the per-CTA shared table (tests force partition splits)
 *   "partitioned_partitions"   0 = library picks; else the number of partitions
 *   "partitioned_heavy_rows"   0 = library picks; else the partition size (rows) beyond which a partitioned launch falls
 *                              back to the global-table probe (hot keys)
 *   "jit"                      run-time specialisation (see hdk_b200_jit_*): 0 off, 1 compile in the background (default
 *                              when NVRTC is present), 2 compile before the first launch of a shape
 *   "geo_env_refresh"          1 = read the HDK_B200_GEO tuning override on every launch instead of once per process
 *                              (tools/sweep_geo.py) */
HDK_B200_API int hdk_b200_debug_set(const char* name, int value);
HDK_B200_API const char* hdk_b200_last_error(void);
HDK_B200_API int hdk_b200_abi_version(void);
HDK_B200_API int hdk_b200_device_count(void);
/* number of kernels this library has launched in this process (bench "gpu_launches") */
HDK_B200_API uint64_t hdk_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HDK_B200_H */
