"""ctypes mirror of include/hdk_b200.h (the C ABI).  Field order and sizes must match the
header exactly; tests/test_abi.py checks sizeof() of every struct against the library."""
import ctypes as C

MAX_EXPRS = 48
MAX_KEYS = 8
MAX_TARGETS = 24
MAX_SLOTS = 32
MAX_FILTERS = 8
MAX_JOINS = 4
MAX_COLS = 32
ABI_VERSION = 2

EMPTY_KEY_64 = 9223372036854775807
EMPTY_KEY_32 = 2147483647

# host error codes
OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4
# in-band error codes (QE/Execute.h:1019-1031)
ERR_DIV_BY_ZERO = 1
ERR_OUT_OF_SLOTS = 3
ERR_OVERFLOW_OR_UNDERFLOW = 7

INT, FP = 0, 1

(OP_COL, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_CAST, OP_EXTRACT_YEAR, OP_LT, OP_LE, OP_GT, OP_GE,
 OP_EQ, OP_NE, OP_AND, OP_OR, OP_NOT, OP_IS_NULL, OP_UMINUS, OP_CASE) = range(20)

AGG_NONE, AGG_COUNT, AGG_SUM, AGG_MIN, AGG_MAX, AGG_AVG = range(6)
PERFECT_HASH, BASELINE_HASH = 0, 1
STRATEGY_THREAD_PRIVATE, STRATEGY_CTA_SHARED, STRATEGY_GLOBAL, STRATEGY_BASELINE, STRATEGY_REGISTER, STRATEGY_PARTITIONED = range(6)
SMALL_DATE, SIGNED, UNSIGNED, DOUBLE = range(4)


class Type(C.Structure):
    _fields_ = [("kind", C.c_int8), ("width", C.c_int8), ("nullable", C.c_int8), ("pad", C.c_int8)]

    def __init__(self, kind=INT, width=8, nullable=0):
        super().__init__(kind, width, int(bool(nullable)), 0)

    def key(self):
        return (self.kind, self.width, self.nullable)

    def __repr__(self):
        return f"{'fp' if self.kind else 'int'}{self.width * 8}{'?' if self.nullable else ''}"


class Expr(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("aux", C.c_int32), ("type", Type),
                ("guard", C.c_int32), ("ival", C.c_int64), ("fval", C.c_double)]


class Target(C.Structure):
    _fields_ = [("agg", C.c_int32), ("arg", C.c_int32), ("type", Type), ("arg_type", Type),
                ("skip_null_val", C.c_int32), ("key_index", C.c_int32), ("slot", C.c_int32), ("pad", C.c_int32)]


class Qmd(C.Structure):
    _fields_ = [("hash_type", C.c_int32), ("keyless", C.c_int32), ("target_idx_for_key", C.c_int32),
                ("output_columnar", C.c_int32), ("entry_count", C.c_uint32), ("key_count", C.c_int32),
                ("key_width", C.c_int32), ("slot_count", C.c_int32), ("min_val", C.c_int64),
                ("max_val", C.c_int64), ("bucket", C.c_int64), ("has_nulls", C.c_int32), ("pad", C.c_int32),
                ("slot_padded", C.c_int8 * MAX_SLOTS), ("slot_logical", C.c_int8 * MAX_SLOTS),
                ("init_vals", C.c_int64 * MAX_SLOTS)]


class Key(C.Structure):
    _fields_ = [("expr", C.c_int32), ("has_nulls", C.c_int32), ("min_val", C.c_int64), ("max_val", C.c_int64),
                ("bucket", C.c_int64), ("cardinality", C.c_int64)]


class Join(C.Structure):
    _fields_ = [("key_expr", C.c_int32), ("one_to_many", C.c_int32), ("min_key", C.c_int64),
                ("max_key", C.c_int64), ("null_val", C.c_int64), ("key_nullable", C.c_int32), ("payload_by_slot", C.c_int32),
                ("entry_count", C.c_int64), ("n_key_exprs", C.c_int32), ("key_width", C.c_int32),
                ("key_exprs", C.c_int32 * MAX_KEYS)]


class Plan(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_exprs", C.c_int32), ("n_filters", C.c_int32),
                ("n_keys", C.c_int32), ("n_targets", C.c_int32), ("n_joins", C.c_int32), ("n_cols", C.c_int32),
                ("pad", C.c_int32), ("exprs", Expr * MAX_EXPRS), ("filters", C.c_int32 * MAX_FILTERS),
                ("keys", Key * MAX_KEYS), ("targets", Target * MAX_TARGETS), ("joins", Join * MAX_JOINS)]


class KernelParams(C.Structure):
    _fields_ = [("col_buffers", C.c_void_p), ("num_fragments", C.c_uint64), ("literals", C.c_void_p),
                ("num_rows", C.c_void_p), ("frag_row_offsets", C.c_void_p), ("max_matched", C.c_int32),
                ("total_matched", C.c_void_p), ("init_agg_vals", C.c_void_p), ("groupby_buf", C.c_void_p),
                ("error_codes", C.c_void_p), ("num_tables", C.c_uint32), ("join_hash_tables", C.c_void_p),
                ("inner_col_buffers", C.c_void_p), ("total_rows_hint", C.c_uint64)]


class JitStats(C.Structure):
    _fields_ = [("shapes_compiled", C.c_uint64), ("shapes_failed", C.c_uint64), ("shapes_pending", C.c_uint64), ("launches", C.c_uint64),
                ("last_compile_ms", C.c_double), ("total_compile_ms", C.c_double), ("available", C.c_int)]


VARIANT_JIT = 1000


class KernelOptions(C.Structure):
    _fields_ = [("gridDimX", C.c_uint), ("gridDimY", C.c_uint), ("gridDimZ", C.c_uint), ("blockDimX", C.c_uint),
                ("blockDimY", C.c_uint), ("blockDimZ", C.c_uint), ("sharedMemBytes", C.c_uint),
                ("literalsOffset", C.c_uint), ("hoistLiterals", C.c_int)]


class LaunchInfo(C.Structure):
    _fields_ = [("variant", C.c_int32), ("strategy", C.c_int32), ("n_launches", C.c_int32), ("grid", C.c_int32),
                ("block", C.c_int32), ("smem_bytes", C.c_int32), ("n_accumulators", C.c_int32), ("tile_rows", C.c_int32)]


class OrderEntry(C.Structure):
    _fields_ = [("column", C.c_int32), ("is_fp", C.c_int32), ("type_width", C.c_int32), ("nullable", C.c_int32),
                ("is_desc", C.c_int32), ("nulls_first", C.c_int32), ("dict_rank", C.c_void_p), ("dict_size", C.c_int64)]


class ChunkStatsPOD(C.Structure):
    _fields_ = [("min_i", C.c_int64), ("max_i", C.c_int64), ("min_f_enc", C.c_int64), ("max_f_enc", C.c_int64),
                ("null_count", C.c_uint64), ("row_count", C.c_uint64)]


class WorkTableLayout(C.Structure):
    _fields_ = [("n_cells", C.c_uint64), ("sum_cells", C.c_uint64), ("sum_i64_cells", C.c_uint64),
                ("min_cells", C.c_uint64), ("max_cells", C.c_uint64)]


class JoinChunk(C.Structure):
    _fields_ = [("col_buff", C.c_void_p), ("num_elems", C.c_size_t), ("row_id", C.c_size_t)]


class JoinColumn(C.Structure):
    _fields_ = [("col_chunks_buff", C.c_void_p), ("col_chunks_buff_sz", C.c_size_t), ("num_chunks", C.c_size_t),
                ("num_elems", C.c_size_t), ("elem_sz", C.c_size_t)]


class JoinColumnTypeInfo(C.Structure):
    _fields_ = [("elem_sz", C.c_size_t), ("min_val", C.c_int64), ("max_val", C.c_int64), ("null_val", C.c_int64),
                ("uses_bw_eq", C.c_int), ("translated_null_val", C.c_int64), ("column_type", C.c_int)]


def int_null(width):
    """Shared/InlineNullValues.h:33-39"""
    return {1: -(1 << 7), 2: -(1 << 15), 4: -(1 << 31), 8: -(1 << 63)}[width]


FLT_MIN = 1.1754943508222875e-38
DBL_MIN = 2.2250738585072014e-308


def fp_null(width):
    return FLT_MIN if width == 4 else DBL_MIN
