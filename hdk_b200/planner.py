"""Host-side planning for the hot path: query memory descriptor + plan POD.

Restates, for the supported plan shapes only, the decisions HDK takes in
  * get_col_range_info                QE/MemoryLayoutBuilder.cpp:91-238
  * get_keyless_info                  QE/MemoryLayoutBuilder.cpp:249-416
  * pick_target_compact_width         QE/MemoryLayoutBuilder.cpp:559-652
  * pick_baseline_key_width           QE/MemoryLayoutBuilder.cpp:654-690
  * build_query_memory_descriptor     QE/MemoryLayoutBuilder.cpp:795-994
  * ColSlotContext                    omniscidb/ResultSet/ColSlotContext.cpp:34-87
  * init_agg_val_vec / get_agg_initial_val  QE/OutputBufferInitialization.cpp:30-68, 112-258
and lowers an ir.ExecutionUnit to the C-ABI `hdk_b200_plan` (include/hdk_b200.h).
In a real drop-in HDK builds the QueryMemoryDescriptor itself and hands it over; this module
exists so the pyhdk-shaped façade can run the same shapes stand-alone.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Callable, List, Optional

from . import abi, ir


@dataclass
class Config:
    """Shared/Config.h:42-62 (group_by section) — defaults are the reference's."""
    bigint_count: bool = False
    default_max_groups_buffer_entry_guess: int = 16384
    big_group_threshold: int = 16384
    baseline_threshold: int = 1000000
    enable_columnar_output: bool = False
    # hdk_b200 extension: probe OneToOne joins through a presence bitmap + slot-ordered inner columns
    join_payload_by_slot: bool = True
    # perfect join tables up to this many entries (PerfectJoinHashTable.cpp:139-151); beyond: baseline join table
    max_perfect_join_entries: int = (1 << 31) // 4
    # Config::exec.codegen.null_div_by_zero: a zero divisor yields NULL instead of ERR_DIV_BY_ZERO
    null_div_by_zero: bool = False


class UnsupportedPlan(Exception):
    """Plan shape outside the hot path.  Never a CPU fallback (north_star)."""


INT_MAX = {1: 127, 2: 32767, 4: 2147483647, 8: 9223372036854775807}
FLT_MAX = 3.4028234663852886e+38
DBL_MAX = 1.7976931348623157e+308


def _dbits(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", x))[0]


def _fbits(x: float) -> int:
    return struct.unpack("<i", struct.pack("<f", x))[0]


def _logical_width(t: ir.SqlType) -> int:
    return t.width


# ---------------------------------------------------------------------------------------------
@dataclass
class TargetInfo:
    """Shared/TargetInfo.h:34-45"""
    is_agg: bool
    agg: int
    type: ir.SqlType
    arg_type: Optional[ir.SqlType]
    skip_null_val: bool
    arg: Optional[ir.Expr]
    key_index: int = -1

    @property
    def compact_type(self) -> ir.SqlType:
        """get_compact_type, Shared/SqlTypesLayout.h:36-55"""
        if not self.is_agg or self.arg_type is None:
            return self.type
        if self.agg in (abi.AGG_MIN, abi.AGG_MAX):
            return self.arg_type
        return self.type.with_nullable(self.arg_type.nullable)

    @property
    def float_argument_input(self) -> bool:
        return self.is_agg and self.agg in (abi.AGG_AVG, abi.AGG_SUM, abi.AGG_MIN, abi.AGG_MAX) and \
            self.arg_type is not None and self.arg_type.is_fp and self.arg_type.width == 4


_AGG_CODE = {"count": abi.AGG_COUNT, "sum": abi.AGG_SUM, "min": abi.AGG_MIN, "max": abi.AGG_MAX, "avg": abi.AGG_AVG}


def get_target_info(e: ir.Expr, groupby: List[ir.Expr], bigint_count: bool) -> TargetInfo:
    if isinstance(e, ir.AggExpr):
        code = _AGG_CODE[e.agg]
        if e.arg is None:
            return TargetInfo(True, abi.AGG_COUNT, ir.int_t(8 if bigint_count else 4, False), None, False, None)
        at = e.arg.type
        if code == abi.AGG_AVG:
            t = ir.int_t(8, at.nullable) if not at.is_fp else at
            return TargetInfo(True, code, t, at, at.nullable, e.arg)
        if code == abi.AGG_COUNT:
            return TargetInfo(True, code, ir.int_t(8 if bigint_count else 4, False), at, at.nullable, e.arg)
        return TargetInfo(True, code, e.type, at, at.nullable, e.arg)
    try:
        kidx = groupby.index(e)
    except ValueError:
        raise UnsupportedPlan("non-aggregate target that is not a group key (projection is outside the hot path)")
    return TargetInfo(False, abi.AGG_NONE, e.type, None, False, e, kidx)


def get_agg_initial_val(agg: int, t: ir.SqlType, min_byte_width_to_compact: int) -> int:
    """QE/OutputBufferInitialization.cpp:112-258 with enable_compaction=true (group by)."""
    byte_width = max(t.width, min_byte_width_to_compact)
    if agg in (abi.AGG_COUNT, abi.AGG_AVG):
        return 0
    if agg == abi.AGG_SUM:
        if t.nullable:
            if t.is_fp:
                return _fbits(abi.FLT_MIN) if byte_width == 4 else _dbits(abi.fp_null(t.width))
            return abi.int_null(t.width)
        return 0
    if agg == abi.AGG_MIN:
        if t.is_fp:
            if byte_width == 4:
                return _fbits(FLT_MAX) if not t.nullable else _fbits(abi.FLT_MIN)
            return _dbits(DBL_MAX) if not t.nullable else _dbits(abi.fp_null(t.width))
        return INT_MAX[byte_width] if not t.nullable else abi.int_null(t.width)
    if agg == abi.AGG_MAX:
        if t.is_fp:
            if byte_width == 4:
                return _fbits(-FLT_MAX) if not t.nullable else _fbits(abi.FLT_MIN)
            return _dbits(-DBL_MAX) if not t.nullable else _dbits(abi.fp_null(t.width))
        return (-INT_MAX[byte_width] - 1) if not t.nullable else abi.int_null(t.width)
    raise UnsupportedPlan(f"agg {agg}")


# ---------------------------------------------------------------------------------------------
@dataclass
class ColRangeInfo:
    hash_type: int
    min: int
    max: int
    bucket: int
    has_nulls: bool

    def bucketed_cardinality(self) -> int:
        c = self.max - self.min
        if self.bucket:
            c //= self.bucket
        return c + 1 + (1 if self.has_nulls else 0)


def get_expr_range_info(e: ir.Expr, col_stats) -> ColRangeInfo:
    r = ir.expr_range(e, col_stats)
    if r.kind == "int":
        if r.lo > r.hi:
            return ColRangeInfo(abi.BASELINE_HASH, 0, -1, 0, r.has_nulls)
        return ColRangeInfo(abi.PERFECT_HASH, int(r.lo), int(r.hi), r.bucket, r.has_nulls)
    return ColRangeInfo(abi.BASELINE_HASH, 0, 0, 0, False)


def get_col_range_info(unit: ir.ExecutionUnit, col_stats, cfg: Config) -> ColRangeInfo:
    gb = unit.groupby_exprs
    if len(gb) != 1:
        cardinality, has_nulls = 1, False
        for g in gb:
            ri = get_expr_range_info(g, col_stats)
            if ri.hash_type != abi.PERFECT_HASH:
                return ColRangeInfo(abi.BASELINE_HASH, 0, 0, 0, False)
            cardinality *= ri.bucketed_cardinality()
            has_nulls = has_nulls or ri.has_nulls
            if cardinality >= (1 << 63):
                return ColRangeInfo(abi.BASELINE_HASH, 0, (1 << 63) - 1, 0, False)
        if not cardinality or cardinality > cfg.baseline_threshold:
            return ColRangeInfo(abi.BASELINE_HASH, 0, cardinality, 0, has_nulls)
        return ColRangeInfo(abi.PERFECT_HASH, 0, cardinality, 0, has_nulls)
    ri = get_expr_range_info(gb[0], col_stats)
    if gb[0].type.kind == "timestamp" and gb[0].type.unit > 1 and unit.quals:
        return ColRangeInfo(abi.BASELINE_HASH, 0, 0, 0, False)
    col_count = len(gb) + len(unit.target_exprs)
    max_entry_count = (1 << 30) // (col_count * 8)
    if (ri.max - ri.min) >= max_entry_count and not ri.bucket:
        return ColRangeInfo(abi.BASELINE_HASH, ri.min, ri.max, 0, ri.has_nulls)
    return ri


def get_keyless_info(infos: List[TargetInfo], col_stats) -> tuple:
    keyless, found, index = True, False, 0
    for ti in infos:
        if not found and ti.is_agg:
            r = ir.expr_range(ti.arg, col_stats) if ti.arg is not None else None
            if ti.agg == abi.AGG_AVG:
                index += 1
                if not (ti.arg_type.nullable and (r.kind == "invalid" or r.has_nulls)):
                    found = True
            elif ti.agg == abi.AGG_COUNT:
                if not (ti.arg is not None and ti.arg_type.nullable and (r.kind == "invalid" or r.has_nulls)):
                    found = True
            elif ti.agg == abi.AGG_SUM:
                if ti.arg_type.nullable:
                    if r.kind != "invalid" and not r.has_nulls:
                        found = True
                elif r.kind in ("int", "fp") and (r.hi < 0 or r.lo > 0):
                    found = True
            elif ti.agg == abi.AGG_MIN:
                init = get_agg_initial_val(ti.agg, ti.compact_type, 4 if ti.float_argument_input else 8)
                if r.kind == "fp":
                    # Deviation, on purpose: the reference (MemoryLayoutBuilder.cpp:339-348) takes a nullable MIN whose values
                    # all lie below the NULL sentinel (DBL_MIN / FLT_MIN, tiny positive numbers — i.e. all-negative data) as
                    # the emptiness marker without looking at has_nulls, and then loses every group whose values are all
                    # NULL: the slot still holds its initial value.  It guards MAX against exactly this (:368-373).
                    init_f = struct.unpack("<d", struct.pack("<q", init))[0]
                    found = r.hi < init_f and not r.has_nulls
                elif r.kind == "int":
                    found = r.hi < init
            elif ti.agg == abi.AGG_MAX:
                if r.kind != "invalid" and not r.has_nulls:
                    init = get_agg_initial_val(ti.agg, ti.compact_type, 4 if ti.float_argument_input else 8)
                    if r.kind == "fp":
                        init_f = struct.unpack("<d", struct.pack("<q", init))[0]
                        found = r.lo > init_f
                    elif r.kind == "int":
                        found = r.lo > init
            else:
                keyless = False
        if not keyless:
            break
        if not found:
            index += 1
    return keyless and found, index


def pick_target_compact_width(unit: ir.ExecutionUnit, infos: List[TargetInfo], total_tuples: int, cfg: Config,
                              crt_min_byte_width: int = 8) -> int:
    if cfg.bigint_count:
        return 8
    compact = 0
    if len(unit.groupby_exprs) != 1:
        compact = crt_min_byte_width
    if not compact:
        for ti in infos:
            if ti.is_agg and ti.arg is not None:
                compact = crt_min_byte_width
                break
            if ti.is_agg:
                continue
            if (ti.type.kind == "int" and ti.type.width <= 4) or ti.type.kind == "dict":
                continue
            compact = crt_min_byte_width
            break
    if not compact:
        return 4 if total_tuples <= 0xFFFFFFFF else crt_min_byte_width
    widths = []
    for ti in infos:
        widths.append(_logical_width(ti.compact_type))
        if ti.agg == abi.AGG_AVG:
            widths.append(8)
    return max([compact] + widths)


def pick_baseline_key_width(unit: ir.ExecutionUnit, col_stats) -> int:
    w = 4
    for g in unit.groupby_exprs:
        r = ir.expr_range(g, col_stats)
        if r.kind == "invalid":
            cw = 8
        elif r.kind == "int":
            if g.type.width == 8 and r.has_nulls:
                cw = 8
            else:
                cw = 4 if (r.lo > -(1 << 31) and r.hi < abi.EMPTY_KEY_32 - 1) else 8
        else:
            cw = 8
        w = max(w, cw)
    return w


@dataclass
class PlannedQuery:
    plan: abi.Plan
    qmd: abi.Qmd
    infos: List[TargetInfo]
    columns: List[str]                  # outer table physical columns, index = OP_COL.b
    inner_columns: List[List[str]]      # per join: inner table physical columns
    unit: ir.ExecutionUnit


def build_query(unit: ir.ExecutionUnit, col_stats: Callable, total_tuples: int, cfg: Config = Config(),
                max_groups_buffer_entry_count: Optional[int] = None, force_hash_type: Optional[int] = None,
                output_columnar: Optional[bool] = None) -> PlannedQuery:
    # Non-grouped aggregates (SELECT COUNT(*), SUM(x) ... without GROUP BY; the reference's
    # QueryDescriptionType::NonGroupedAggregate) run as the degenerate group-by: zero keys, one keyless entry.
    if cfg.null_div_by_zero:
        f = ir.with_null_div_by_zero
        unit = ir.ExecutionUnit(unit.table, [f(g) for g in unit.groupby_exprs], [f(t) for t in unit.target_exprs], unit.target_names,
                                [f(q) for q in unit.quals],
                                [ir.JoinSpec(j.inner_table, f(j.outer_key), j.inner_key_column, [(f(o), c) for o, c in j.more_keys],
                                             j.inner_key_types)
                                 for j in unit.joins], unit.order_by, unit.limit, unit.n_hidden)
    non_grouped = not unit.groupby_exprs
    if non_grouped and any(not isinstance(t, ir.AggExpr) for t in unit.target_exprs):
        raise UnsupportedPlan("a query without GROUP BY may only select aggregates")
    if len(unit.groupby_exprs) > abi.MAX_KEYS:
        raise UnsupportedPlan("too many group keys")
    for g in unit.groupby_exprs:
        if g.type.is_fp:
            raise UnsupportedPlan("floating-point group keys")
    columnar = cfg.enable_columnar_output if output_columnar is None else output_columnar
    infos = [get_target_info(t, unit.groupby_exprs, cfg.bigint_count) for t in unit.target_exprs]
    if non_grouped:
        # NonGroupedAggregate: every aggregate over an argument skips NULLs and starts from the NULL sentinel, whatever the
        # argument's declared nullability, so that no input row ⇒ NULL (TargetExprCodegenBuilder::operator(),
        # QE/TargetExprBuilder.cpp:546-552; init_agg_val_vec, QE/OutputBufferInitialization.cpp:56-59)
        for ti in infos:
            if ti.is_agg and ti.arg is not None:
                ti.skip_null_val = True
                ti.arg_type = ti.arg_type.with_nullable(True)
                if ti.agg != abi.AGG_COUNT:
                    ti.type = ti.type.with_nullable(True)

    cri = get_col_range_info(unit, col_stats, cfg)
    if force_hash_type is not None and force_hash_type != cri.hash_type:
        cri = ColRangeInfo(force_hash_type, cri.min, cri.max, cri.bucket, cri.has_nulls)
        if force_hash_type == abi.PERFECT_HASH:
            raise UnsupportedPlan("cannot force perfect hash on this key range")

    q = abi.Qmd()
    q.hash_type = cri.hash_type
    q.output_columnar = int(columnar)
    q.key_count = len(unit.groupby_exprs)
    min_slot_size = pick_target_compact_width(unit, infos, total_tuples, cfg)

    keys = []
    for g in unit.groupby_exprs:
        ri = get_expr_range_info(g, col_stats)
        keys.append(ri)

    if cri.hash_type == abi.PERFECT_HASH:
        keyless, tidx = get_keyless_info(infos, col_stats)
        q.keyless = int(keyless and not cri.bucket)
        if non_grouped:
            q.keyless, columnar = 1, False      # one entry, nothing to key; the row is always part of the result
            q.output_columnar = 0
            tidx = tidx if keyless else 0
        # (keyless + columnar is a layout the reference produces — MemoryLayoutBuilder.cpp:864-880 does not look at the
        #  columnar hint — with its quirk that get_columnar_group_bin_offset then writes into the first SLOT column,
        #  QE/RowFuncBuilder.cpp:604-607; finalize.cu reproduces it)
        q.target_idx_for_key = tidx
        q.key_width = 8
        if non_grouped:
            q.entry_count = 1
            q.min_val = q.max_val = q.bucket = q.has_nulls = 0
        elif len(unit.groupby_exprs) > 1:
            q.entry_count = cri.max
            q.min_val, q.max_val, q.bucket, q.has_nulls = 0, cri.max, 0, int(cri.has_nulls)
        else:
            q.entry_count = max(cri.bucketed_cardinality(), 1)
            q.min_val, q.max_val, q.bucket, q.has_nulls = cri.min, cri.max, cri.bucket, int(cri.has_nulls)
        key_targets_have_slots = True
        # slots narrower than 8 bytes make the reference retry without compaction when there is no GROUP BY
        # (CompilationRetryNoCompaction, QE/TargetExprBuilder.cpp:520-526)
        padded = 8 if non_grouped else min_slot_size
    else:
        q.keyless = 0
        q.target_idx_for_key = -1
        q.key_width = 8 if columnar else pick_baseline_key_width(unit, col_stats)
        q.entry_count = max_groups_buffer_entry_count or cfg.default_max_groups_buffer_entry_guess
        q.min_val = q.max_val = q.bucket = 0
        q.has_nulls = 0
        key_targets_have_slots = False
        padded = 8

    # ColSlotContext
    slot = 0
    slots_of = []
    for ti in infos:
        if not ti.is_agg and not key_targets_have_slots:
            slots_of.append(-1)
            continue
        slots_of.append(slot)
        lw = _logical_width(ti.compact_type)
        q.slot_padded[slot] = padded
        q.slot_logical[slot] = lw
        if ti.is_agg:
            mbw = 4 if ti.float_argument_input else padded
            q.init_vals[slot] = get_agg_initial_val(ti.agg, ti.compact_type, mbw)
        else:
            q.init_vals[slot] = 0
        slot += 1
        if ti.agg == abi.AGG_AVG:
            q.slot_padded[slot] = padded
            q.slot_logical[slot] = 8
            q.init_vals[slot] = 0
            slot += 1
        if slot > abi.MAX_SLOTS:
            raise UnsupportedPlan("too many slots")
    q.slot_count = slot

    # ---- plan POD -------------------------------------------------------------------------
    p = abi.Plan()
    p.abi_version = abi.ABI_VERSION
    columns: List[str] = []
    inner_columns: List[List[str]] = [[] for _ in unit.joins]
    node_of = {}

    def emit(op, a=0, b=0, aux=0, t=None, ival=0, fval=0.0, guard=0):
        n = p.n_exprs
        if n >= abi.MAX_EXPRS:
            raise UnsupportedPlan("expression too large")
        e = p.exprs[n]
        e.op, e.a, e.b, e.aux, e.type, e.ival, e.fval, e.guard = op, a, b, aux, t.abi(), int(ival), float(fval), guard
        p.n_exprs = n + 1
        return n

    raises_memo = {}

    def can_raise(e: ir.Expr) -> bool:
        """Does evaluating e possibly raise a row error (division by zero, checked integer overflow)?"""
        if e not in raises_memo:
            own = (isinstance(e, ir.BinOp) and ((e.op == "/" and not e.null_on_zero) or (e.overflow_check and not e.type.is_fp))) or \
                  (isinstance(e, ir.Cast) and not e.type.is_fp and not e.arg.type.is_fp and e.arg.type.width > e.type.width) or \
                  (isinstance(e, ir.UMinus) and not e.type.is_fp and not e.arg.type.nullable)
            raises_memo[e] = own or any(can_raise(c) for c in e.children())
        return raises_memo[e]

    BOOL_NN, BOOL_N = ir.SqlType("bool", 1, False), ir.SqlType("bool", 1, True)
    unsafe_memo = {}

    def unsafe_division(e: ir.Expr) -> bool:
        """contains_unsafe_division (QE/LogicalIR.cpp:26-53): a division whose divisor is not a non-zero literal"""
        if e not in unsafe_memo:
            own = isinstance(e, ir.BinOp) and e.op == "/" and not e.null_on_zero and \
                not (isinstance(e.rhs, ir.Const) and e.rhs.value not in (None, 0, 0.0))
            unsafe_memo[e] = own or any(unsafe_division(c) for c in e.children())
        return unsafe_memo[e]

    def both(g: Optional[int], c: int) -> int:
        return c if g is None else emit(abi.OP_AND, g, c, t=BOOL_N)

    def lower(e: ir.Expr, guard: Optional[int] = None) -> int:
        """guard: the node that is true for the rows where the reference would execute e at all (the CASE arm it sits
        in, QE/CaseIR.cpp:66-93).  Only nodes that can raise carry it — everything else is pure and evaluated for every
        row — so a sub-expression shared between an arm and the rest of the query is one node unless it can raise."""
        if not can_raise(e):
            guard = None
        memo_key = (e, guard)
        if memo_key in node_of:
            return node_of[memo_key]
        g = 0 if guard is None else guard + 1
        if isinstance(e, ir.ColumnRef):
            if e.table == 0:
                if e.column not in columns:
                    columns.append(e.column)
                idx = columns.index(e.column)
            else:
                cl = inner_columns[e.table - 1]
                if e.column not in cl:
                    cl.append(e.column)
                idx = cl.index(e.column)
            n = emit(abi.OP_COL, e.table, idx, 1 if e.type.date_in_days else 0, e.type, ival=e.phys_width)
        elif isinstance(e, ir.Const):
            if e.value is None:       # the NULL arm of a CASE: the type's sentinel
                t = e.type.with_nullable(True)
                n = emit(abi.OP_CONST, t=t, ival=0 if t.is_fp else abi.int_null(t.width),
                         fval=abi.fp_null(t.width) if t.is_fp else 0.0)
            elif isinstance(e.value, str):
                raise UnsupportedPlan("string literal outside a comparison with a dictionary-encoded column")
            else:
                n = emit(abi.OP_CONST, t=e.type.with_nullable(False),
                         ival=0 if e.type.is_fp else int(e.value), fval=float(e.value))
        elif isinstance(e, ir.BinOp):
            a, b = lower(e.lhs, guard), lower(e.rhs, guard)
            op = {"+": abi.OP_ADD, "-": abi.OP_SUB, "*": abi.OP_MUL, "/": abi.OP_DIV}[e.op]
            n = emit(op, a, b, (1 if (e.overflow_check and not e.type.is_fp) else 0) | (2 if e.null_on_zero else 0), e.type, guard=g)
        elif isinstance(e, ir.UMinus):
            n = emit(abi.OP_UMINUS, lower(e.arg, guard), t=e.type, guard=g)
        elif isinstance(e, ir.Cast):
            n = emit(abi.OP_CAST, lower(e.arg, guard), t=e.type, guard=g)
        elif isinstance(e, ir.ExtractYear):
            at = e.arg.type
            units = at.unit if at.kind == "timestamp" else 1
            n = emit(abi.OP_EXTRACT_YEAR, lower(e.arg, guard), t=e.type.with_nullable(at.nullable), ival=units)
        elif isinstance(e, ir.Case):
            # WHEN i is reached when no earlier WHEN was true; THEN i additionally needs WHEN i; the arms nest from the
            # last one outwards: CASE(c1, v1, CASE(c2, v2, else))
            reach, conds, thens = guard, [], []
            pieces = [x for arm in e.arms for x in arm] + [e.else_]
            for i, (c, v) in enumerate(e.arms):
                cn = lower(c, reach)
                conds.append(cn)
                thens.append(lower(v, both(reach, cn) if can_raise(v) else None))
                if any(can_raise(x) for x in pieces[2 * i + 2:]):      # what follows needs "no WHEN so far was true"
                    nt = emit(abi.OP_NOT, cn, t=c.type)
                    if c.type.nullable:
                        nt = emit(abi.OP_OR, emit(abi.OP_IS_NULL, cn, t=BOOL_NN), nt, t=BOOL_NN)
                    reach = both(reach, nt)
            n = lower(e.else_, reach)
            for cn, tn in reversed(list(zip(conds, thens))):
                n = emit(abi.OP_CASE, cn, tn, t=e.type, ival=n)
        elif isinstance(e, ir.Cmp):
            a, b = lower(e.lhs, guard), lower(e.rhs, guard)
            op = {"<": abi.OP_LT, "<=": abi.OP_LE, ">": abi.OP_GT, ">=": abi.OP_GE, "=": abi.OP_EQ,
                  "<>": abi.OP_NE}[e.op]
            n = emit(op, a, b, t=e.type)
        elif isinstance(e, ir.Logic):
            if e.op == "not":
                n = emit(abi.OP_NOT, lower(e.args[0], guard), t=e.type)
            else:
                # codegenLogicalShortCircuit (QE/LogicalIR.cpp:193-298): an operand with an unsafe division is generated
                # behind the other one and only runs when that one does not already decide the result
                # (false AND …, true OR …, NULL).  Here: safe operands first, each unsafe one guarded by what precedes it.
                args = [x for x in e.args if not unsafe_division(x)] + [x for x in e.args if unsafe_division(x)]
                n = None
                for x in args:
                    if n is None:
                        n = lower(x, guard)
                        continue
                    if unsafe_division(x):
                        decided = n if e.op == "and" else emit(abi.OP_NOT, n, t=BOOL_N)   # x runs when n is TRUE / FALSE
                        nx = lower(x, both(guard, decided))
                        # where x does not run the reference's phi yields n itself: FALSE / TRUE, or NULL when n is NULL
                        # (nullcheck_fail_bb) — not NULL AND x.  NULL in x's place gives exactly that through AND / OR.
                        null_bool = emit(abi.OP_CONST, t=BOOL_N, ival=abi.int_null(1))
                        nx = emit(abi.OP_CASE, decided, nx, t=BOOL_N, ival=null_bool)
                    else:
                        nx = lower(x, guard)
                    n = emit(abi.OP_AND if e.op == "and" else abi.OP_OR, n, nx, t=BOOL_N if unsafe_division(x) else e.type)
        elif isinstance(e, ir.IsNull):
            n = emit(abi.OP_IS_NULL, lower(e.arg, guard), t=e.type)
        else:
            raise UnsupportedPlan(f"expression {type(e).__name__}")
        node_of[memo_key] = n
        return n

    # joins first so that a join's key node precedes every inner-table column
    if len(unit.joins) > abi.MAX_JOINS:
        raise UnsupportedPlan("too many joins")
    p.n_joins = len(unit.joins)
    for j, js in enumerate(unit.joins):
        nodes = [lower(k) for k in js.outer_keys]
        p.joins[j].key_expr = max(nodes)        # the probe point: every component has been evaluated by then
        p.joins[j].key_nullable = int(js.outer_key.type.nullable)
        p.joins[j].null_val = abi.int_null(js.outer_key.type.width)
        wide = False
        if len(nodes) == 1:
            lo, hi, _ = col_stats(j + 1, js.inner_key_column)
            wide = lo is not None and (hi - lo + 1) > cfg.max_perfect_join_entries
        if len(nodes) > 1 or wide:
            # composite key, or a single key whose range is too wide for a perfect table (TooManyHashEntries,
            # JHT/PerfectJoinHashTable.cpp:139-151) → baseline join table (JHT/BaselineJoinHashTable.cpp)
            if len(nodes) > abi.MAX_KEYS:
                raise UnsupportedPlan("too many join key components")
            p.joins[j].n_key_exprs = len(nodes)
            for i, nd in enumerate(nodes):
                p.joins[j].key_exprs[i] = nd
            # BaselineJoinHashTable::getKeyComponentWidth (JHT/BaselineJoinHashTable.cpp:502-509): 8 as soon as one INNER
            # key column is wider than 4 bytes (the table is built over the inner values: a narrower width would truncate
            # them and alias e.g. 2^32 + 5 with 5); a wide outer expression needs 8 too, for the same reason on the probe
            inner_wide = any(t.width > 4 for t in js.inner_key_types)
            p.joins[j].key_width = 8 if (inner_wide or any(k.type.width > 4 for k in js.outer_keys)) else 4
    if len(unit.quals) > abi.MAX_FILTERS:
        raise UnsupportedPlan("too many filters")
    p.n_filters = len(unit.quals)
    # quals with an unsafe division are deferred behind the others and only run for rows those accept
    # (should_defer_eval, QE/LogicalIR.cpp:57-74; Executor::compileBody's primary / deferred quals)
    primary = [f for f in unit.quals if not unsafe_division(f)]
    deferred = [f for f in unit.quals if unsafe_division(f)]
    passed = None
    for i, f in enumerate(primary):
        p.filters[i] = lower(f)
        passed = both(passed, p.filters[i]) if deferred else None
    for i, f in enumerate(deferred):
        p.filters[len(primary) + i] = lower(f, passed)
    p.n_keys = len(unit.groupby_exprs)
    for i, g in enumerate(unit.groupby_exprs):
        k = p.keys[i]
        k.expr = lower(g)
        ri = keys[i]
        if cri.hash_type == abi.PERFECT_HASH:
            k.has_nulls = int(ri.has_nulls)
            k.min_val, k.max_val, k.bucket = ri.min, ri.max, ri.bucket
            k.cardinality = ri.bucketed_cardinality()
        else:
            k.has_nulls = 0
            k.min_val = k.max_val = k.bucket = 0
            k.cardinality = 0
    if len(infos) > abi.MAX_TARGETS:
        raise UnsupportedPlan("too many targets")
    p.n_targets = len(infos)
    for i, ti in enumerate(infos):
        t = p.targets[i]
        t.agg = ti.agg
        t.type = ti.type.abi()
        t.arg_type = (ti.arg_type or ti.type).abi()
        t.skip_null_val = int(ti.skip_null_val)
        t.key_index = ti.key_index
        t.slot = slots_of[i]
        if ti.is_agg:
            t.arg = lower(ti.arg) if ti.arg is not None else -1
        else:
            t.arg = p.keys[ti.key_index].expr
    p.n_cols = len(columns)
    if p.n_cols > abi.MAX_COLS:
        raise UnsupportedPlan("too many columns")
    return PlannedQuery(p, q, infos, columns, inner_columns, unit)
