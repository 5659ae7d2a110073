"""Minimal relational IR for the hot path: typed scalar expressions, aggregate targets and the
execution unit the Executor hands to the kernels.

Mirrors the subset of hdk::ir (omniscidb/IR/Expr.h, Type.h) and RelAlgExecutionUnit
(omniscidb/QueryEngine/RelAlgExecutionUnit.h:132-241) that the named plan shapes use.
Host-side planning only — nothing here touches data.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

from . import abi


# ---------------------------------------------------------------------------------------------
# types
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class SqlType:
    """kind: 'int' | 'fp' | 'bool' | 'timestamp' | 'date' | 'dict'.
    width = logical byte width; unit = units per second for timestamps (1, 1e3, 1e6, 1e9);
    date: unit 'days' means the chunk stores days in `phys_width` bytes and decodes to seconds
    (fixed_width_small_date_decode, QE/DecodersImpl.h:153-161)."""
    kind: str
    width: int
    nullable: bool = True
    unit: int = 1
    date_in_days: bool = False
    dict_id: int = 0

    @property
    def is_fp(self):
        return self.kind == "fp"

    @property
    def is_integer_like(self):
        return self.kind in ("int", "bool", "timestamp", "date", "dict")

    def with_nullable(self, n):
        return SqlType(self.kind, self.width, bool(n), self.unit, self.date_in_days, self.dict_id)

    def abi(self) -> abi.Type:
        return abi.Type(abi.FP if self.is_fp else abi.INT, self.width, self.nullable)

    def null_value(self):
        return abi.fp_null(self.width) if self.is_fp else abi.int_null(self.width)


def int_t(width=8, nullable=True):
    return SqlType("int", width, nullable)


def fp_t(width=8, nullable=True):
    return SqlType("fp", width, nullable)


BOOL = SqlType("bool", 1, True)


# ---------------------------------------------------------------------------------------------
# expressions
# ---------------------------------------------------------------------------------------------
class Expr:
    type: SqlType

    def children(self) -> Sequence["Expr"]:
        return ()


@dataclass(frozen=True)
class ColumnRef(Expr):
    table: int          # 0 = outer (fact) table, j > 0 = inner table of join j-1
    column: str
    type: SqlType
    phys_width: int


@dataclass(frozen=True)
class Const(Expr):
    value: object
    type: SqlType


@dataclass(frozen=True)
class BinOp(Expr):
    op: str             # + - * /
    lhs: Expr
    rhs: Expr
    type: SqlType
    overflow_check: bool = True
    null_on_zero: bool = False      # division under Config.null_div_by_zero: NULL instead of ERR_DIV_BY_ZERO (safe_div_*)

    def children(self):
        return (self.lhs, self.rhs)


@dataclass(frozen=True)
class UMinus(Expr):
    arg: Expr
    type: SqlType

    def children(self):
        return (self.arg,)


@dataclass(frozen=True)
class Cast(Expr):
    arg: Expr
    type: SqlType

    def children(self):
        return (self.arg,)


@dataclass(frozen=True)
class ExtractYear(Expr):
    arg: Expr
    type: SqlType = SqlType("int", 8, True)

    def children(self):
        return (self.arg,)


@dataclass(frozen=True)
class Cmp(Expr):
    op: str             # < <= > >= = <>
    lhs: Expr
    rhs: Expr
    type: SqlType = BOOL

    def children(self):
        return (self.lhs, self.rhs)


@dataclass(frozen=True)
class Logic(Expr):
    op: str             # and or not
    args: tuple
    type: SqlType = BOOL

    def children(self):
        return self.args


@dataclass(frozen=True)
class IsNull(Expr):
    arg: Expr
    type: SqlType = SqlType("bool", 1, False)

    def children(self):
        return (self.arg,)


@dataclass(frozen=True)
class Case(Expr):
    """CASE WHEN c1 THEN v1 [WHEN c2 THEN v2 …] ELSE ve END (hdk::ir::CaseExpr).  Every value already has `type`."""
    arms: tuple         # ((when, then), …)
    else_: Expr
    type: SqlType

    def children(self):
        return tuple(x for arm in self.arms for x in arm) + (self.else_,)


@dataclass(frozen=True)
class AggExpr(Expr):
    agg: str            # count sum min max avg
    arg: Optional[Expr]
    type: SqlType

    def children(self):
        return (self.arg,) if self.arg is not None else ()


def common_numeric_type(a: SqlType, b: SqlType) -> SqlType:
    """hdk::ir::BinOper::commonNumericType (omniscidb/IR/Expr.cpp): fp beats int, wider beats narrower."""
    nullable = a.nullable or b.nullable
    if a.is_fp or b.is_fp:
        w = max(a.width if a.is_fp else 0, b.width if b.is_fp else 0)
        # an 8-byte integer mixed with float widens to double
        if w == 4 and ((not a.is_fp and a.width == 8) or (not b.is_fp and b.width == 8)):
            w = 8
        return fp_t(w, nullable)
    return int_t(max(a.width, b.width), nullable)


def cast_to(e: Expr, t: SqlType) -> Expr:
    same = (e.type.is_fp == t.is_fp) and e.type.width == t.width
    if same:
        return e
    if isinstance(e, Const):
        v = e.value
        if v is None:
            return Const(None, t)
        return Const(float(v) if t.is_fp else int(v), t.with_nullable(False))
    return Cast(e, t.with_nullable(e.type.nullable))


def make_case(arms, else_: Optional[Expr]) -> Expr:
    """The arms unify to their common numeric type; no ELSE means ELSE NULL, and then (or with any nullable arm) the
    result is nullable — what the reference's CASE normalisation produces (omniscidb/IR/Expr.cpp, CaseExpr typing)."""
    values = [v for _, v in arms] + ([else_] if else_ is not None else [])
    typed = [v for v in values if not (isinstance(v, Const) and v.value is None)]
    if not typed:
        raise NotImplementedError("CASE with only NULL values")
    for c, _ in arms:
        if c.type.kind != "bool":
            raise NotImplementedError("CASE condition must be boolean")
    t = typed[0].type
    if all(v.type.kind == "int" or v.type.is_fp for v in typed):
        for v in typed[1:]:
            t = common_numeric_type(t, v.type)
    elif any((v.type.kind, v.type.width, getattr(v.type, "unit", 0)) != (t.kind, t.width, getattr(t, "unit", 0)) for v in typed) \
            or t.kind in ("dict", "bool"):
        raise NotImplementedError("CASE over non-numeric values of different types")
    t = t.with_nullable(len(typed) < len(values) or else_ is None or any(v.type.nullable for v in typed))

    def conv(v):
        if isinstance(v, Const) and v.value is None:
            return Const(None, t)
        v = cast_to(v, t)
        return v
    return Case(tuple((c, conv(v)) for c, v in arms), conv(else_) if else_ is not None else Const(None, t), t)


def join_key_for(outer: Expr, inner_type: SqlType) -> Expr:
    """The probe-side expression of one equi-join component.  The join table is built over the inner column's stored
    values; a days-encoded date column stores days while its expressions decode to seconds, so a date = date component
    probes with the outer column's stored days too.  Any other pairing with a days-encoded column is refused, and so are
    dictionary-encoded keys (two dictionaries: the ids are unrelated)."""
    if inner_type.kind == "dict" or outer.type.kind == "dict":
        raise NotImplementedError("join on dictionary-encoded columns needs a dictionary translation")
    if inner_type.is_fp or outer.type.is_fp:
        # the reference refuses to hash-join on a floating-point key (HashJoinFail → loop join, outside the hot path);
        # probing with the raw bits of a double would silently match nothing
        raise NotImplementedError("hash join on a floating-point key")
    if inner_type.date_in_days or outer.type.date_in_days:
        if isinstance(outer, ColumnRef) and outer.type.date_in_days and inner_type.date_in_days and outer.phys_width == 4:
            return ColumnRef(outer.table, outer.column, SqlType("int", 4, outer.type.nullable), outer.phys_width)
        raise NotImplementedError("join of a days-encoded date column with anything but another one")
    if (inner_type.kind == "timestamp" or outer.type.kind == "timestamp") and \
            (inner_type.kind, inner_type.unit) != (outer.type.kind, outer.type.unit):
        raise NotImplementedError("join of timestamps of different precision")
    return outer


def with_null_div_by_zero(e: Expr) -> Expr:
    """The expression under Config::exec.codegen.null_div_by_zero (QE/ArithmeticIR.cpp:587-597): every division yields NULL
    for a zero divisor (safe_div_*), so it and everything computed from it become nullable."""
    import dataclasses

    def rebuild(x):
        if isinstance(x, tuple):
            return tuple(rebuild(y) for y in x)
        if not isinstance(x, Expr) or not dataclasses.is_dataclass(x):
            return x
        kw = {f.name: rebuild(getattr(x, f.name)) for f in dataclasses.fields(x) if f.name != "type"}
        kids = [v for v in kw.values() if isinstance(v, Expr)] + \
               [z for v in kw.values() if isinstance(v, tuple) for y in v for z in (y if isinstance(y, tuple) else (y,)) if isinstance(z, Expr)]
        t = x.type
        if isinstance(x, BinOp) and x.op == "/":
            kw["null_on_zero"] = True
            t = t.with_nullable(True)
        elif isinstance(x, AggExpr):
            if x.agg != "count" and kids:
                t = t.with_nullable(t.nullable or kids[0].type.nullable)
        elif not isinstance(x, (ColumnRef, Const, IsNull)):
            t = t.with_nullable(t.nullable or any(k.type.nullable for k in kids))
        return type(x)(**kw, type=t)
    return rebuild(e)


def make_binop(op: str, lhs: Expr, rhs: Expr) -> Expr:
    t = common_numeric_type(lhs.type, rhs.type)
    lhs, rhs = cast_to(lhs, t), cast_to(rhs, t)
    t = t.with_nullable(lhs.type.nullable or rhs.type.nullable)
    return BinOp(op, lhs, rhs, t)


def make_cmp(op: str, lhs: Expr, rhs: Expr) -> Expr:
    lt, rt = lhs.type, rhs.type
    if lt.is_fp or rt.is_fp or (lt.kind == "int" and rt.kind == "int"):
        t = common_numeric_type(lt, rt)
        lhs, rhs = cast_to(lhs, t), cast_to(rhs, t)
    elif isinstance(rhs, Const) and rhs.value is not None:
        rhs = Const(int(rhs.value), SqlType("int", lt.width, False))
    elif isinstance(lhs, Const) and lhs.value is not None:
        lhs = Const(int(lhs.value), SqlType("int", rt.width, False))
    return Cmp(op, lhs, rhs, SqlType("bool", 1, lhs.type.nullable or rhs.type.nullable))


def make_agg(agg: str, arg: Optional[Expr], bigint_count=False) -> AggExpr:
    """Result types as the reference derives them (Shared/TargetInfo.h get_target_info and
    hdk::ir::AggExpr typing): COUNT → int32 (int64 with bigint_count), SUM(int) → int64,
    SUM/MIN/MAX(fp) → same fp, AVG → double."""
    agg = agg.lower()
    if agg == "count":
        return AggExpr("count", arg, int_t(8 if bigint_count else 4, False))
    assert arg is not None
    at = arg.type
    if at.kind in ("dict", "bool"):
        # the reference throws for MIN / MAX / SUM / AVG over strings ("Aggregate on … is not supported",
        # ArrowBasedExecuteTest.cpp:2840 expects it); ids are not values
        raise NotImplementedError(f"{agg.upper()} over a {'dictionary-encoded string' if at.kind == 'dict' else 'boolean'} is not supported")
    if agg == "sum":
        t = fp_t(at.width, at.nullable) if at.is_fp else int_t(8, at.nullable)
    elif agg in ("min", "max"):
        t = at
    elif agg == "avg":
        t = fp_t(8, at.nullable)
    else:
        raise NotImplementedError(f"aggregate {agg} is outside the hot path")
    return AggExpr(agg, arg, t)


# ---------------------------------------------------------------------------------------------
# expression ranges (QE/ExpressionRange.cpp): integer [min, max], has_nulls; fp [min, max]
# ---------------------------------------------------------------------------------------------
@dataclass
class Range:
    kind: str           # 'int' | 'fp' | 'invalid'
    lo: float = 0
    hi: float = 0
    has_nulls: bool = False
    bucket: int = 0


def year_of(seconds: int) -> int:
    """omniscidb/Utils/ExtractFromTime.cpp:260-271 (general path; equals the fast path where valid)."""
    day = seconds // 86400
    era = (day - 11017) // 146097
    doe = day - 11017 - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - (1 if doe == 146096 else 0)) // 365
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    marjan = 31 + 30 + 31 + 30 + 31 + 31 + 30 + 31 + 30 + 31
    return 2000 + era * 400 + yoe + (1 if marjan <= doy else 0)


def expr_range(e: Expr, col_stats) -> Range:
    """col_stats(table, column) -> (min, max, has_nulls) from chunk metadata."""
    if isinstance(e, ColumnRef):
        lo, hi, hn = col_stats(e.table, e.column)
        if lo is None:
            return Range("invalid")
        if e.type.is_fp:
            return Range("fp", float(lo), float(hi), hn)
        if e.type.date_in_days:
            # a DATE column's range counts in days: bucket = 86400 s (getExpressionRange(ColumnVar), QE/ExpressionRange.cpp:
            # 553-558 → get_conservative_datetrunc_bucket(kDay)); GROUP BY a date is then a perfect hash over the days
            return Range("int", int(lo) * 86400, int(hi) * 86400, hn, 86400)
        return Range("int", int(lo), int(hi), hn)
    if isinstance(e, Const):
        if e.value is None or isinstance(e.value, str):
            return Range("invalid")
        return Range("fp" if e.type.is_fp else "int", e.value, e.value, False)
    if isinstance(e, Cast):
        r = expr_range(e.arg, col_stats)
        if r.kind == "fp" and not e.type.is_fp:
            return Range("int", math.floor(r.lo), math.ceil(r.hi), r.has_nulls)
        if r.kind == "int" and e.type.is_fp:
            return Range("fp", float(r.lo), float(r.hi), r.has_nulls)
        return r
    if isinstance(e, ExtractYear):
        r = expr_range(e.arg, col_stats)
        if r.kind != "int":
            return Range("invalid")
        u = e.arg.type.unit if e.arg.type.kind == "timestamp" else 1
        # C++ integer division truncates toward zero (ExpressionRange.cpp:812-820)
        tdiv = lambda a, b: int(a / b) if b != 1 else a  # noqa: E731
        return Range("int", year_of(tdiv(r.lo, u)), year_of(tdiv(r.hi, u)), r.has_nulls)
    if isinstance(e, BinOp):
        a, b = expr_range(e.lhs, col_stats), expr_range(e.rhs, col_stats)
        if "invalid" in (a.kind, b.kind):
            return Range("invalid")
        hn = a.has_nulls or b.has_nulls
        kind = "fp" if e.type.is_fp else "int"
        if e.op == "+":
            return Range(kind, a.lo + b.lo, a.hi + b.hi, hn)
        if e.op == "-":
            return Range(kind, a.lo - b.hi, a.hi - b.lo, hn)
        if e.op == "*":
            c = [a.lo * b.lo, a.lo * b.hi, a.hi * b.lo, a.hi * b.hi]
            return Range(kind, min(c), max(c), hn)
        if e.op == "/" and kind == "int" and a.kind == "int" and b.kind == "int" and b.lo * b.hi > 0:
            # ExpressionRange::div (QE/ExpressionRange.cpp:199-218): only when the divisor's interval excludes 0
            tdiv = lambda x, y: abs(x) // abs(y) * (1 if (x >= 0) == (y >= 0) else -1)   # noqa: E731  (C truncation)
            c = [tdiv(a.lo, b.lo), tdiv(a.lo, b.hi), tdiv(a.hi, b.lo), tdiv(a.hi, b.hi)]
            return Range(kind, min(c), max(c), hn or e.null_on_zero)
        return Range("invalid")
    if isinstance(e, Case):
        # the union of the arms' ranges; a NULL arm only adds has_nulls (getExpressionRange(CaseExpr), ExpressionRange.cpp)
        out = None
        for v in [v for _, v in e.arms] + [e.else_]:
            if isinstance(v, Const) and v.value is None:
                r = None
            else:
                r = expr_range(v, col_stats)
                if r.kind == "invalid":
                    return r
            if r is None:
                if out is not None:
                    out.has_nulls = True
                else:
                    out = Range("null", has_nulls=True)
            elif out is None or out.kind == "null":
                out = Range(r.kind, r.lo, r.hi, r.has_nulls or out is not None)
            else:
                if out.kind != r.kind:
                    return Range("invalid")
                out = Range(r.kind, min(out.lo, r.lo), max(out.hi, r.hi), out.has_nulls or r.has_nulls)
        return out if out is not None and out.kind != "null" else Range("invalid")
    return Range("invalid")


# ---------------------------------------------------------------------------------------------
# the execution unit
# ---------------------------------------------------------------------------------------------
@dataclass
class JoinSpec:
    """Inner equi-join of the outer table with one inner table (JoinCondition / buildHashTableForQualifier,
    QE/Execute.cpp:3692).  One integer key → perfect join table when its range allows; several key columns (or a
    range too wide for a perfect table) → baseline join table.  `outer_key` / `inner_key_column` are the first pair;
    `more_keys` holds the remaining (outer expression, inner column) pairs of a composite key."""
    inner_table: str
    outer_key: Expr
    inner_key_column: str
    more_keys: List[tuple] = field(default_factory=list)
    inner_key_types: List[SqlType] = field(default_factory=list)   # types of the inner key columns (all components)

    @property
    def outer_keys(self):
        return [self.outer_key] + [o for o, _ in self.more_keys]

    @property
    def inner_key_columns(self):
        return [self.inner_key_column] + [c for _, c in self.more_keys]


@dataclass
class ExecutionUnit:
    table: str
    groupby_exprs: List[Expr]
    target_exprs: List[Expr]            # ColumnRef/expr equal to a group key, or AggExpr
    target_names: List[str]
    quals: List[Expr] = field(default_factory=list)
    joins: List[JoinSpec] = field(default_factory=list)
    order_by: List[tuple] = field(default_factory=list)   # (target index, is_desc, nulls_first) — hdk::ir::OrderEntry
    limit: Optional[int] = None
    n_hidden: int = 0                   # trailing targets that only serve ORDER BY and are dropped from the answer
