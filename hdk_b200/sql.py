"""A small SQL front end for the hot-path plan shapes.

HDK parses SQL with Apache Calcite over JNI (omniscidb/Calcite, out of scope: no JVM here) and
turns the relational algebra into a RelAlgExecutionUnit (QE/WorkUnitBuilder.cpp).  This module
accepts the SELECT … FROM … [JOIN … ON a = b] [WHERE …] GROUP BY … [ORDER BY …] [LIMIT n] subset the
named configs use and produces the same execution-unit shape (ir.ExecutionUnit), so that the
parity tests can be written as SQL strings like the reference's own (`c("SELECT …", dt)`,
omniscidb/Tests/ArrowBasedExecuteTest.cpp).  Anything else raises UnsupportedPlan.
"""
from __future__ import annotations

import datetime
import re
from typing import Dict, List, Optional, Tuple

from . import ir
from .planner import UnsupportedPlan

_TOKEN = re.compile(r"""
    \s*(?:
      (?P<num>\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+|\d+)
    | (?P<str>'(?:[^']|'')*')
    | (?P<id>[A-Za-z_][A-Za-z_0-9]*|"[^"]+")
    | (?P<op><>|!=|<=|>=|[-+*/=<>(),.])
    )""", re.X)

_KEYWORDS = {"select", "from", "where", "group", "by", "order", "limit", "join", "inner", "on", "and", "or",
             "not", "as", "is", "null", "asc", "desc", "extract", "year", "cast", "date", "timestamp",
             "between", "count", "sum", "min", "max", "avg", "case", "when", "then", "else", "end", "nulls", "first", "last",
             # recognised only to be rejected: they must never be taken for a table alias or a column name
             "in", "left", "right", "full", "outer", "cross", "natural", "having", "distinct", "over", "union", "using"}


def _tokenize(s: str):
    pos, out = 0, []
    s = s.strip().rstrip(";")
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m or m.end() == pos:
            raise UnsupportedPlan(f"cannot tokenize SQL at: {s[pos:pos + 20]!r}")
        pos = m.end()
        if m.group("num"):
            out.append(("num", m.group("num")))
        elif m.group("str"):
            out.append(("str", m.group("str")[1:-1].replace("''", "'")))
        elif m.group("id"):
            t = m.group("id")
            if t.startswith('"'):
                out.append(("id", t[1:-1]))
            elif t.lower() in _KEYWORDS:
                out.append(("kw", t.lower()))
            else:
                out.append(("id", t))
        else:
            out.append(("op", m.group("op")))
    out.append(("eof", ""))
    return out


_SQL_TYPES = {"int": ("int", 4), "integer": ("int", 4), "bigint": ("int", 8), "smallint": ("int", 2),
              "tinyint": ("int", 1), "double": ("fp", 8), "float": ("fp", 4), "real": ("fp", 4)}


class _Parser:
    def __init__(self, sql: str, tables: Dict[str, object]):
        self.toks = _tokenize(sql)
        self.i = 0
        self.tables = tables           # name -> storage.Table
        self.scopes: List[Tuple[str, str, int]] = []   # (alias, table name, table index)
        self.aliases: Dict[str, ir.Expr] = {}

    # -- token helpers
    def peek(self, k=0):
        return self.toks[self.i + k]

    def eat(self, kind=None, val=None):
        t = self.toks[self.i]
        if (kind and t[0] != kind) or (val is not None and t[1] != val):
            raise UnsupportedPlan(f"SQL: expected {val or kind}, got {t[1]!r}")
        self.i += 1
        return t

    def accept(self, kind, val=None):
        t = self.toks[self.i]
        if t[0] == kind and (val is None or t[1] == val):
            self.i += 1
            return True
        return False

    # -- name resolution
    def resolve(self, qual: Optional[str], col: str) -> ir.Expr:
        if qual is None and col.lower() in self.aliases:
            return self.aliases[col.lower()]      # select-list alias used in GROUP BY / ORDER BY
        hits = []
        for alias, tname, tidx in self.scopes:
            if qual is not None and qual.lower() not in (alias.lower(), tname.lower()):
                continue
            t = self.tables[tname]
            for cname, ci in t.columns.items():
                if cname.lower() == col.lower():
                    hits.append(ir.ColumnRef(tidx, cname, ci.type, ci.phys_width))
        if len(hits) != 1:
            raise UnsupportedPlan(f"SQL: cannot resolve column {qual + '.' if qual else ''}{col}")
        return hits[0]

    # -- expressions (precedence: or < and < not < cmp < add < mul < unary)
    def expr(self):
        e = self.and_expr()
        while self.accept("kw", "or"):
            e = ir.Logic("or", (e, self.and_expr()))
        return e

    def and_expr(self):
        e = self.not_expr()
        while self.accept("kw", "and"):
            e = ir.Logic("and", (e, self.not_expr()))
        return e

    def not_expr(self):
        if self.accept("kw", "not"):
            return ir.Logic("not", (self.not_expr(),))
        return self.cmp_expr()

    def _coerce_dict_literal(self, a, b, op):
        """'literal' vs a dictionary-encoded column: the literal becomes its dictionary id (ids are table-wide,
        ArrowStorage keeps one dictionary per column).  A string that is not in the dictionary gets an id no row has."""
        def conv(col, lit):
            if isinstance(lit, ir.Const) and isinstance(lit.value, str):
                if not (isinstance(col, ir.ColumnRef) and col.type.kind == "dict"):
                    raise UnsupportedPlan("string literal compared with a non-dictionary expression")
                if op not in ("=", "<>"):
                    raise UnsupportedPlan("only = and <> are defined between a dictionary column and a string literal")
                tname = [tn for _, tn, tidx in self.scopes if tidx == col.table][0]
                d = self.tables[tname].columns[col.column].dictionary or []
                return ir.Const(d.index(lit.value) if lit.value in d else len(d), ir.SqlType("int", 4, False))
            return lit
        return conv(b, a), conv(a, b)

    def _equals(self, op, lhs, rhs):
        lhs, rhs = self._coerce_temporal(lhs, rhs)
        literal = lambda e: isinstance(e, ir.Const) and isinstance(e.value, str)   # noqa: E731
        if not (literal(lhs) or literal(rhs)):
            if lhs.type.kind == "dict" and rhs.type.kind == "dict" and not (
                    isinstance(lhs, ir.ColumnRef) and isinstance(rhs, ir.ColumnRef) and (lhs.table, lhs.column) == (rhs.table, rhs.column)):
                # ids of two dictionaries are unrelated; the reference translates one dictionary into the other
                # (StringDictionaryTranslationMgr) — out of scope here, and never compared by id
                raise UnsupportedPlan("comparison of two dictionary-encoded columns needs a dictionary translation")
            if (lhs.type.kind == "dict") != (rhs.type.kind == "dict"):
                raise UnsupportedPlan("comparison of a dictionary-encoded column with a non-string expression")
        lhs, rhs = self._coerce_dict_literal(lhs, rhs, op)
        unit_of = lambda t: t.unit if t.kind == "timestamp" else (1 if t.kind == "date" else None)   # noqa: E731
        if unit_of(lhs.type) and unit_of(rhs.type) and unit_of(lhs.type) != unit_of(rhs.type):
            raise UnsupportedPlan("comparison of temporal values of different precision")
        return ir.make_cmp(op, lhs, rhs)

    def cmp_expr(self):
        lhs = self.add_expr()
        t = self.peek()
        if t[0] == "op" and t[1] in ("<", "<=", ">", ">=", "=", "<>", "!="):
            self.i += 1
            rhs = self.add_expr()
            return self._equals("<>" if t[1] == "!=" else t[1], lhs, rhs)
        neg_in = t == ("kw", "not") and self.peek(1) == ("kw", "in")
        if t == ("kw", "in") or neg_in:
            # x [NOT] IN (a, b, ...): a disjunction of equalities (three-valued: a NULL x stays NULL and is filtered)
            self.i += 2 if neg_in else 1
            self.eat("op", "(")
            e = None
            while True:
                c = self._equals("=", lhs, self.add_expr())
                e = c if e is None else ir.Logic("or", (e, c))
                if not self.accept("op", ","):
                    break
            self.eat("op", ")")
            return ir.Logic("not", (e,)) if neg_in else e
        if t == ("kw", "is"):
            self.i += 1
            neg = self.accept("kw", "not")
            self.eat("kw", "null")
            e = ir.IsNull(lhs)
            return ir.Logic("not", (e,), ir.SqlType("bool", 1, False)) if neg else e
        if t == ("kw", "between"):
            self.i += 1
            lo = self.add_expr()
            self.eat("kw", "and")
            hi = self.add_expr()
            l1, lo = self._coerce_temporal(lhs, lo)
            l2, hi = self._coerce_temporal(lhs, hi)
            return ir.Logic("and", (ir.make_cmp(">=", l1, lo), ir.make_cmp("<=", l2, hi)))
        return lhs

    def _coerce_temporal(self, a, b):
        """DATE/TIMESTAMP literal vs temporal column: express the literal in the column's unit."""
        def conv(col, lit):
            if isinstance(lit, ir.Const) and isinstance(lit.value, datetime.datetime) :
                secs = int((lit.value - datetime.datetime(1970, 1, 1)).total_seconds())
                ct = col.type
                if ct.kind == "timestamp":
                    v = secs * ct.unit
                elif ct.kind == "date":
                    v = secs           # date-in-days columns decode to seconds
                else:
                    raise UnsupportedPlan("date literal compared with a non-temporal expression")
                return ir.Const(v, ir.SqlType("int", 8, False))
            return lit
        return conv(b, a), conv(a, b)

    def add_expr(self):
        e = self.mul_expr()
        while self.peek()[0] == "op" and self.peek()[1] in "+-":
            op = self.eat()[1]
            e = ir.make_binop(op, e, self.mul_expr())
        return e

    def mul_expr(self):
        e = self.unary()
        while self.peek()[0] == "op" and self.peek()[1] in "*/":
            op = self.eat()[1]
            e = ir.make_binop(op, e, self.unary())
        return e

    def unary(self):
        if self.accept("op", "-"):
            a = self.unary()
            if isinstance(a, ir.Const):
                return ir.Const(-a.value, a.type)
            return ir.UMinus(a, a.type)
        if self.accept("op", "+"):
            return self.unary()
        return self.primary()

    def primary(self):
        t = self.peek()
        if t[0] == "num":
            self.i += 1
            s = t[1]
            if re.fullmatch(r"\d+", s):
                v = int(s)
                return ir.Const(v, ir.SqlType("int", 4 if -(1 << 31) < v < (1 << 31) else 8, False))
            return ir.Const(float(s), ir.SqlType("fp", 8, False))
        if t[0] == "op" and t[1] == "(":
            self.i += 1
            e = self.expr()
            self.eat("op", ")")
            return e
        if t[0] == "kw" and t[1] in ("date", "timestamp") and self.peek(1)[0] == "str":
            self.i += 1
            s = self.eat("str")[1]
            fmt = "%Y-%m-%d" if len(s) <= 10 else "%Y-%m-%d %H:%M:%S"
            return ir.Const(datetime.datetime.strptime(s, fmt), ir.SqlType("timestamp", 8, False))
        if t[0] == "str":
            self.i += 1
            return ir.Const(t[1], ir.SqlType("dict", 4, False))
        if t[0] == "kw" and t[1] in ("count", "sum", "min", "max", "avg"):
            self.i += 1
            self.eat("op", "(")
            if t[1] == "count" and self.accept("op", "*"):
                self.eat("op", ")")
                return ir.make_agg("count", None, self.bigint_count)
            arg = self.expr()
            self.eat("op", ")")
            try:
                return ir.make_agg(t[1], arg, self.bigint_count)
            except NotImplementedError as ex:
                raise UnsupportedPlan(str(ex))
        if t == ("kw", "null"):
            self.i += 1
            return ir.Const(None, ir.SqlType("int", 4, True))       # only meaningful as a CASE value, which retypes it
        if t == ("kw", "case"):
            # searched CASE, or simple CASE x WHEN v … (rewritten to x = v)
            self.i += 1
            subject = None if self.peek() == ("kw", "when") else self.expr()
            arms, else_ = [], None
            while self.accept("kw", "when"):
                c = self.expr()
                if subject is not None:
                    c = self._equals("=", subject, c)
                self.eat("kw", "then")
                arms.append((c, self.expr()))
            if not arms:
                raise UnsupportedPlan("CASE without WHEN")
            if self.accept("kw", "else"):
                else_ = self.expr()
            self.eat("kw", "end")
            try:
                return ir.make_case(arms, else_)
            except NotImplementedError as ex:
                raise UnsupportedPlan(str(ex))
        if t == ("kw", "extract"):
            self.i += 1
            self.eat("op", "(")
            self.eat("kw", "year")
            self.eat("kw", "from")
            arg = self.expr()
            self.eat("op", ")")
            if arg.type.kind not in ("timestamp", "date"):
                raise UnsupportedPlan("EXTRACT from a non-temporal expression")
            return ir.ExtractYear(arg, ir.SqlType("int", 8, arg.type.nullable))
        if t == ("kw", "cast"):
            self.i += 1
            self.eat("op", "(")
            arg = self.expr()
            self.eat("kw", "as")
            tn = self.eat()[1].lower()
            if tn == "double" and self.peek()[1].lower() == "precision":
                self.i += 1
            self.eat("op", ")")
            if tn not in _SQL_TYPES:
                raise UnsupportedPlan(f"CAST to {tn}")
            kind, w = _SQL_TYPES[tn]
            return ir.cast_to(arg, ir.SqlType(kind, w, arg.type.nullable))
        if t[0] == "id":
            self.i += 1
            if self.accept("op", "."):
                col = self.eat("id")[1]
                return self.resolve(t[1], col)
            return self.resolve(None, t[1])
        raise UnsupportedPlan(f"SQL: unexpected token {t[1]!r}")

    # -- statement
    def parse(self, bigint_count=False) -> ir.ExecutionUnit:
        self.bigint_count = bigint_count
        self.eat("kw", "select")
        # the select list needs the FROM scopes: scan ahead for FROM at depth 0
        depth, j = 0, self.i
        while not (self.toks[j] == ("kw", "from") and depth == 0):
            if self.toks[j] == ("op", "("):
                depth += 1
            elif self.toks[j] == ("op", ")"):
                depth -= 1
            elif self.toks[j][0] == "eof":
                raise UnsupportedPlan("SQL: missing FROM")
            j += 1
        select_start, self.i = self.i, j + 1
        table, alias = self._table_ref()
        self.scopes.append((alias, table, 0))
        joins_raw = []
        while self.peek() in (("kw", "join"), ("kw", "inner"), ("op", ",")):
            if self.accept("op", ","):
                # implicit join: FROM a, b WHERE a.x = b.x — the equalities against b's columns are taken out of the WHERE
                # clause below (what the reference's join-qual extraction does, QE/RelAlgExecutor / RelAlgOptimizer)
                t2, a2 = self._table_ref()
                self.scopes.append((a2, t2, len(joins_raw) + 1))
                joins_raw.append((t2, None))
                continue
            if self.accept("kw", "inner"):
                pass
            self.eat("kw", "join")
            t2, a2 = self._table_ref()
            self.scopes.append((a2, t2, len(joins_raw) + 1))
            self.eat("kw", "on")
            cond = self.expr()
            joins_raw.append((t2, cond))
        after_from = self.i
        # select list
        self.i = select_start
        targets, names = [], []
        while True:
            e = self.expr()
            name = None
            if self.accept("kw", "as"):
                name = self.eat()[1]
            elif self.peek()[0] == "id":
                name = self.eat("id")[1]
            targets.append(e)
            names.append(name)
            if not self.accept("op", ","):
                break
        if self.peek() != ("kw", "from"):
            raise UnsupportedPlan(f"SQL: unexpected {self.peek()[1]!r} in select list")
        self.i = after_from
        for e, n in zip(targets, names):
            if n is not None and not isinstance(e, ir.AggExpr):
                self.aliases[n.lower()] = e
        quals = []
        if self.accept("kw", "where"):
            visible, self.aliases = self.aliases, {}       # select-list aliases are not visible in WHERE
            quals = _split_conjuncts(self.expr())
            self.aliases = visible
        groupby = []
        if self.accept("kw", "group"):
            self.eat("kw", "by")
            while True:
                g = self.expr()
                if isinstance(g, ir.Const) and isinstance(g.value, int) and 1 <= g.value <= len(targets):
                    g = targets[g.value - 1]      # GROUP BY ordinal
                groupby.append(g)
                if not self.accept("op", ","):
                    break
        order = []
        n_hidden = 0
        if self.accept("kw", "order"):
            self.eat("kw", "by")
            while True:
                save = self.i
                key = None
                if self.peek()[0] == "id" and self.peek()[1] in [n for n in names if n]:
                    key = self.eat()[1]
                else:
                    self.i = save
                    e = self.expr()
                    if isinstance(e, ir.Const) and isinstance(e.value, int):
                        if not 1 <= e.value <= len(targets) - n_hidden:
                            raise UnsupportedPlan(f"ORDER BY ordinal {e.value} is not a select item")
                        key = e.value - 1
                    elif e in targets:
                        key = targets.index(e)
                    elif e in groupby:
                        # ordered by a group key the query does not project: carried as a trailing hidden target
                        targets.append(e)
                        names.append(f"__order_by_{len(targets)}")
                        n_hidden += 1
                        key = len(targets) - 1
                    else:
                        raise UnsupportedPlan("ORDER BY expression must be a select item or a group key")
                desc = False
                if self.accept("kw", "desc"):
                    desc = True
                else:
                    self.accept("kw", "asc")
                # Calcite's default (RelFieldCollation): NULLs sort high — last when ascending, first when descending
                nulls_first = desc
                if self.accept("kw", "nulls"):
                    if self.accept("kw", "last"):
                        nulls_first = False
                    else:
                        self.eat("kw", "first")
                        nulls_first = True
                order.append((key, desc, nulls_first))
                if not self.accept("op", ","):
                    break
        limit = None
        if self.accept("kw", "limit"):
            limit = int(self.eat("num")[1])
        self.eat("eof")
        joins = []
        def conjuncts(c):
            if isinstance(c, ir.Logic) and c.op == "and":
                out = []
                for x in c.args:
                    out += conjuncts(x)
                return out
            return [c]

        strip_cast = lambda x: x.arg if isinstance(x, ir.Cast) else x  # noqa: E731

        def max_table(e):
            return max([e.table] if isinstance(e, ir.ColumnRef) else [0] + [max_table(c) for c in e.children()])

        for t2, cond in joins_raw:
            jidx = len(joins) + 1
            if cond is None:
                # comma join: WHERE conjuncts `inner column of this table = expression over earlier tables`
                mine, rest = [], []
                for c in quals:
                    ok = False
                    if isinstance(c, ir.Cmp) and c.op == "=":
                        l, r = strip_cast(c.lhs), strip_cast(c.rhs)
                        ok = (isinstance(r, ir.ColumnRef) and r.table == jidx and max_table(l) < jidx) or \
                             (isinstance(l, ir.ColumnRef) and l.table == jidx and max_table(r) < jidx)
                    (mine if ok else rest).append(c)
                if not mine:
                    raise UnsupportedPlan("cross join: no equality between the comma-joined table and the tables before it")
                quals = rest
                cond = mine[0] if len(mine) == 1 else ir.Logic("and", tuple(mine))
            pairs = []
            for c in conjuncts(cond):
                if not (isinstance(c, ir.Cmp) and c.op == "="):
                    raise UnsupportedPlan("only (conjunctions of) equi-join conditions are on the hot path")
                # the implicit widening Cast the comparison puts on the narrower side is dropped (both sides are compared
                # as integers of the join table's key width); any other Cast — narrowing, int <-> fp — changes the value
                # and stays part of the outer expression (or refuses the plan on the inner side)
                def strip(x):
                    if isinstance(x, ir.Cast) and not x.type.is_fp and not x.arg.type.is_fp and x.type.width >= x.arg.type.width \
                            and x.type.kind == x.arg.type.kind:
                        return x.arg
                    return x
                lhs, rhs = strip(c.lhs), strip(c.rhs)
                if isinstance(rhs, ir.ColumnRef) and rhs.table == jidx:
                    outer, inner = lhs, rhs
                elif isinstance(lhs, ir.ColumnRef) and lhs.table == jidx:
                    outer, inner = rhs, lhs
                else:
                    raise UnsupportedPlan("join condition must compare an outer expression with an inner column")
                try:
                    pairs.append((ir.join_key_for(outer, inner.type), inner.column, inner.type))
                except NotImplementedError as ex:
                    raise UnsupportedPlan(str(ex))
            joins.append(ir.JoinSpec(t2, pairs[0][0], pairs[0][1], [(o, c) for o, c, _ in pairs[1:]], [t for _, _, t in pairs]))
        final_names = []
        for i, (e, n) in enumerate(zip(targets, names)):
            if n is None:
                n = e.column if isinstance(e, ir.ColumnRef) else f"EXPR${i}"
            final_names.append(n)
        order = [(final_names.index(k) if isinstance(k, str) else k, d, nf) for k, d, nf in order]
        return ir.ExecutionUnit(table, groupby, targets, final_names, quals, joins, order, limit, n_hidden)

    def _table_ref(self):
        name = self.eat("id")[1]
        real = None
        for t in self.tables:
            if t.lower() == name.lower():
                real = t
        if real is None:
            raise UnsupportedPlan(f"SQL: unknown table {name}")
        alias = real
        if self.accept("kw", "as"):
            alias = self.eat("id")[1]
        elif self.peek()[0] == "id":
            alias = self.eat("id")[1]
        return real, alias


def _split_conjuncts(e: ir.Expr) -> List[ir.Expr]:
    if isinstance(e, ir.Logic) and e.op == "and":
        out = []
        for a in e.args:
            out += _split_conjuncts(a)
        return out
    return [e]


def parse(sql: str, tables: Dict[str, object], bigint_count=False) -> ir.ExecutionUnit:
    return _Parser(sql, tables).parse(bigint_count)
