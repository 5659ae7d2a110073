// hdk_b200/csrc/lower.cu — host-side lowering of the C-ABI plan + query memory descriptor into
// the compact device plan (common.cuh).  Pure host code; no kernels here.
//
// What is restated from the reference here is *layout and call selection*, not arithmetic:
//   slot offsets      ColSlotContext::getAlignedPaddedSizeForRange (omniscidb/ResultSet/ColSlotContext.cpp:143-158)
//   row / buffer size QueryMemoryDescriptor::getRowSize / getBufferSizeBytes (QueryMemoryDescriptor.cpp:240-256, 457-481)
//   per-slot agg fn   TargetExprCodegen::codegenAggregate (QE/TargetExprBuilder.cpp:278-465)
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "common.cuh"
#include "partagg.cuh"

namespace hb {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;
DebugKnobs g_debug = {0, -1, -1, 0, 0, 0, 1, 0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static inline size_t align8(size_t x) { return (x + 7) & ~size_t(7); }

static int find_or_add_acc(DPlan& p, uint8_t kind, int arg, uint8_t arg_nullable) {
  for (int i = 0; i < p.n_acc; ++i)
    if (p.accs[i].kind == kind && p.accs[i].arg == arg && p.accs[i].arg_nullable == arg_nullable) return i;
  if (p.n_acc >= kMaxAcc) return -1;
  DAcc& a = p.accs[p.n_acc];
  a.kind = kind;
  a.arg = int8_t(arg);
  a.arg_nullable = arg_nullable;
  a.bytes = (kind == ACC_CNT_ALL || kind == ACC_CNT_NN) ? 4 : 8;
  return p.n_acc++;
}

static int acc_class(uint8_t kind) {  // 0 sum_i, 1 sum_f, 2 min, 3 max
  switch (kind) {
    case ACC_CNT_ALL: case ACC_CNT_NN: case ACC_SUM_I: return 0;
    case ACC_SUM_F: return 1;
    case ACC_MIN_I: case ACC_MIN_F: return 2;
    default: return 3;
  }
}

int lower_plan(const hdk_b200_plan* plan, const hdk_b200_qmd* q, Lowered* out) {
  if (!plan || !q || !out) { set_error("null plan/qmd"); return HDK_B200_E_INVALID; }
  if (plan->abi_version != HDK_B200_ABI_VERSION) { set_error("plan ABI version %d != %d", plan->abi_version, HDK_B200_ABI_VERSION); return HDK_B200_E_INVALID; }
  if (plan->n_exprs < 0 || plan->n_exprs > HDK_B200_MAX_EXPRS || plan->n_keys < 0 || plan->n_keys > HDK_B200_MAX_KEYS ||
      plan->n_targets < 1 || plan->n_targets > HDK_B200_MAX_TARGETS || plan->n_filters < 0 ||
      plan->n_filters > HDK_B200_MAX_FILTERS || plan->n_joins < 0 || plan->n_joins > HDK_B200_MAX_JOINS ||
      plan->n_cols < 0 || plan->n_cols > HDK_B200_MAX_COLS) {
    set_error("plan counts out of range");
    return HDK_B200_E_INVALID;
  }
  if (q->slot_count < 0 || q->slot_count > HDK_B200_MAX_SLOTS || q->entry_count == 0 || q->key_count != plan->n_keys) {
    set_error("query memory descriptor inconsistent with plan");
    return HDK_B200_E_INVALID;
  }
  memset(out, 0, sizeof(*out));
  DPlan& p = out->plan;
  DLayout& L = out->layout;
  p.n_exprs = plan->n_exprs;
  p.n_filters = plan->n_filters;
  p.n_keys = plan->n_keys;
  p.n_joins = plan->n_joins;
  p.n_cols = plan->n_cols;
  p.entry_count = q->entry_count;
  p.hash_type = q->hash_type;

  // ---- expressions
  for (int i = 0; i < plan->n_exprs; ++i) {
    const hdk_b200_expr& e = plan->exprs[i];
    DExpr& d = p.exprs[i];
    d.op = uint8_t(e.op);
    d.aux = uint8_t(e.aux);
    d.kind = uint8_t(e.type.kind);
    d.width = uint8_t(e.type.width);
    d.nullable = uint8_t(e.type.nullable);
    d.a = int8_t(e.a);
    d.b = int8_t(e.b);
    const bool unary = e.op == HDK_B200_OP_CAST || e.op == HDK_B200_OP_EXTRACT_YEAR || e.op == HDK_B200_OP_NOT ||
                       e.op == HDK_B200_OP_IS_NULL || e.op == HDK_B200_OP_UMINUS;
    const bool binary = (e.op >= HDK_B200_OP_ADD && e.op <= HDK_B200_OP_DIV) || (e.op >= HDK_B200_OP_LT && e.op <= HDK_B200_OP_OR);
    if ((unary || binary) && (e.a < 0 || e.a >= i)) { set_error("expr %d: operand a=%d not topologically earlier", i, e.a); return HDK_B200_E_INVALID; }
    if (binary && (e.b < 0 || e.b >= i)) { set_error("expr %d: operand b=%d not topologically earlier", i, e.b); return HDK_B200_E_INVALID; }
    if (e.type.width != 1 && e.type.width != 2 && e.type.width != 4 && e.type.width != 8) { set_error("expr %d: bad width", i); return HDK_B200_E_INVALID; }
    if (e.guard < 0 || e.guard > i) { set_error("expr %d: guard %d not topologically earlier", i, e.guard); return HDK_B200_E_INVALID; }
    d.guard = uint8_t(e.guard);
    switch (e.op) {
      case HDK_B200_OP_COL: {
        if (e.a < 0 || e.a > plan->n_joins || e.b < 0 || e.b >= HDK_B200_MAX_COLS) { set_error("expr %d: bad column ref", i); return HDK_B200_E_INVALID; }
        const int pw = int(e.ival);
        if (pw != 1 && pw != 2 && pw != 4 && pw != 8) { set_error("expr %d: bad physical width", i); return HDK_B200_E_INVALID; }
        if (e.type.kind == HDK_B200_FP && pw != 4 && pw != 8) { set_error("expr %d: bad fp width", i); return HDK_B200_E_INVALID; }
        d.imm.i = pw;
        if (e.a == 0) {
          if (e.b >= plan->n_cols) { set_error("expr %d: column %d >= n_cols", i, e.b); return HDK_B200_E_INVALID; }
          if (p.col_width[e.b] && p.col_width[e.b] != pw) { set_error("column %d used with two widths", e.b); return HDK_B200_E_INVALID; }
          p.col_width[e.b] = uint8_t(pw);
        }
        break;
      }
      case HDK_B200_OP_CONST:
        if (e.type.kind == HDK_B200_FP) d.imm.f = e.fval; else d.imm.i = e.ival;
        break;
      case HDK_B200_OP_EXTRACT_YEAR:
        d.imm.i = e.ival > 0 ? e.ival : 1;
        break;
      case HDK_B200_OP_CASE:
        if (e.a < 0 || e.a >= i || e.b < 0 || e.b >= i || e.ival < 0 || e.ival >= i) { set_error("expr %d: CASE operand not topologically earlier", i); return HDK_B200_E_INVALID; }
        for (int arm : {e.b, int(e.ival)})
          if (plan->exprs[arm].type.kind != e.type.kind || plan->exprs[arm].type.width != e.type.width) {
            set_error("expr %d: CASE arm %d does not have the node's type", i, arm); return HDK_B200_E_INVALID;
          }
        d.imm.i = e.ival;
        break;
      default:
        if (!(unary || binary)) { set_error("expr %d: unknown op %d", i, e.op); return HDK_B200_E_UNSUPPORTED; }
        d.imm.i = 0;
    }
  }
  for (int c = 0; c < plan->n_cols; ++c) {
    if (!p.col_width[c]) p.col_width[c] = 8;  // unused column: never staged
    out->stage_row_bytes += p.col_width[c];
  }
  for (int i = 0; i < plan->n_filters; ++i) {
    if (plan->filters[i] < 0 || plan->filters[i] >= plan->n_exprs) { set_error("bad filter node"); return HDK_B200_E_INVALID; }
    p.filters[i] = int8_t(plan->filters[i]);
    p.exprs[p.filters[i]].aux |= kAuxInQual;
  }
  // Everything a qual is made of: the reference generates the quals ahead of the filter branch and a failed check returns
  // from the row function at once (ArithmeticIR.cpp `CreateRet(ERR_…)`), so those errors do not depend on the row passing;
  // keys and targets live inside the filter-true block (QE/RowFuncBuilder.cpp:400-513).
  for (int i = plan->n_exprs - 1; i >= 0; --i) {
    const DExpr& d = p.exprs[i];
    if (!(d.aux & kAuxInQual)) continue;
    if (d.guard) p.exprs[d.guard - 1].aux |= kAuxInQual;
    if (d.op == HDK_B200_OP_COL || d.op == HDK_B200_OP_CONST) continue;
    p.exprs[d.a].aux |= kAuxInQual;
    const bool unary = d.op == HDK_B200_OP_CAST || d.op == HDK_B200_OP_EXTRACT_YEAR || d.op == HDK_B200_OP_NOT ||
                       d.op == HDK_B200_OP_IS_NULL || d.op == HDK_B200_OP_UMINUS;
    if (!unary) p.exprs[d.b].aux |= kAuxInQual;
    if (d.op == HDK_B200_OP_CASE) p.exprs[int(d.imm.i)].aux |= kAuxInQual;
  }
  // ---- joins
  for (int j = 0; j < plan->n_joins; ++j) {
    const hdk_b200_join& s = plan->joins[j];
    if (s.key_expr < 0 || s.key_expr >= plan->n_exprs) { set_error("bad join key node"); return HDK_B200_E_INVALID; }
    DJoin& d = p.joins[j];
    d.min_key = s.min_key; d.max_key = s.max_key; d.null_val = s.null_val;
    d.key_expr = s.key_expr; d.key_nullable = uint8_t(s.key_nullable); d.one_to_many = uint8_t(s.one_to_many);
    if (s.payload_by_slot && s.one_to_many) { set_error("payload_by_slot needs a one-to-one table"); return HDK_B200_E_INVALID; }
    if (s.payload_by_slot < 0 || s.payload_by_slot > 2) { set_error("payload_by_slot must be 0, 1 or 2"); return HDK_B200_E_INVALID; }
    d.by_slot = uint8_t(s.payload_by_slot);
    p.join_entry_count[j] = s.entry_count;
    if (s.n_key_exprs < 0 || s.n_key_exprs > HDK_B200_MAX_KEYS) { set_error("join %d: bad n_key_exprs", j); return HDK_B200_E_INVALID; }
    if (s.n_key_exprs > 0) {
      if (s.payload_by_slot) { set_error("join %d: payload_by_slot is a perfect-table layout", j); return HDK_B200_E_INVALID; }
      if (s.key_width != 4 && s.key_width != 8) { set_error("join %d: key_width must be 4 or 8", j); return HDK_B200_E_INVALID; }
      d.n_key_exprs = uint8_t(s.n_key_exprs);
      d.key_width = uint8_t(s.key_width);
      for (int i = 0; i < s.n_key_exprs; ++i) {
        if (s.key_exprs[i] < 0 || s.key_exprs[i] > s.key_expr) { set_error("join %d: component %d must not come after key_expr", j, i); return HDK_B200_E_INVALID; }
        if (plan->exprs[s.key_exprs[i]].type.kind != HDK_B200_INT) { set_error("join %d: floating-point join keys are not supported", j); return HDK_B200_E_UNSUPPORTED; }
        d.key_exprs[i] = int8_t(s.key_exprs[i]);
      }
    }
    // every inner-table column of join j must come after the join's key node
    for (int i = 0; i <= s.key_expr; ++i)
      if (plan->exprs[i].op == HDK_B200_OP_COL && plan->exprs[i].a == j + 1) { set_error("inner column of join %d precedes its key node", j); return HDK_B200_E_INVALID; }
  }
  // ---- keys
  int64_t mult = 1;
  for (int k = 0; k < plan->n_keys; ++k) {
    const hdk_b200_key& s = plan->keys[k];
    if (s.expr < 0 || s.expr >= plan->n_exprs) { set_error("bad key node"); return HDK_B200_E_INVALID; }
    if (plan->exprs[s.expr].type.kind != HDK_B200_INT) { set_error("floating-point group keys are not supported"); return HDK_B200_E_UNSUPPORTED; }
    DKey& d = p.keys[k];
    d.expr = s.expr;
    d.width = uint8_t(plan->exprs[s.expr].type.width);
    d.has_nulls = uint8_t(s.has_nulls != 0);
    d.min_val = s.min_val;
    d.null_translated = s.max_val + (s.bucket ? s.bucket : 1);
    d.mult = mult;
    d.card = s.cardinality;
    d.bucket = s.bucket > 1 ? s.bucket : 0;
    if (q->hash_type == HDK_B200_PERFECT_HASH) {
      if (s.bucket < 0) { set_error("key %d: negative bucket", k); return HDK_B200_E_INVALID; }
      if (s.cardinality <= 0) { set_error("key %d: bad cardinality", k); return HDK_B200_E_INVALID; }
      mult *= s.cardinality;
    }
  }
  if (plan->n_keys == 0 && !(q->hash_type == HDK_B200_PERFECT_HASH && q->keyless && q->entry_count == 1 && !q->output_columnar)) {
    // non-grouped aggregate (QueryDescriptionType::NonGroupedAggregate): zero keys = one keyless row-wise entry
    set_error("a plan without group keys needs a keyless perfect-hash descriptor with one entry");
    return HDK_B200_E_INVALID;
  }
  if (q->hash_type == HDK_B200_PERFECT_HASH) {
    if (plan->n_keys == 1) {
      if (uint64_t(plan->keys[0].cardinality) > q->entry_count) { set_error("perfect hash: key cardinality exceeds entry_count"); return HDK_B200_E_INVALID; }
    } else if (uint64_t(mult) > q->entry_count) { set_error("perfect hash: key space exceeds entry_count"); return HDK_B200_E_INVALID; }
    if (q->key_width != 8) { set_error("perfect hash needs 8-byte keys"); return HDK_B200_E_INVALID; }
  } else if (q->hash_type == HDK_B200_BASELINE_HASH) {
    if (q->key_width != 4 && q->key_width != 8) { set_error("baseline key width must be 4 or 8"); return HDK_B200_E_INVALID; }
    if (q->keyless) { set_error("baseline hash cannot be keyless"); return HDK_B200_E_INVALID; }
    if (q->output_columnar && q->key_width != 8) { set_error("columnar baseline needs 8-byte keys"); return HDK_B200_E_INVALID; }
  } else { set_error("unknown hash type"); return HDK_B200_E_UNSUPPORTED; }

  // ---- layout
  L.entry_count = q->entry_count;
  L.key_count = q->key_count;
  L.key_width = q->key_width;
  L.keyless = q->keyless;
  L.columnar = q->output_columnar;
  L.slot_count = q->slot_count;
  L.target_idx_for_key = q->target_idx_for_key;
  size_t rw_off = 0, col_off = 0;
  const size_t E = q->entry_count;
  if (q->output_columnar && !q->keyless) col_off = size_t(q->key_count) * align8(8 * E);
  L.key_bytes = uint32_t(q->keyless ? 0 : align8(size_t(q->key_count) * q->key_width));
  for (int s = 0; s < q->slot_count; ++s) {
    const int w = q->slot_padded[s];
    if (w != 0 && w != 4 && w != 8) { set_error("slot %d: padded width %d unsupported", s, w); return HDK_B200_E_UNSUPPORTED; }
    if (w == 8) rw_off = align8(rw_off);
    L.slots[s].off = int32_t(rw_off);
    L.slots[s].col_off = col_off;
    L.slots[s].padded = uint8_t(w);
    L.slots[s].init_val = q->init_vals[s];
    L.slots[s].acc = L.slots[s].acc_cnt = -1;
    L.slots[s].arg = -1;
    rw_off += w;
    col_off += align8(size_t(w) * E);
  }
  L.row_bytes = uint32_t(align8(L.key_bytes + rw_off));
  if (q->keyless && (q->target_idx_for_key < 0 || q->target_idx_for_key >= q->slot_count)) { set_error("keyless: bad target_idx_for_key"); return HDK_B200_E_INVALID; }

  // ---- targets → slots and accumulators
  p.n_acc = 0;
  find_or_add_acc(p, ACC_CNT_ALL, -1, 0);  // accumulator 0: rows per group (emptiness, COUNT(*))
  bool slot_used[HDK_B200_MAX_SLOTS] = {false};
  for (int t = 0; t < plan->n_targets; ++t) {
    const hdk_b200_target& tg = plan->targets[t];
    if (tg.slot < 0) {
      if (tg.agg != HDK_B200_AGG_NONE) { set_error("target %d: aggregate without a slot", t); return HDK_B200_E_INVALID; }
      continue;
    }
    const int nslots = tg.agg == HDK_B200_AGG_AVG ? 2 : 1;
    if (tg.slot + nslots > q->slot_count) { set_error("target %d: slot out of range", t); return HDK_B200_E_INVALID; }
    if (q->slot_padded[tg.slot] == 0) continue;
    const bool has_arg = tg.arg >= 0;
    if (tg.agg != HDK_B200_AGG_NONE && tg.agg != HDK_B200_AGG_COUNT && !has_arg) { set_error("target %d: aggregate needs an argument", t); return HDK_B200_E_INVALID; }
    if (has_arg && tg.arg >= plan->n_exprs) { set_error("target %d: bad arg node", t); return HDK_B200_E_INVALID; }
    for (int which = 0; which < nslots; ++which) {
      DSlot& s = L.slots[tg.slot + which];
      slot_used[tg.slot + which] = true;
      const int padded = q->slot_padded[tg.slot + which];
      hdk_b200_type arg_type = tg.type;
      if (tg.agg == HDK_B200_AGG_NONE) {
        if (tg.key_index < 0 || tg.key_index >= plan->n_keys) { set_error("target %d: bad key_index", t); return HDK_B200_E_INVALID; }
        arg_type = plan->exprs[plan->keys[tg.key_index].expr].type;
      } else if (has_arg) {
        arg_type = plan->exprs[tg.arg].type;
      }
      const bool is_fp_arg = has_arg && arg_type.kind == HDK_B200_FP;
      const bool float_arg_input = (tg.agg == HDK_B200_AGG_AVG || tg.agg == HDK_B200_AGG_SUM || tg.agg == HDK_B200_AGG_MIN ||
                                    tg.agg == HDK_B200_AGG_MAX) && is_fp_arg && arg_type.width == 4;
      const bool is_count_in_avg = tg.agg == HDK_B200_AGG_AVG && which == 1;
      s.bytes = uint8_t((float_arg_input && !is_count_in_avg) ? 4 : padded);
      s.skip_null = uint8_t(tg.skip_null_val && has_arg && tg.agg != HDK_B200_AGG_NONE);
      s.arg = int16_t(has_arg ? tg.arg : -1);
      s.arg_kind = uint8_t(arg_type.kind);
      s.arg_width = uint8_t(arg_type.width);
      s.arg_nullable = uint8_t(arg_type.nullable);
      s.key_index = tg.key_index;
      const uint8_t skip = s.skip_null ? 1 : 0;
      if (tg.agg == HDK_B200_AGG_NONE) {
        s.op = SLOT_KEY;
        s.arg = int16_t(plan->keys[tg.key_index].expr);
        s.key_width = uint8_t(arg_type.width);
        s.key_nullable = uint8_t(arg_type.nullable);
        s.is_fp = 0;
      } else if (tg.agg == HDK_B200_AGG_COUNT || is_count_in_avg) {
        s.op = SLOT_COUNT;
        s.is_fp = 0;
        if (skip) {
          // reference quirk (convertNullIfAny, QE/RowFuncBuilder.cpp:803-860): COUNT(int64 arg) with an
          // int32 COUNT type compares the value truncated to 32 bits with INT32_MIN.
          uint8_t mode = 1;
          if (tg.agg == HDK_B200_AGG_COUNT && !is_fp_arg && arg_type.width == 8 && tg.type.width == 4) mode = 2;
          s.count_mode = mode;
          s.acc = int8_t(find_or_add_acc(p, ACC_CNT_NN, tg.arg, mode));
        } else {
          s.acc = 0;
        }
      } else {
        const bool is_sum = tg.agg == HDK_B200_AGG_SUM || tg.agg == HDK_B200_AGG_AVG;
        s.op = is_sum ? SLOT_SUM : (tg.agg == HDK_B200_AGG_MIN ? SLOT_MIN : SLOT_MAX);
        s.is_fp = uint8_t(is_fp_arg);
        s.is_avg_sum = uint8_t(tg.agg == HDK_B200_AGG_AVG);
        uint8_t kind = is_sum ? (is_fp_arg ? ACC_SUM_F : ACC_SUM_I)
                              : (tg.agg == HDK_B200_AGG_MIN ? (is_fp_arg ? ACC_MIN_F : ACC_MIN_I)
                                                            : (is_fp_arg ? ACC_MAX_F : ACC_MAX_I));
        s.acc = int8_t(find_or_add_acc(p, kind, tg.arg, skip));
        if (skip) s.acc_cnt = int8_t(find_or_add_acc(p, ACC_CNT_NN, tg.arg, 1));
        if (!is_fp_arg && s.bytes != 8) { set_error("target %d: integer aggregate in a %d-byte slot is not produced by the reference layout", t, s.bytes); return HDK_B200_E_UNSUPPORTED; }
      }
      if (s.acc < -0 && s.op != SLOT_KEY) { set_error("too many accumulators"); return HDK_B200_E_UNSUPPORTED; }
      if ((s.op != SLOT_KEY && s.acc < 0) || (skip && s.op >= SLOT_SUM && s.acc_cnt < 0)) { set_error("too many accumulators"); return HDK_B200_E_UNSUPPORTED; }
    }
  }
  for (int s = 0; s < q->slot_count; ++s)
    if (q->slot_padded[s] && !slot_used[s]) { set_error("slot %d is not produced by any target", s); return HDK_B200_E_INVALID; }

  // order accumulators by merge class [sum_i | sum_f | min | max], keeping CNT_ALL at index 0
  {
    DAcc sorted[kMaxAcc];
    int remap[kMaxAcc];
    int n = 0;
    int counts[4] = {0, 0, 0, 0};
    for (int cls = 0; cls < 4; ++cls)
      for (int i = 0; i < p.n_acc; ++i)
        if (acc_class(p.accs[i].kind) == cls) { remap[i] = n; sorted[n++] = p.accs[i]; ++counts[cls]; }
    memcpy(p.accs, sorted, sizeof(DAcc) * p.n_acc);
    for (int s = 0; s < q->slot_count; ++s) {
      if (L.slots[s].acc >= 0) L.slots[s].acc = int8_t(remap[L.slots[s].acc]);
      if (L.slots[s].acc_cnt >= 0) L.slots[s].acc_cnt = int8_t(remap[L.slots[s].acc_cnt]);
    }
    out->n_sum_i = counts[0]; out->n_sum_f = counts[1]; out->n_min = counts[2]; out->n_max = counts[3];
  }
  out->work_table_bytes = size_t(p.n_acc) * E * 8;   // perfect: [n_acc][E]; baseline: [E][n_acc]
  return HDK_B200_OK;
}

// Structural signature of a lowered plan: everything the scan kernel's control flow depends on, nothing
// that is data (literals, key ranges, entry counts).  Pre-compiled shapes are matched by it.
uint64_t plan_signature(const DPlan& p) {
  uint64_t h = 1469598103934665603ULL;
  auto mix = [&](uint64_t v) { for (int i = 0; i < 8; ++i) { h ^= (v >> (8 * i)) & 0xff; h *= 1099511628211ULL; } };
  mix(p.n_exprs); mix(p.n_filters); mix(p.n_keys); mix(p.n_joins); mix(p.n_acc); mix(p.n_cols); mix(p.hash_type);
  for (int i = 0; i < p.n_exprs; ++i) {
    const DExpr& e = p.exprs[i];
    mix(e.op); mix(uint8_t(e.a)); mix(uint8_t(e.b)); mix(e.aux); mix(e.kind); mix(e.width); mix(e.nullable); mix(e.guard);
    if (e.op == HDK_B200_OP_COL || e.op == HDK_B200_OP_EXTRACT_YEAR || e.op == HDK_B200_OP_CASE) mix(uint64_t(e.imm.i));
  }
  for (int i = 0; i < p.n_filters; ++i) mix(uint8_t(p.filters[i]));
  for (int i = 0; i < p.n_keys; ++i) { mix(p.keys[i].expr); mix(p.keys[i].has_nulls); mix(p.keys[i].width); }
  for (int i = 0; i < p.n_joins; ++i) {
    mix(p.joins[i].key_expr); mix(p.joins[i].key_nullable); mix(p.joins[i].one_to_many); mix(p.joins[i].n_key_exprs);
    for (int k = 0; k < p.joins[i].n_key_exprs; ++k) mix(uint8_t(p.joins[i].key_exprs[k]));
  }
  for (int i = 0; i < p.n_acc; ++i) { mix(p.accs[i].kind); mix(uint8_t(p.accs[i].arg)); mix(p.accs[i].arg_nullable); mix(p.accs[i].bytes); }
  for (int i = 0; i < p.n_cols; ++i) mix(p.col_width[i]);
  return h;
}

// Field-by-field equality of what plan_signature hashes: the run-time kernel cache compares a hit with the shape it was
// compiled for, so a 64-bit signature collision degrades to the interpreter instead of running another shape's kernel.
bool same_plan_shape(const DPlan& a, const DPlan& b) {
  if (a.n_exprs != b.n_exprs || a.n_filters != b.n_filters || a.n_keys != b.n_keys || a.n_joins != b.n_joins || a.n_acc != b.n_acc ||
      a.n_cols != b.n_cols || a.hash_type != b.hash_type)
    return false;
  for (int i = 0; i < a.n_exprs; ++i) {
    const DExpr &x = a.exprs[i], &y = b.exprs[i];
    if (x.op != y.op || x.a != y.a || x.b != y.b || x.aux != y.aux || x.kind != y.kind || x.width != y.width || x.nullable != y.nullable ||
        x.guard != y.guard)
      return false;
    if ((x.op == HDK_B200_OP_COL || x.op == HDK_B200_OP_EXTRACT_YEAR || x.op == HDK_B200_OP_CASE) && x.imm.i != y.imm.i) return false;
  }
  for (int i = 0; i < a.n_filters; ++i)
    if (a.filters[i] != b.filters[i]) return false;
  for (int i = 0; i < a.n_keys; ++i)
    if (a.keys[i].expr != b.keys[i].expr || a.keys[i].has_nulls != b.keys[i].has_nulls || a.keys[i].width != b.keys[i].width) return false;
  for (int i = 0; i < a.n_joins; ++i) {
    const DJoin &x = a.joins[i], &y = b.joins[i];
    if (x.key_expr != y.key_expr || x.key_nullable != y.key_nullable || x.one_to_many != y.one_to_many || x.n_key_exprs != y.n_key_exprs) return false;
    for (int k = 0; k < x.n_key_exprs; ++k)
      if (x.key_exprs[k] != y.key_exprs[k]) return false;
  }
  for (int i = 0; i < a.n_acc; ++i)
    if (a.accs[i].kind != b.accs[i].kind || a.accs[i].arg != b.accs[i].arg || a.accs[i].arg_nullable != b.accs[i].arg_nullable || a.accs[i].bytes != b.accs[i].bytes)
      return false;
  for (int i = 0; i < a.n_cols; ++i)
    if (a.col_width[i] != b.col_width[i]) return false;
  return true;
}

// C++ aggregate initialiser of the structural part of a DPlan (consumed by static_shapes.inc and by jit.cu)
int dump_shape_text(const DPlan& p, char* out, size_t cap);
static int dump_shape(const DPlan& p, char* out, size_t cap) { return dump_shape_text(p, out, cap); }
int dump_shape_text(const DPlan& p, char* out, size_t cap) {
  std::string s = "{";
  auto add = [&](const char* fmt, ...) { char b[256]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof(b), fmt, ap); va_end(ap); s += b; };
  add("%d, %d, %d, %d, %d, %d, 0u, %d,\n  {", p.n_exprs, p.n_filters, p.n_keys, p.n_joins, p.n_acc, p.n_cols, p.hash_type);
  for (int i = 0; i < p.n_exprs; ++i) {
    const DExpr& e = p.exprs[i];
    const long long imm = (e.op == HDK_B200_OP_COL || e.op == HDK_B200_OP_EXTRACT_YEAR || e.op == HDK_B200_OP_CASE) ? (long long)e.imm.i : 0;
    add("{%d, %d, %d, %d, %d, %d, %d, %d, {%lldLL}}%s", e.op, e.a, e.b, e.aux, e.kind, e.width, e.nullable, e.guard, imm, i + 1 < p.n_exprs ? ", " : "");
  }
  s += "},\n  {";
  for (int i = 0; i < HDK_B200_MAX_FILTERS; ++i) add("%d%s", i < p.n_filters ? p.filters[i] : 0, i + 1 < HDK_B200_MAX_FILTERS ? ", " : "");
  s += "},\n  {";
  for (int i = 0; i < p.n_keys; ++i) add("{0, 0, 0, 0, %d, %d, %d, 0, 0}%s", p.keys[i].expr, p.keys[i].has_nulls, p.keys[i].width, i + 1 < p.n_keys ? ", " : "");
  s += "},\n  {";
  for (int i = 0; i < p.n_joins; ++i) {
    // (baseline join tables — composite / wide-range keys — probe with their component nodes: part of the structure; the
    //  component width stays a run-time property like the key range)
    const DJoin& j = p.joins[i];
    add("{0, 0, 0, %d, %d, %d, 0, 0, %d, 0, 0, 0, {", j.key_expr, j.key_nullable, j.one_to_many, j.n_key_exprs);
    for (int k = 0; k < HDK_B200_MAX_KEYS; ++k) add("%d%s", k < j.n_key_exprs ? j.key_exprs[k] : 0, k + 1 < HDK_B200_MAX_KEYS ? ", " : "");
    add("}, 0}%s", i + 1 < p.n_joins ? ", " : "");
  }
  s += "},\n  {";
  for (int i = 0; i < p.n_acc; ++i) add("{%d, %d, %d, %d}%s", p.accs[i].kind, p.accs[i].arg, p.accs[i].arg_nullable, p.accs[i].bytes, i + 1 < p.n_acc ? ", " : "");
  s += "},\n  {";
  for (int i = 0; i < HDK_B200_MAX_COLS; ++i) add("%d%s", i < p.n_cols ? p.col_width[i] : 0, i + 1 < HDK_B200_MAX_COLS ? ", " : "");
  s += "},\n  {0, 0, 0, 0}}";
  if (s.size() + 1 > cap) return -1;
  memcpy(out, s.c_str(), s.size() + 1);
  return int(s.size());
}

}  // namespace hb

extern "C" {
// build-time helper of tools/gen_static_shapes.py (not part of the public header): signature + initialiser text
__attribute__((visibility("default"))) int hdk_b200_internal_dump_shape(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd,
                                                                        uint64_t* sig, int* n_exprs, char* out, size_t cap) {
  hb::Lowered lw;
  if (int rc = hb::lower_plan(plan, qmd, &lw)) return rc;
  if (sig) *sig = hb::plan_signature(lw.plan);
  if (n_exprs) *n_exprs = lw.plan.n_exprs;
  return hb::dump_shape(lw.plan, out, cap) < 0 ? HDK_B200_E_NOMEM : HDK_B200_OK;
}

const char* hdk_b200_last_error(void) { return hb::last_error(); }

int hdk_b200_debug_set(const char* name, int value) {
  if (!name) { hb::set_error("null knob name"); return HDK_B200_E_INVALID; }
  const std::string n(name);
  if (n == "force_generic") hb::g_debug.force_generic = value != 0;
  else if (n == "force_strategy") hb::g_debug.force_strategy = value;
  else if (n == "partitioned_aggregation") hb::g_debug.partitioned = value;
  else if (n == "partitioned_table_slots") hb::g_debug.pa_slots = value;
  else if (n == "partitioned_partitions") hb::g_debug.pa_partitions = value;
  else if (n == "partitioned_heavy_rows") hb::g_debug.pa_heavy_rows = value;
  else if (n == "jit") hb::g_debug.jit = value;
  else if (n == "geo_env_refresh") hb::g_debug.geo_env_refresh = value;
  else { hb::set_error("unknown debug knob '%s'", name); return HDK_B200_E_INVALID; }
  return HDK_B200_OK;
}
int hdk_b200_abi_version(void) { return HDK_B200_ABI_VERSION; }
uint64_t hdk_b200_launch_count(void) { return hb::g_launch_count; }
int hdk_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int hdk_b200_plan_check(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, size_t* scratch_bytes) {
  hb::Lowered lw;
  const int rc = hb::lower_plan(plan, qmd, &lw);
  if (rc) return rc;
  if (scratch_bytes) *scratch_bytes = lw.work_table_bytes;
  return HDK_B200_OK;
}

size_t hdk_b200_buffer_size_bytes(const hdk_b200_qmd* q) {
  if (!q) return 0;
  const size_t E = q->entry_count;
  auto a8 = [](size_t x) { return (x + 7) & ~size_t(7); };
  if (q->output_columnar) {
    size_t off = q->keyless ? 0 : size_t(q->key_count) * a8(8 * E);
    for (int s = 0; s < q->slot_count; ++s) off += a8(size_t(q->slot_padded[s]) * E);
    return a8(off);
  }
  size_t off = 0;
  for (int s = 0; s < q->slot_count; ++s) {
    if (q->slot_padded[s] == 8) off = a8(off);
    off += q->slot_padded[s];
  }
  const size_t key_bytes = q->keyless ? 0 : a8(size_t(q->key_count) * q->key_width);
  return a8(key_bytes + off) * E;
}

int hdk_b200_work_table_layout_get(const hdk_b200_plan* plan, const hdk_b200_qmd* qmd, hdk_b200_work_table_layout* out) {
  hb::Lowered lw;
  const int rc = hb::lower_plan(plan, qmd, &lw);
  if (rc) return rc;
  if (qmd->hash_type != HDK_B200_PERFECT_HASH) { hb::set_error("work tables exist for perfect-hash plans only"); return HDK_B200_E_UNSUPPORTED; }
  const uint64_t E = qmd->entry_count;
  out->n_cells = uint64_t(lw.plan.n_acc) * E;
  out->sum_i64_cells = uint64_t(lw.n_sum_i) * E;
  out->sum_cells = uint64_t(lw.n_sum_i + lw.n_sum_f) * E;
  out->min_cells = uint64_t(lw.n_min) * E;
  out->max_cells = uint64_t(lw.n_max) * E;
  return HDK_B200_OK;
}
}
