// hdk_b200/csrc/scan_kernel.cuh — device code of the fused scan → filter → join probe → group-by → aggregate kernel
// (see scan.cu for the execution model).  A header so that it can be instantiated twice: at build time for the generic
// (interpreting) shape and the pre-compiled shapes of static_shapes.inc (scan.cu), and at run time by NVRTC for the shape
// of any other plan (jit.cu).
#pragma once
#include "common.cuh"
#include "accum.cuh"
#include "baseline.cuh"
#include "device_utils.cuh"
#include "eval.cuh"
#include "scan.cuh"
#include "shape.cuh"

namespace hb {

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
struct StageHeader {          // written by the producer, read by consumers
  uint32_t rows;              // rows in this tile
  uint32_t aligned;           // every column slice of the tile starts on a 16-byte boundary (vector loads allowed)
  uint32_t col_off[HDK_B200_MAX_COLS];  // byte offset (from dynamic smem base) of element 0 of column c
};

__device__ __forceinline__ uint64_t lds_elem(const uint8_t* ptr, int w) {
  return w == 8 ? *reinterpret_cast<const uint64_t*>(ptr) : w == 4 ? uint64_t(*reinterpret_cast<const uint32_t*>(ptr))
         : w == 2 ? uint64_t(*reinterpret_cast<const uint16_t*>(ptr)) : uint64_t(*ptr);
}
// one <= 16-byte shared-memory load holding several consecutive elements of a column, and element extraction
template <int BYTES>
__device__ __forceinline__ void lds_vec(const uint8_t* p, uint32_t* words) {
  if constexpr (BYTES == 16) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    words[0] = t.x; words[1] = t.y; words[2] = t.z; words[3] = t.w;
  } else if constexpr (BYTES == 8) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    words[0] = t.x; words[1] = t.y;
  } else if constexpr (BYTES == 4) {
    words[0] = *reinterpret_cast<const uint32_t*>(p);
  } else if constexpr (BYTES == 2) {
    words[0] = *reinterpret_cast<const uint16_t*>(p);
  } else {
    words[0] = *p;
  }
}
template <int W>
__device__ __forceinline__ uint64_t vec_elem(const uint32_t* words, int v) {
  if constexpr (W == 8) return uint64_t(words[2 * v]) | (uint64_t(words[2 * v + 1]) << 32);
  else if constexpr (W == 4) return words[v];
  else if constexpr (W == 2) return (words[v / 2] >> (16 * (v % 2))) & 0xffffu;
  else return (words[v / 4] >> (8 * (v % 4))) & 0xffu;
}
__host__ __device__ constexpr int shape_max_width(const DPlan& p) {
  int m = 1;
  for (int c = 0; c < p.n_cols; ++c) m = p.col_width[c] > m ? p.col_width[c] : m;
  return m;
}
__host__ __device__ constexpr int shape_vec_rows(const DPlan& p) { return 16 / shape_max_width(p); }
// rows a consumer thread handles per iteration of the full-tile loop: a multiple of the vector group
template <class Shape>
__host__ __device__ constexpr int shape_iter_rows() {
  constexpr int vw = shape_vec_rows(Shape::get());
  return Shape::rows_per_iter > vw ? Shape::rows_per_iter / vw * vw : vw;
}

__device__ __forceinline__ uint64_t ldg_elem(const uint8_t* ptr, int w) {
  return w == 8 ? __ldg(reinterpret_cast<const uint64_t*>(ptr)) : w == 4 ? uint64_t(__ldg(reinterpret_cast<const uint32_t*>(ptr)))
         : w == 2 ? uint64_t(__ldg(reinterpret_cast<const uint16_t*>(ptr))) : uint64_t(__ldg(ptr));
}

// group slot of a row (perfect hash): get_group_value_fast / perfect_key_hash incl. translate_null_key
// (QE/GroupByRuntime.cpp:198-213, QE/RowFuncBuilder.cpp:748-801)
template <int kStrategy>
__device__ __forceinline__ void accumulate_one(const ScanArgs& args, uint8_t* bins, int tid, int a, const DAcc acc, uint32_t idx,
                                               int64_t x) {
  if (kStrategy == HDK_B200_STRATEGY_THREAD_PRIVATE) {
    bin_update_private(acc.kind, bins + args.acc_bin_off[a] + (idx * args.consumer_threads + uint32_t(tid)) * acc.bytes, x);
  } else if (kStrategy == HDK_B200_STRATEGY_CTA_SHARED) {
    bin_update_shared_atomic(acc.kind, bins + args.acc_bin_off[a] + size_t(idx) * acc.bytes, x);
  } else if (kStrategy == HDK_B200_STRATEGY_BASELINE) {
    // entry-major work table: the accumulators of one hash entry share a sector
    cell_update_global(acc.kind, args.work_table + size_t(idx) * args.plan.n_acc + a, x);
  } else {
    cell_update_global(acc.kind, args.work_table + size_t(a) * args.plan.entry_count + idx, x);
  }
}

// baseline hash: find / claim the row's entry in the reference-encoded buffer (keys are written there by the claim);
// the aggregates go to the neutral work table cell block of that entry and are encoded by the finalize kernel.
// Returns the entry index or -1 when the table is full (get_group_value returning NULL ⇒ ERR_OUT_OF_SLOTS).
template <class KeyExpr>
__device__ __forceinline__ int64_t baseline_entry(const ScanArgs& args, int n_keys, KeyExpr&& key_expr, const V* vals) {
  const DLayout& L = args.layout;
  int8_t* buf = reinterpret_cast<int8_t*>(args.groupby_buf[0]);
  const uint32_t E = args.plan.entry_count;
  int64_t keys[HDK_B200_MAX_KEYS];
#pragma unroll
  for (int k = 0; k < HDK_B200_MAX_KEYS; ++k) {
    if (k >= n_keys) break;
    const int64_t v = vals[key_expr(k)].i;
    keys[k] = L.key_width == 4 ? int64_t(int32_t(v)) : v;  // castToTypeIn(key, key_width * 8), no NULL translation
  }
  const uint32_t h0 = key_hash_dev(keys, n_keys, L.key_width) % E;
  return L.columnar ? baseline_claim_columnar(reinterpret_cast<int64_t*>(buf), E, keys, n_keys, h0)
         : L.key_width == 4 ? baseline_claim_rowwise<int32_t>(buf, L.row_bytes, E, keys, n_keys, h0)
                            : baseline_claim_rowwise<int64_t>(buf, L.row_bytes, E, keys, n_keys, h0);
}

// ---- one row, run-time plan --------------------------------------------------------------------
// probe of join j with the key node(s) just evaluated: false on a miss (inner join: the row, or this combination of
// matches, is dropped).  One-to-one tables give one inner row id; one-to-many tables the matching set (ids != nullptr).
__device__ __forceinline__ bool probe_join_generic(const ScanArgs& args, int j, const V* vals, int64_t& rid, int32_t& n_matches,
                                                   const int32_t*& ids) {
  const DPlan& p = args.plan;
  const DJoin& jn = p.joins[j];
  n_matches = 1;
  ids = nullptr;
  if (jn.n_key_exprs) {   // composite / wide-range key: baseline join table
    int64_t k64[HDK_B200_MAX_KEYS];
    for (int i = 0; i < jn.n_key_exprs; ++i) k64[i] = vals[jn.key_exprs[i]].i;
    const int8_t* tbl = reinterpret_cast<const int8_t*>(args.join_hash_tables[j]);
    const int64_t E = p.join_entry_count[j];
    if (jn.one_to_many) {
      // composite-key dictionary, then offsets | counts | payload indexed by the key's position in it
      // (BaselineJoinHashTable one-to-many layout; HashJoin::codegenMatchingSet)
      const int64_t slot = jn.key_width == 4 ? baseline_dict_index<int32_t>(tbl, E, k64, jn.n_key_exprs)
                                             : baseline_dict_index<int64_t>(tbl, E, k64, jn.n_key_exprs);
      if (slot < 0) return false;
      const int32_t* otm = reinterpret_cast<const int32_t*>(tbl + size_t(E) * size_t(jn.n_key_exprs) * size_t(jn.key_width));
      const int32_t off = __ldg(otm + slot);
      if (off < 0) return false;
      n_matches = __ldg(otm + E + slot);
      ids = otm + 2 * E + off;
      return n_matches > 0;
    }
    rid = jn.key_width == 4 ? baseline_join_probe<int32_t>(tbl, E, k64, jn.n_key_exprs) : baseline_join_probe<int64_t>(tbl, E, k64, jn.n_key_exprs);
    return rid >= 0;
  }
  // hash_join_idx[_nullable] (QE/GroupByRuntime.cpp:298-329)
  const int64_t key = vals[jn.key_expr].i;
  if ((jn.key_nullable && key == jn.null_val) || key < jn.min_key || key > jn.max_key) return false;
  const int32_t* table = reinterpret_cast<const int32_t*>(args.join_hash_tables[j]);
  const int64_t slot = key - jn.min_key;
  if (jn.one_to_many) {
    // offsets | counts | payload (JHT/PerfectJoinHashTable.cpp:861-886)
    const int64_t E = p.join_entry_count[j];
    const int32_t off = __ldg(table + slot);
    if (off < 0) return false;
    n_matches = __ldg(table + E + slot);
    ids = table + 2 * E + off;
    return n_matches > 0;
  }
  if (jn.by_slot) {
    // presence bitmap + slot-ordered inner columns (hdk_b200_gather_join_payload_on_device)
    if (jn.by_slot == 1 && !((__ldg(reinterpret_cast<const uint32_t*>(table) + (slot >> 5)) >> (slot & 31)) & 1u)) return false;
    rid = slot;   // (by_slot == 2: every slot of [min_key, max_key] is occupied, no bitmap)
    return true;
  }
  rid = __ldg(table + slot);
  return rid >= 0;
}

// The joins of a plan are nested loops (the reference's JoinLoop nest, QE/IRCodegen.cpp: a Singleton loop for a one-to-one
// table, a Set loop over HashJoin::codegenMatchingSet's (offset, count) for a one-to-many table).  Level l evaluates the
// nodes between the key of join l-1 and the key of join l (joins taken by key node: args.join_order), then probes join l;
// the innermost level evaluates the rest, filters and accumulates once per combination of matches.  Iterative: cur[l] is
// the position inside level l's matching set, a finished level advances the nearest outer one that has matches left.
template <int kStrategy>
__device__ __forceinline__ void process_row_generic(const ScanArgs& args, const uint8_t* smem, const uint32_t* col_off, uint32_t r,
                                                    uint8_t* bins, int tid, V* vals, int32_t& my_err) {
  const DPlan& p = args.plan;
  const int J = p.n_joins;
  int64_t rowid[HDK_B200_MAX_JOINS];
  int32_t n_matches[HDK_B200_MAX_JOINS], cur[HDK_B200_MAX_JOINS];
  const int32_t* match_ids[HDK_B200_MAX_JOINS];
  int32_t row_err_at[HDK_B200_MAX_JOINS + 1], qual_err_at[HDK_B200_MAX_JOINS + 1];   // first error up to and including a level
  auto load_outer = [&](int c, int w) -> uint64_t { return lds_elem(smem + col_off[c] + size_t(r) * w, w); };
  auto load_inner = [&](int j, int c, int w) -> uint64_t {
    return ldg_elem(reinterpret_cast<const uint8_t*>(args.inner_col_buffers[j * HDK_B200_MAX_COLS + c]) + size_t(rowid[j]) * w, w);
  };
  int lvl = 0;
  for (;;) {
    bool dead = false;
    for (; lvl <= J; ++lvl) {
      const int n0 = lvl == 0 ? 0 : p.joins[args.join_order[lvl - 1]].key_expr + 1;
      const int n1 = lvl == J ? p.n_exprs : p.joins[args.join_order[lvl]].key_expr + 1;
      if (lvl > 0 && match_ids[lvl - 1]) rowid[args.join_order[lvl - 1]] = __ldg(match_ids[lvl - 1] + cur[lvl - 1]);
      int32_t row_err = lvl ? row_err_at[lvl - 1] : 0, qual_err = lvl ? qual_err_at[lvl - 1] : 0;
      for (int n = n0; n < n1; ++n) {
        int32_t e = 0;
        vals[n] = eval_node(p, p.exprs[n], vals, e, load_outer, load_inner);
        if (e) { int32_t& dst = (p.exprs[n].aux & kAuxInQual) ? qual_err : row_err; if (!dst) dst = e; }
      }
      row_err_at[lvl] = row_err;
      qual_err_at[lvl] = qual_err;
      if (lvl < J) {
        cur[lvl] = 0;
        const int j = args.join_order[lvl];
        if (!probe_join_generic(args, j, vals, rowid[j], n_matches[lvl], match_ids[lvl])) { dead = true; break; }
      }
    }
    if (!dead) {
      const int32_t row_err = row_err_at[J], qual_err = qual_err_at[J];
      bool pass = true;
      for (int f = 0; f < p.n_filters; ++f) pass = pass && (vals[p.filters[f]].i > 0);
      if (qual_err) {
        my_err = my_err > 0 ? my_err : qual_err;   // an error inside a qual: raised whether or not the row passes
      } else if (pass && row_err) {
        my_err = my_err > 0 ? my_err : row_err;
      } else if (pass) {
        uint32_t idx = 0;
        bool have = true;
        if (kStrategy == HDK_B200_STRATEGY_BASELINE) {
          const int64_t entry = baseline_entry(args, p.n_keys, [&](int k) { return p.keys[k].expr; }, vals);
          if (entry == kClaimTimedOut) { my_err = my_err > 0 ? my_err : HDK_B200_ERR_CLAIM_TIMEOUT; have = false; }
          else if (entry < 0) { if (my_err <= 0) my_err = -HDK_B200_ERR_OUT_OF_SLOTS; have = false; }
          idx = uint32_t(entry);
        } else {
          int64_t h = 0;
          for (int k = 0; k < p.n_keys; ++k) {
            const DKey& ky = p.keys[k];
            int64_t v = vals[ky.expr].i;
            if (ky.has_nulls && v == int_null_of(ky.width)) v = ky.null_translated;
            int64_t term = v - ky.min_val;
            if (ky.bucket) term /= ky.bucket;   // (get_group_value_fast / perfect_key_hash divide by the bucket)
            h += term * ky.mult;
          }
          idx = uint32_t(h);
          if (idx >= p.entry_count) { my_err = my_err > 0 ? my_err : 1003; have = false; }  // key outside the range the layout was built for
        }
        if (have) {
          for (int a = 0; a < p.n_acc; ++a) {
            const DAcc acc = p.accs[a];
            // shared-memory bins of a CNT_NN accumulator count the NULL rows (rare) instead of the non-NULL ones;
            // the flush converts: non-null = rows - nulls.  The global work table always holds non-null counts.
            const bool count_nulls = acc.kind == ACC_CNT_NN && kStrategy != HDK_B200_STRATEGY_GLOBAL && kStrategy != HDK_B200_STRATEGY_BASELINE;
            if (acc_arg_is_null(p, acc, vals) != count_nulls) continue;
            accumulate_one<kStrategy>(args, bins, tid, a, acc, idx, acc_input(p, acc, vals));
          }
        }
      }
    }
    // next combination: advance the innermost level that was probed and has matches left
    int l = (dead ? lvl : J) - 1;
    while (l >= 0 && ++cur[l] >= n_matches[l]) --l;
    if (l < 0) break;
    lvl = l + 1;
  }
}

// ---- one row, compile-time plan structure ------------------------------------------------------
// Split in two phases so that a thread can evaluate several rows (independent load / probe chains in
// flight together) before it touches the accumulators.  `raw` holds the row's column elements already
// fetched from the staged tile.  Returns false when the row is dropped (join miss / filter / error).
template <class Shape>
__device__ __forceinline__ bool eval_row_static(const ScanArgs& args, const uint64_t* raw, V* vals, uint32_t& idx, uint32_t& max_idx,
                                                int32_t& my_err) {
  constexpr DPlan sp = Shape::get();
  const DPlan& rp = args.plan;  // literals, key ranges, entry count
  int64_t rowid[HDK_B200_MAX_JOINS];
  int32_t row_err = 0, qual_err = 0;
  bool alive = true;
  auto load_outer = [&](int c, int) -> uint64_t { return raw[c]; };
  auto load_inner = [&](int j, int c, int w) -> uint64_t {
    return alive ? ldg_elem(reinterpret_cast<const uint8_t*>(args.inner_col_buffers[j * HDK_B200_MAX_COLS + c]) + size_t(rowid[j]) * w, w) : 0;
  };
  static_for<0, sp.n_exprs>([&](auto I) {
    constexpr int n = decltype(I)::value;
    constexpr DPlan sp = Shape::get();  // (a captured constexpr object is not a constant expression inside the lambda)
    DExpr e = sp.exprs[n];
    if constexpr (sp.exprs[n].op == HDK_B200_OP_CONST) e.imm = rp.exprs[n].imm;  // literals are run-time; widths / units are structure
    int32_t err = 0;
    vals[n] = eval_node(sp, e, vals, err, load_outer, load_inner);
    if constexpr ((sp.exprs[n].aux & kAuxInQual) != 0) { if (err && !qual_err) qual_err = err; }
    else { if (err && !row_err) row_err = err; }
    static_for<0, sp.n_joins>([&](auto J) {
      constexpr int j = decltype(J)::value;
      constexpr DPlan sp = Shape::get();
      if constexpr (sp.joins[j].key_expr == n && sp.joins[j].n_key_exprs > 0) {
        int64_t k64[HDK_B200_MAX_KEYS];
        static_for<0, sp.joins[j].n_key_exprs>([&](auto Kc) {
          constexpr DPlan sp = Shape::get();
          k64[decltype(Kc)::value] = vals[sp.joins[j].key_exprs[decltype(Kc)::value]].i;
        });
        int64_t ridx = -1;
        if (alive) {
          const int8_t* tbl = reinterpret_cast<const int8_t*>(args.join_hash_tables[j]);
          ridx = rp.joins[j].key_width == 4 ? baseline_join_probe<int32_t>(tbl, rp.join_entry_count[j], k64, sp.joins[j].n_key_exprs)
                                            : baseline_join_probe<int64_t>(tbl, rp.join_entry_count[j], k64, sp.joins[j].n_key_exprs);
        }
        alive = alive && ridx >= 0;
        rowid[j] = ridx;
      } else if constexpr (sp.joins[j].key_expr == n) {
        const DJoin& jn = rp.joins[j];
        const int64_t key = vals[n].i;
        bool hit = alive && !((sp.joins[j].key_nullable && key == jn.null_val) || key < jn.min_key || key > jn.max_key);
        int64_t ridx = -1;
        if (hit) {
          const int64_t slot = key - jn.min_key;
          if (jn.by_slot == 2) {        // slot-ordered inner columns, every slot occupied
            ridx = slot;
          } else if (jn.by_slot == 1) { // presence bitmap + slot-ordered inner columns: the row id is the slot
            const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(args.join_hash_tables[j]) + (slot >> 5));
            ridx = ((word >> (slot & 31)) & 1u) ? slot : -1;
          } else {
            ridx = __ldg(reinterpret_cast<const int32_t*>(args.join_hash_tables[j]) + slot);
          }
        }
        alive = hit && ridx >= 0;
        rowid[j] = ridx;
      }
    });
  });
  if (!alive) return false;
  if (qual_err) { my_err = my_err > 0 ? my_err : qual_err; return false; }   // raised whether or not the row passes
  bool pass = true;
  static_for<0, sp.n_filters>([&](auto F) {
    constexpr DPlan sp = Shape::get();
    pass = pass && (vals[sp.filters[decltype(F)::value]].i > 0);
  });
  if (!pass) return false;
  if (row_err) { my_err = my_err > 0 ? my_err : row_err; return false; }
  if constexpr (sp.hash_type == HDK_B200_PERFECT_HASH) {
    int64_t h = 0;
    static_for<0, sp.n_keys>([&](auto K) {
      constexpr int k = decltype(K)::value;
      constexpr DPlan sp = Shape::get();
      const DKey& ky = rp.keys[k];
      int64_t v = vals[sp.keys[k].expr].i;
      if (sp.keys[k].has_nulls && v == int_null_of(sp.keys[k].width)) v = ky.null_translated;
      int64_t term = v - ky.min_val;
      if (ky.bucket) term /= ky.bucket;     // run-time property of the key range (uniform branch, not taken for plain keys)
      if constexpr (k == 0) h = term; else h += term * ky.mult;
    });
    idx = uint32_t(h);
    max_idx = max(max_idx, idx);   // a key outside the range the layout was built for is reported once per tile (error 1003)
    if (idx >= rp.entry_count) return false;
  } else {
    const int64_t entry = baseline_entry(args, sp.n_keys, [&](int k) { constexpr DPlan sp = Shape::get(); return sp.keys[k].expr; }, vals);
    if (entry == kClaimTimedOut) { my_err = my_err > 0 ? my_err : HDK_B200_ERR_CLAIM_TIMEOUT; return false; }
    if (entry < 0) { if (my_err <= 0) my_err = -HDK_B200_ERR_OUT_OF_SLOTS; return false; }
    idx = uint32_t(entry);
  }
  return true;
}

__host__ __device__ constexpr bool shape_has_wide_acc(const DPlan& p) {
  for (int a = 0; a < p.n_acc; ++a)
    if (p.accs[a].bytes == 8) return true;
  return false;
}

template <class Shape, int kStrategy>
__device__ __forceinline__ void accumulate_row_static(const ScanArgs& args, const V* vals, uint32_t idx, uint8_t* bins, int tid,
                                                      int32_t& my_err) {
  constexpr DPlan sp = Shape::get();
  if constexpr (kStrategy == HDK_B200_STRATEGY_CTA_SHARED && shape_has_wide_acc(sp)) {
    // Counters use the native 32-bit shared atomics.  The 64-bit ones are CAS loops that collapse when lanes of one
    // warp hit the same bin, so lanes holding the same group take turns: round r updates the r-th lane of each group.
    static_for<0, sp.n_acc>([&](auto A) {
      constexpr int a = decltype(A)::value;
      constexpr DPlan sp = Shape::get();
      constexpr DAcc acc = sp.accs[a];
      if constexpr (acc.bytes == 4) {
        constexpr bool count_nulls = acc.kind == ACC_CNT_NN;
        if (acc_arg_is_null(sp, acc, vals) == count_nulls) accumulate_one<kStrategy>(args, bins, tid, a, acc, idx, acc_input(sp, acc, vals));
      }
    });
    // (with many groups two lanes rarely meet and the plain CAS retry is cheaper than finding the peers:
    //  measured on config 1, 1000 groups: 0.103 ms without, 0.134 ms with)
    uint32_t rank = 0, rounds = 0;
    if (args.plan.entry_count < 256) {
      const uint32_t active = __activemask();
      const uint32_t peers = __match_any_sync(active, idx);
      rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
      rounds = __reduce_max_sync(active, rank);
    }
    for (uint32_t r = 0; r <= rounds; ++r) {
      if (rank == r) {
        static_for<0, sp.n_acc>([&](auto A) {
          constexpr int a = decltype(A)::value;
          constexpr DPlan sp = Shape::get();
          constexpr DAcc acc = sp.accs[a];
          if constexpr (acc.bytes == 8) {
            if (!acc_arg_is_null(sp, acc, vals)) accumulate_one<kStrategy>(args, bins, tid, a, acc, idx, acc_input(sp, acc, vals));
          }
        });
      }
    }
  } else {
    static_for<0, sp.n_acc>([&](auto A) {
      constexpr int a = decltype(A)::value;
      constexpr DPlan sp = Shape::get();
      constexpr DAcc acc = sp.accs[a];
      constexpr bool count_nulls = acc.kind == ACC_CNT_NN && kStrategy != HDK_B200_STRATEGY_GLOBAL && kStrategy != HDK_B200_STRATEGY_BASELINE;
      if (acc_arg_is_null(sp, acc, vals) == count_nulls) accumulate_one<kStrategy>(args, bins, tid, a, acc, idx, acc_input(sp, acc, vals));
    });
  }
}

// ---- REGISTER strategy: few groups, many aggregates (TPC-H Q1) ------------------------------------
// Every consumer thread keeps the accumulators of all G (<= 8) groups in registers: a row updates group g's set
// under the predicate idx == g.  No shared-memory traffic per row; the price is G predicated updates per
// accumulator, which the integer / fp64 pipes absorb while the kernel waits for HBM.  COUNT(arg) counts the (rare)
// NULL rows in a small per-CTA table instead of spending G registers: non-null = rows - nulls.
constexpr int kRegGroups = 8;
constexpr int kRegThreads = 384 + 32;   // most consumer threads + producer warp of a REGISTER-strategy CTA

template <class Shape, int G>
struct RegAcc {
  static constexpr int NA = Shape::get().n_acc > 0 ? Shape::get().n_acc : 1;
  int64_t wide[G][NA];     // SUM / MIN / MAX cells (fp64 as bits); entries of other accumulators are dead and cost nothing
  uint32_t cnt[G];         // rows of the group (accumulator 0 is always CNT_ALL, see lower.cu)
};

template <class Shape, int G>
__device__ __forceinline__ void reg_init(RegAcc<Shape, G>& r) {
  constexpr DPlan sp = Shape::get();
#pragma unroll
  for (int g = 0; g < G; ++g) {
    r.cnt[g] = 0;
    static_for<0, sp.n_acc>([&](auto A) {
      constexpr int a = decltype(A)::value;
      constexpr DPlan sp = Shape::get();
      r.wide[g][a] = acc_identity(sp.accs[a].kind);
    });
  }
}

template <class Shape, int G>
__device__ __forceinline__ void reg_accumulate(RegAcc<Shape, G>& r, uint32_t* null_bins, bool ok, const V* vals, uint32_t idx) {
  constexpr DPlan sp = Shape::get();
  // inputs masked to the accumulator's identity when the row is dropped or the argument is NULL, so that the
  // per-group update needs the single predicate idx == g
  int64_t x[RegAcc<Shape, G>::NA];
  static_for<0, sp.n_acc>([&](auto A) {
    constexpr int a = decltype(A)::value;
    constexpr DPlan sp = Shape::get();
    constexpr DAcc acc = sp.accs[a];
    const bool is_null = acc_arg_is_null(sp, acc, vals);
    if constexpr (acc.kind == ACC_CNT_NN) {
      if (ok && is_null) atomicAdd(&null_bins[idx * sp.n_acc + a], 1u);
    } else if constexpr (acc.kind == ACC_SUM_F) {
      x[a] = (ok && !is_null) ? acc_input(sp, acc, vals) : int64_t(0x8000000000000000ULL);   // -0.0: s + (-0.0) == s for every s
    } else if constexpr (acc.kind != ACC_CNT_ALL) {
      x[a] = (ok && !is_null) ? acc_input(sp, acc, vals) : acc_identity(acc.kind);
    }
  });
  if (!ok) idx = 0xffffffffu;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (idx == uint32_t(g)) {
      r.cnt[g] += 1u;
      static_for<0, sp.n_acc>([&](auto A) {
        constexpr int a = decltype(A)::value;
        constexpr DPlan sp = Shape::get();
        constexpr DAcc acc = sp.accs[a];
        if constexpr (acc.kind == ACC_SUM_I) r.wide[g][a] += x[a];
        else if constexpr (acc.kind == ACC_SUM_F) r.wide[g][a] = __double_as_longlong(__longlong_as_double(r.wide[g][a]) + __longlong_as_double(x[a]));
        else if constexpr (acc.kind == ACC_MIN_I || acc.kind == ACC_MIN_F) r.wide[g][a] = min(r.wide[g][a], x[a]);
        else if constexpr (acc.kind == ACC_MAX_I || acc.kind == ACC_MAX_F) r.wide[g][a] = max(r.wide[g][a], x[a]);
      });
    }
  }
}

// warp-reduce every (group, accumulator) pair and merge lane 0's result into the global work table
template <class Shape, int G>
__device__ __forceinline__ void reg_flush(const ScanArgs& args, RegAcc<Shape, G>& r, const uint32_t* null_bins, int warp, int lane) {
  constexpr DPlan sp = Shape::get();
  const uint32_t E = args.plan.entry_count;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (uint32_t(g) >= E) break;
    int64_t rows = int64_t(r.cnt[g]);   // (a thread sees < 2^32 rows per launch; the warp total is summed in 64 bits)
    for (int d = 16; d; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
    static_for<0, sp.n_acc>([&](auto A) {
      constexpr int a = decltype(A)::value;
      constexpr DPlan sp = Shape::get();
      constexpr DAcc acc = sp.accs[a];
      int64_t x;
      if constexpr (acc.kind == ACC_CNT_ALL) {
        x = rows;
      } else if constexpr (acc.kind == ACC_CNT_NN) {
        x = rows - (warp == 0 ? int64_t(null_bins[g * sp.n_acc + a]) : 0);   // the CTA's NULL rows are subtracted once
      } else if constexpr (acc.kind == ACC_SUM_I) {
        x = r.wide[g][a];
        for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      } else if constexpr (acc.kind == ACC_SUM_F) {
        double s = __longlong_as_double(r.wide[g][a]);
        for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        x = __double_as_longlong(s);
      } else if constexpr (acc.kind == ACC_MIN_I || acc.kind == ACC_MIN_F) {
        x = r.wide[g][a];
        for (int d = 16; d; d >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, d));
      } else {
        x = r.wide[g][a];
        for (int d = 16; d; d >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, d));
      }
      if (lane == 0 && x != acc_identity(acc.kind)) cell_update_global(acc.kind, args.work_table + size_t(a) * E + g, x);
    });
  }
}

template <int kStrategy, class Shape, int G = kRegGroups>
__global__ void __launch_bounds__(kStrategy == HDK_B200_STRATEGY_REGISTER ? kRegThreads : kThreads, kStrategy == HDK_B200_STRATEGY_REGISTER ? 1 : 2)
scan_kernel(const __grid_constant__ ScanArgs args) {
  extern __shared__ __align__(128) uint8_t smem[];
  if (args.run_if && *args.run_if == 0) return;
  const DPlan& p = args.plan;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nct = int(args.consumer_threads);   // consumer threads (multiple of 32); the warp after them is the producer
  const int ncw = nct >> 5;
  const bool is_producer = warp == ncw;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kStages;
  StageHeader* hdr = reinterpret_cast<StageHeader*>(smem + 128);
  uint32_t* tile_prefix = reinterpret_cast<uint32_t*>(smem + args.off_tile_prefix);  // [nfrag + 1]
  uint8_t* bins = smem + args.off_bins;

  // ---- prologue: barriers, tile prefix, bins
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {  // (all kStages barriers are initialised; args.n_stages of them are used)
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], ncw);
    }
    mbar_fence_init();
  }
  if (warp == 0) {
    // inclusive scan of per-fragment tile counts, 32 fragments per step
    uint32_t carry = 0;
    if (lane == 0) tile_prefix[0] = 0;
    for (uint32_t f0 = 0; f0 < args.num_fragments; f0 += 32) {
      const uint32_t f = f0 + lane;
      uint32_t t = 0;
      if (f < args.num_fragments) {
        const int64_t rows = args.num_rows[f];
        t = rows > 0 ? uint32_t((rows + args.tile_rows - 1) / args.tile_rows) : 0;
      }
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, t, d);
        if (lane >= d) t += o;
      }
      if (f < args.num_fragments) tile_prefix[f + 1] = carry + t;
      carry += __shfl_sync(0xffffffffu, t, 31);
    }
  }
  if (kStrategy == HDK_B200_STRATEGY_REGISTER && !is_producer) {
    for (uint32_t i = tid; i < uint32_t(G) * kMaxAcc; i += nct) reinterpret_cast<uint32_t*>(bins)[i] = 0;
  }
  if (!is_producer && (kStrategy == HDK_B200_STRATEGY_THREAD_PRIVATE || kStrategy == HDK_B200_STRATEGY_CTA_SHARED)) {
    // initialise bins to the accumulators' identities
    for (int a = 0; a < p.n_acc; ++a) {
      const DAcc acc = p.accs[a];
      uint8_t* base = bins + args.acc_bin_off[a];
      const uint32_t n = kStrategy == HDK_B200_STRATEGY_THREAD_PRIVATE ? p.entry_count * nct : p.entry_count;
      if (acc.bytes == 4) {
        for (uint32_t i = tid; i < n; i += nct) reinterpret_cast<uint32_t*>(base)[i] = 0;
      } else {
        const int64_t id = acc_identity(acc.kind);
        for (uint32_t i = tid; i < n; i += nct) reinterpret_cast<int64_t*>(base)[i] = id;
      }
    }
  }
  __syncthreads();
  const uint32_t total_tiles = tile_prefix[args.num_fragments];

  if (is_producer) {
    // =============================== producer warp ===============================
    // Lane c owns column c: it keeps the current fragment's chunk pointer in a register (re-read only when
    // the tile walk enters a new fragment), patches its own head / tail bytes and issues its own bulk copy,
    // so a tile costs one short pass without dependent global loads.
    const uint64_t policy = policy_evict_first();
    const int c = lane;
    const bool has_col = c < p.n_cols;
    const uint32_t w = has_col ? p.col_width[c] : 0;
    const uint32_t region_off = has_col ? args.col_region_off[c] : 0;
    uint32_t stage = 0, phase = 0, frag = 0, cur_frag = 0xffffffffu;   // phase: parity of the ring round being filled
    bool first_round = true;
    const uint8_t* col_base = nullptr;
    uint64_t frag_rows = 0;
    for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      if (!first_round) mbar_wait(&empty_bar[stage], phase ^ 1);
      while (tile_prefix[frag + 1] <= t) ++frag;  // tiles are visited in increasing order
      if (frag != cur_frag) {
        cur_frag = frag;
        frag_rows = uint64_t(args.num_rows[frag]);
        if (has_col) col_base = reinterpret_cast<const uint8_t*>(args.col_buffers[size_t(frag) * p.n_cols + c]);
      }
      const uint64_t row0 = uint64_t(t - tile_prefix[frag]) * args.tile_rows;
      const uint32_t rows = uint32_t(min(uint64_t(args.tile_rows), frag_rows - row0));
      uint8_t* stage_base = smem + args.off_stages + size_t(stage) * args.stage_bytes;
      uint32_t mid = 0, head = 0, m = 0;
      const uint8_t* src = nullptr;
      if (has_col) {
        src = col_base + row0 * w;
        const uint32_t bytes = rows * w;
        m = uint32_t(reinterpret_cast<uintptr_t>(src) & 15u);
        head = m ? min(16u - m, bytes) : 0u;
        mid = (bytes - head) & ~15u;
        const uint32_t tail = bytes - head - mid;
        uint8_t* dst = stage_base + region_off + m;
        for (uint32_t i = 0; i < head; ++i) dst[i] = src[i];                          // generic-proxy byte patches
        for (uint32_t i = 0; i < tail; ++i) dst[head + mid + i] = src[head + mid + i];
        hdr[stage].col_off[c] = uint32_t(dst - smem);
      }
      uint32_t tx_bytes = mid;
      for (int d = 16; d; d >>= 1) tx_bytes += __shfl_xor_sync(0xffffffffu, tx_bytes, d);
      const bool aligned = __ballot_sync(0xffffffffu, m != 0) == 0;   // every column slice starts on a 16-byte boundary
      if (lane == 0) {
        hdr[stage].rows = rows;
        hdr[stage].aligned = aligned ? 1u : 0u;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
      __syncwarp();
      if (mid) bulk_g2s(stage_base + region_off + m + head, src + head, mid, &full_bar[stage], policy);
      if (++stage == args.n_stages) { stage = 0; phase ^= 1; first_round = false; }
    }
  } else {
    // =============================== consumer warps ===============================
    int32_t my_err = 0;
    uint32_t stage = 0, phase = 0;
    RegAcc<Shape, G> racc;   // REGISTER strategy only (dead otherwise)
    uint32_t* null_bins = reinterpret_cast<uint32_t*>(bins);   // REGISTER: [G][n_acc] NULL-row counters of COUNT(arg)
    if constexpr (kStrategy == HDK_B200_STRATEGY_REGISTER && Shape::is_static) reg_init<Shape, G>(racc);
    for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&full_bar[stage], phase);
      const uint32_t rows = hdr[stage].rows;
      if constexpr (Shape::is_static) {
        constexpr DPlan sp = Shape::get();
        constexpr int VW = shape_vec_rows(sp);            // consecutive rows a lane takes with one <= 16-byte load per column
        constexpr int U = shape_iter_rows<Shape>() / VW;  // vector groups per thread per iteration
        constexpr int NC = sp.n_cols > 0 ? sp.n_cols : 1;
        constexpr int NE = sp.n_exprs > 0 ? sp.n_exprs : 1;
        uint32_t max_idx = 0;
        if (rows == args.tile_rows && hdr[stage].aligned && args.full_iters) {
          // full tile (all but a fragment's last, 16-byte aligned chunks): every thread runs the same number of
          // iterations; per iteration U vector loads per column, then U*VW rows evaluated, then accumulated
          const uint8_t* cptr[NC];
          static_for<0, sp.n_cols>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value;
            constexpr DPlan sp = Shape::get();
            cptr[c] = smem + args.off_stages + stage * args.stage_bytes + args.col_region_off[c] + uint32_t(tid) * (VW * sp.col_width[c]);
          });
          for (uint32_t i = 0; i < args.full_iters; ++i) {
            uint32_t words[U][NC][4];
#pragma unroll
            for (int u = 0; u < U; ++u)
              static_for<0, sp.n_cols>([&](auto Cc) {
                constexpr int c = decltype(Cc)::value;
                constexpr DPlan sp = Shape::get();
                constexpr int gw = VW * sp.col_width[c];
                lds_vec<gw>(cptr[c] + size_t(i * U + u) * nct * gw, words[u][c]);
              });
            V vals[U * VW][NE];
            uint32_t idx[U * VW];
            bool ok[U * VW];
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
              for (int v = 0; v < VW; ++v) {
                uint64_t raw[NC];
                static_for<0, sp.n_cols>([&](auto Cc) {
                  constexpr int c = decltype(Cc)::value;
                  constexpr DPlan sp = Shape::get();
                  raw[c] = vec_elem<sp.col_width[c]>(words[u][c], v);
                });
                idx[u * VW + v] = 0;
                ok[u * VW + v] = eval_row_static<Shape>(args, raw, vals[u * VW + v], idx[u * VW + v], max_idx, my_err);
              }
            }
#pragma unroll
            for (int r = 0; r < U * VW; ++r) {
              if constexpr (kStrategy == HDK_B200_STRATEGY_REGISTER) reg_accumulate<Shape, G>(racc, null_bins, ok[r], vals[r], idx[r]);
              else if (ok[r]) accumulate_row_static<Shape, kStrategy>(args, vals[r], idx[r], bins, tid, my_err);
            }
          }
        } else {
          const uint8_t* cbase[NC];
          static_for<0, sp.n_cols>([&](auto Cc) {
            constexpr int c = decltype(Cc)::value;
            cbase[c] = smem + hdr[stage].col_off[c];
          });
          for (uint32_t r = tid; r < rows; r += nct) {
            uint64_t raw[NC];
            static_for<0, sp.n_cols>([&](auto Cc) {
              constexpr int c = decltype(Cc)::value;
              constexpr DPlan sp = Shape::get();
              constexpr int w = sp.col_width[c];
              raw[c] = lds_elem(cbase[c] + size_t(r) * w, w);
            });
            V vals[NE];
            uint32_t idx = 0;
            const bool ok = eval_row_static<Shape>(args, raw, vals, idx, max_idx, my_err);
            if constexpr (kStrategy == HDK_B200_STRATEGY_REGISTER) reg_accumulate<Shape, G>(racc, null_bins, ok, vals, idx);
            else if (ok) accumulate_row_static<Shape, kStrategy>(args, vals, idx, bins, tid, my_err);
          }
        }
        if (sp.hash_type == HDK_B200_PERFECT_HASH && max_idx >= p.entry_count && my_err <= 0) my_err = 1003;
      } else {
        V vals[HDK_B200_MAX_EXPRS];
        const uint32_t* col_off = hdr[stage].col_off;
        for (uint32_t r = tid; r < rows; r += nct) process_row_generic<kStrategy>(args, smem, col_off, r, bins, tid, vals, my_err);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == args.n_stages) { stage = 0; phase ^= 1; }
    }
    if (my_err) record_error(args.error_codes, my_err);

    // ---- flush block partials into the global work table
    if constexpr (kStrategy == HDK_B200_STRATEGY_REGISTER) {
      named_bar_sync(1, nct);   // every consumer's NULL counts are in
      if constexpr (Shape::is_static) reg_flush<Shape, G>(args, racc, null_bins, warp, lane);
    }
    if (kStrategy == HDK_B200_STRATEGY_THREAD_PRIVATE || kStrategy == HDK_B200_STRATEGY_CTA_SHARED) {
      named_bar_sync(1, nct);
      if (kStrategy == HDK_B200_STRATEGY_THREAD_PRIVATE) {
        // one (acc, group) pair per warp step: lanes stride the private copies
        const uint32_t pairs = uint32_t(p.n_acc) * p.entry_count;
        for (uint32_t pr = warp; pr < pairs; pr += ncw) {
          const uint32_t a = pr / p.entry_count, g = pr % p.entry_count;
          const DAcc acc = p.accs[a];
          const uint8_t* base = bins + args.acc_bin_off[a] + size_t(g) * nct * acc.bytes;
          int64_t x;
          if (acc.bytes == 4) {
            uint64_t s = 0;
            for (int i = lane; i < nct; i += 32) s += reinterpret_cast<const uint32_t*>(base)[i];
            if (acc.kind == ACC_CNT_NN) {  // bins hold NULL counts: non-null = rows (accumulator 0) - nulls
              const uint8_t* rows_base = bins + args.acc_bin_off[0] + size_t(g) * nct * 4;
              uint64_t rws = 0;
              for (int i = lane; i < nct; i += 32) rws += reinterpret_cast<const uint32_t*>(rows_base)[i];
              s = rws - s;
            }
            for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            x = int64_t(s);
          } else if (acc.kind == ACC_SUM_I) {
            int64_t s = 0;
            for (int i = lane; i < nct; i += 32) s += reinterpret_cast<const int64_t*>(base)[i];
            for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            x = s;
          } else if (acc.kind == ACC_SUM_F) {
            double s = 0.0;
            for (int i = lane; i < nct; i += 32) s += reinterpret_cast<const double*>(base)[i];
            for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            x = __double_as_longlong(s);
          } else if (acc.kind == ACC_MIN_I || acc.kind == ACC_MIN_F) {
            int64_t s = INT64_MAX;
            for (int i = lane; i < nct; i += 32) s = min(s, reinterpret_cast<const int64_t*>(base)[i]);
            for (int d = 16; d; d >>= 1) s = min(s, __shfl_xor_sync(0xffffffffu, s, d));
            x = s;
          } else {
            int64_t s = INT64_MIN;
            for (int i = lane; i < nct; i += 32) s = max(s, reinterpret_cast<const int64_t*>(base)[i]);
            for (int d = 16; d; d >>= 1) s = max(s, __shfl_xor_sync(0xffffffffu, s, d));
            x = s;
          }
          if (lane == 0) {
            if (x != acc_identity(acc.kind))
              cell_update_global(acc.kind, args.work_table + size_t(a) * p.entry_count + g, x);
          }
        }
      } else {
        for (int a = 0; a < p.n_acc; ++a) {
          const DAcc acc = p.accs[a];
          const uint8_t* base = bins + args.acc_bin_off[a];
          for (uint32_t g = tid; g < p.entry_count; g += nct) {
            int64_t x = acc.bytes == 4 ? int64_t(reinterpret_cast<const uint32_t*>(base)[g]) : reinterpret_cast<const int64_t*>(base)[g];
            if (acc.kind == ACC_CNT_NN) x = int64_t(reinterpret_cast<const uint32_t*>(bins + args.acc_bin_off[0])[g]) - x;
            if (x != acc_identity(acc.kind)) cell_update_global(acc.kind, args.work_table + size_t(a) * p.entry_count + g, x);
          }
        }
      }
    }
  }

  // ---- multi-GPU exchange: the last CTA to finish publishes this GPU's partial table to every peer ----
  if (args.n_peers) {
    __shared__ bool is_last;
    __threadfence();     // this thread's atomics on the local work table are visible device-wide
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(args.ticket, 1ull) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
      __threadfence();
      const uint64_t n2 = args.n_cells / 2;   // 16-byte pieces (the table is 16-byte aligned)
      for (uint32_t pr = 0; pr < args.n_peers; ++pr) {
        longlong2* dst = reinterpret_cast<longlong2*>(args.peer_slot[pr]);
        const longlong2* src = reinterpret_cast<const longlong2*>(args.work_table);
        for (uint64_t i = tid; i < n2; i += blockDim.x) dst[i] = __ldcg(src + i);
        if ((args.n_cells & 1) && tid == 0) args.peer_slot[pr][args.n_cells - 1] = __ldcg(args.work_table + args.n_cells - 1);
      }
      __threadfence_system();   // the copies are visible on the peers before the flags
      __syncthreads();
      if (uint32_t(tid) < args.n_peers) *reinterpret_cast<volatile unsigned long long*>(args.peer_flag[tid]) = args.epoch;
    }
  }
}

}  // namespace hb
