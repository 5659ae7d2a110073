// hdk_b200/csrc/scan.cu — the fused scan → filter → join probe → group-by → aggregate kernel.
//
// Replaces the JIT'd multifrag_query_hoisted_literals / query_group_by_template / row_func of the
// reference (QE/RuntimeFunctions.cpp:1692-1726, QE/QueryTemplateGenerator.cpp:488-772) and the
// runtime it calls per row (QE/GroupByRuntime.cpp, QE/cuda_mapd_rt.cu:423-1083).
//
// Execution model (sm_100a):
//   * persistent grid, a multiple of the SM count; every CTA walks tiles t = blockIdx.x, +gridDim.x
//   * one producer warp streams each tile's column slices global → shared with 1-D TMA bulk copies
//     (cp.async.bulk … mbarrier::complete_tx, SASS UBLKCP) through a kStages-deep mbarrier ring;
//     unaligned heads/tails (< 16 B per column) are patched with byte copies, so any chunk pointer
//     and any row count work on the same path
//   * kConsumerWarps consumer warps evaluate the expression DAG per row out of shared memory and
//     accumulate into NEUTRAL accumulators (SUM→0, MIN→+max, MAX→−max, counters) — the same
//     representation merges across CTAs (atomics) and across GPUs (NCCL SUM/MIN/MAX):
//       THREAD_PRIVATE : per-thread bins in shared memory, [acc][group][thread], no atomics,
//                        bank-conflict free by construction          (few groups)
//       CTA_SHARED     : one table per CTA in shared memory, shared atomics   (up to ~100 KB)
//       GLOBAL         : straight into the global work table with RED atomics
//   * a finalize kernel converts the work table into the reference's buffer encoding
//     (finalize.cu), so the result is byte-compatible with QueryMemoryDescriptor.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "accum.cuh"
#include "baseline.cuh"
#include "device_utils.cuh"
#include "eval.cuh"
#include "partagg.cuh"
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
template <int kStrategy>
__device__ __forceinline__ void process_row_generic(const ScanArgs& args, const uint8_t* smem, const uint32_t* col_off, uint32_t r,
                                                    uint8_t* bins, int tid, V* vals, int32_t& my_err) {
  const DPlan& p = args.plan;
  int64_t rowid[HDK_B200_MAX_JOINS];
  auto load_outer = [&](int c, int w) -> uint64_t { return lds_elem(smem + col_off[c] + size_t(r) * w, w); };
  auto load_inner = [&](int j, int c, int w) -> uint64_t {
    return ldg_elem(reinterpret_cast<const uint8_t*>(args.inner_col_buffers[j * HDK_B200_MAX_COLS + c]) + size_t(rowid[j]) * w, w);
  };
  int32_t row_err = 0, qual_err = 0;
  bool dropped = false;
  // 1:N join: nodes up to the key are evaluated once, the rest once per match
  int n_matches = 1;
  const int32_t* match_ids = nullptr;
  int split = p.n_exprs;
  if (p.n_joins == 1 && p.joins[0].one_to_many) split = p.joins[0].key_expr + 1;
  for (int n = 0; n < split && !dropped; ++n) {
    int32_t e = 0;
    vals[n] = eval_node(p, p.exprs[n], vals, e, load_outer, load_inner);
    if (e) { int32_t& dst = (p.exprs[n].aux & kAuxInQual) ? qual_err : row_err; if (!dst) dst = e; }
    for (int j = 0; j < p.n_joins; ++j) {
      const DJoin& jn = p.joins[j];
      if (jn.key_expr != n) continue;
      if (jn.n_key_exprs) {   // composite / wide-range key: baseline join table
        int64_t k64[HDK_B200_MAX_KEYS];
        for (int i = 0; i < jn.n_key_exprs; ++i) k64[i] = vals[jn.key_exprs[i]].i;
        const int8_t* tbl = reinterpret_cast<const int8_t*>(args.join_hash_tables[j]);
        if (jn.one_to_many) {
          // composite-key dictionary, then offsets | counts | payload indexed by the key's position in it
          // (BaselineJoinHashTable one-to-many layout; HashJoin::codegenMatchingSet)
          const int64_t E = p.join_entry_count[j];
          const int64_t slot = jn.key_width == 4 ? baseline_dict_index<int32_t>(tbl, E, k64, jn.n_key_exprs)
                                                 : baseline_dict_index<int64_t>(tbl, E, k64, jn.n_key_exprs);
          if (slot < 0) { dropped = true; break; }
          const int32_t* otm = reinterpret_cast<const int32_t*>(tbl + size_t(E) * size_t(jn.n_key_exprs) * size_t(jn.key_width));
          const int32_t off = __ldg(otm + slot);
          if (off < 0) { dropped = true; break; }
          n_matches = __ldg(otm + E + slot);
          match_ids = otm + 2 * E + off;
          continue;
        }
        const int64_t rid = jn.key_width == 4 ? baseline_join_probe<int32_t>(tbl, p.join_entry_count[j], k64, jn.n_key_exprs)
                                              : baseline_join_probe<int64_t>(tbl, p.join_entry_count[j], k64, jn.n_key_exprs);
        if (rid < 0) { dropped = true; break; }
        rowid[j] = rid;
        continue;
      }
      // hash_join_idx[_nullable] (QE/GroupByRuntime.cpp:298-329)
      const int64_t key = vals[n].i;
      if ((jn.key_nullable && key == jn.null_val) || key < jn.min_key || key > jn.max_key) { dropped = true; break; }
      const int32_t* table = reinterpret_cast<const int32_t*>(args.join_hash_tables[j]);
      const int64_t slot = key - jn.min_key;
      if (jn.one_to_many) {
        // offsets | counts | payload (JHT/PerfectJoinHashTable.cpp:861-886)
        const int64_t E = p.join_entry_count[j];
        const int32_t off = __ldg(table + slot);
        if (off < 0) { dropped = true; break; }
        n_matches = __ldg(table + E + slot);
        match_ids = table + 2 * E + off;
      } else if (jn.by_slot) {
        // presence bitmap + slot-ordered inner columns (hdk_b200_gather_join_payload_on_device)
        if (jn.by_slot == 1 && !((__ldg(reinterpret_cast<const uint32_t*>(table) + (slot >> 5)) >> (slot & 31)) & 1u)) { dropped = true; break; }
        rowid[j] = slot;   // (by_slot == 2: every slot of [min_key, max_key] is occupied, no bitmap)
      } else {
        const int32_t idx = __ldg(table + slot);
        if (idx < 0) { dropped = true; break; }
        rowid[j] = idx;
      }
    }
  }
  if (dropped) return;
  for (int mi = 0; mi < n_matches; ++mi) {
    if (match_ids) rowid[0] = __ldg(match_ids + mi);
    for (int n = split; n < p.n_exprs; ++n) {
      int32_t e = 0;
      vals[n] = eval_node(p, p.exprs[n], vals, e, load_outer, load_inner);
      if (e) { int32_t& dst = (p.exprs[n].aux & kAuxInQual) ? qual_err : row_err; if (!dst) dst = e; }
    }
    if (qual_err) { my_err = my_err > 0 ? my_err : qual_err; continue; }   // an error inside a qual: raised whether or not the row passes
    bool pass = true;
    for (int f = 0; f < p.n_filters; ++f) pass = pass && (vals[p.filters[f]].i > 0);
    if (!pass) continue;
    if (row_err) { my_err = my_err > 0 ? my_err : row_err; continue; }
    uint32_t idx;
    if (kStrategy == HDK_B200_STRATEGY_BASELINE) {
      const int64_t entry = baseline_entry(args, p.n_keys, [&](int k) { return p.keys[k].expr; }, vals);
      if (entry < 0) { if (my_err <= 0) my_err = -HDK_B200_ERR_OUT_OF_SLOTS; continue; }
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
      if (idx >= p.entry_count) { my_err = my_err > 0 ? my_err : 1003; continue; }  // key outside the range the layout was built for
    }
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

typedef void (*ScanKernelFn)(const ScanArgs);
struct StaticEntry {
  uint64_t sig;
  const char* name;
  int iter_rows;       // rows per consumer thread per iteration of the full-tile loop
  ScanKernelFn fn[5];  // THREAD_PRIVATE, CTA_SHARED, GLOBAL, BASELINE, REGISTER (8 groups)
  ScanKernelFn reg_fn[4];  // REGISTER kernels for <= 2, 4, 6, 8 groups
};
// perfect-hash shapes get the three accumulation strategies, baseline-hash shapes the in-place one
// REGISTER only where it can win: several wide accumulators (it replaces n_acc shared-memory updates per row by
// kRegGroups predicated register updates per accumulator) and a register budget that fits
__host__ __device__ constexpr bool shape_wants_registers(const DPlan& p) {
  int wide = 0;
  for (int a = 0; a < p.n_acc; ++a) wide += p.accs[a].bytes == 8;
  return p.hash_type == HDK_B200_PERFECT_HASH && p.n_joins == 0 && wide >= 3 && wide * 2 + (p.n_acc - wide) <= 18;
}
template <int ID, int S, int G = kRegGroups>
constexpr ScanKernelFn pick_kernel() {
  if constexpr (S == HDK_B200_STRATEGY_REGISTER) {
    if constexpr (shape_wants_registers(StaticShape<ID>::get())) return scan_kernel<S, StaticShape<ID>, G>;
    else return nullptr;
  } else if constexpr ((StaticShape<ID>::get().hash_type == HDK_B200_BASELINE_HASH) == (S == HDK_B200_STRATEGY_BASELINE)) {
    return scan_kernel<S, StaticShape<ID>>;
  } else {
    return nullptr;
  }
}
#define HB_STATIC_SHAPE(ID, SIG, NAME, RPI, ...)                                                              \
  {SIG, NAME, shape_iter_rows<StaticShape<ID>>(), {pick_kernel<ID, HDK_B200_STRATEGY_THREAD_PRIVATE>(), pick_kernel<ID, HDK_B200_STRATEGY_CTA_SHARED>(), \
               pick_kernel<ID, HDK_B200_STRATEGY_GLOBAL>(), pick_kernel<ID, HDK_B200_STRATEGY_BASELINE>(), \
               pick_kernel<ID, HDK_B200_STRATEGY_REGISTER>()},                                                    \
   {pick_kernel<ID, HDK_B200_STRATEGY_REGISTER, 2>(), pick_kernel<ID, HDK_B200_STRATEGY_REGISTER, 4>(),                 \
    pick_kernel<ID, HDK_B200_STRATEGY_REGISTER, 6>(), pick_kernel<ID, HDK_B200_STRATEGY_REGISTER, 8>()}},
static const StaticEntry kStaticShapes[] = {
#include "static_shapes.inc"
    {0, nullptr, 0, {nullptr, nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}}};
#undef HB_STATIC_SHAPE

// ---------------------------------------------------------------------------------------------
// work table initialisation
// ---------------------------------------------------------------------------------------------
// accumulator-major [n_acc][E] for perfect hash (merge classes are contiguous for the all-reduce),
// entry-major [E][n_acc] for baseline hash (one entry's cells share a sector)
__global__ void init_work_table_kernel(int64_t* w, uint64_t E, int n_acc, bool entry_major, const __grid_constant__ AccKinds kinds, const int* run_if) {
  if (run_if && *run_if == 0) return;
  const uint64_t n = uint64_t(n_acc) * E;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
    w[i] = acc_identity(kinds.kind[entry_major ? i % uint64_t(n_acc) : i / E]);
}

int init_work_table(const Lowered& lw, int64_t* work_table, cudaStream_t stream, const int* run_if) {
  AccKinds kinds{};
  for (int a = 0; a < lw.plan.n_acc; ++a) kinds.kind[a] = lw.plan.accs[a].kind;
  const uint64_t n = uint64_t(lw.plan.n_acc) * lw.plan.entry_count;
  const int block = 256;
  const int grid = int(std::min<uint64_t>((n + block - 1) / block, uint64_t(sm_count()) * 8));
  init_work_table_kernel<<<std::max(grid, 1), block, 0, stream>>>(work_table, lw.plan.entry_count, lw.plan.n_acc,
                                                                  lw.plan.hash_type == HDK_B200_BASELINE_HASH, kinds, run_if);
  HB_LAUNCH_CHECK();
  return HDK_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// host: geometry + launch
// ---------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// launch geometry per (plan shape, entry count, fragment count, input size class, device, forced choices): per-device
// cache, nothing query-specific in it (literals, pointers and key ranges travel in the arguments of every launch)
struct GeoKey {
  uint64_t sig, entries;
  uint32_t n_frag, hint_bucket, dev, baseline, force_generic, force_strategy, grid_x;
  bool operator==(const GeoKey& o) const {
    return sig == o.sig && entries == o.entries && n_frag == o.n_frag && hint_bucket == o.hint_bucket && dev == o.dev && baseline == o.baseline &&
           force_generic == o.force_generic && force_strategy == o.force_strategy && grid_x == o.grid_x;
  }
};
struct GeoEntry {
  GeoKey key;
  ScanArgs a;            // only the geometry fields are used
  ScanKernelFn kern;
  int grid, block, variant, strategy;
  size_t smem_bytes;
};
static std::mutex g_geo_mutex;
static std::vector<GeoEntry> g_geo_cache;
static const char* geo_env() {   // tuning hook (tools/sweep_geo.py), read once
  static const char* env = getenv("HDK_B200_GEO");
  return env;
}

static int launch_scan_impl(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                            int64_t* work_table, bool baseline, cudaStream_t stream, hdk_b200_launch_info* info,
                            const ExchangeTargets* xchg = nullptr, const int* run_if = nullptr) {
  const DPlan& p = lw.plan;
  if (params->num_fragments > kMaxFragments) { set_error("more than %d fragments per launch", kMaxFragments); return HDK_B200_E_UNSUPPORTED; }
  ScanArgs a{};
  a.plan = p;
  a.col_buffers = params->col_buffers;
  a.num_rows = params->num_rows;
  a.num_fragments = uint32_t(params->num_fragments);
  a.join_hash_tables = params->join_hash_tables;
  a.inner_col_buffers = params->inner_col_buffers;
  a.work_table = work_table;
  a.error_codes = params->error_codes;
  a.layout = lw.layout;
  a.groupby_buf = params->groupby_buf;
  a.run_if = run_if;
  if (xchg) {
    a.n_peers = xchg->n_peers;
    a.n_cells = uint64_t(p.n_acc) * p.entry_count;
    a.epoch = xchg->epoch;
    a.ticket = xchg->ticket;
    for (uint32_t i = 0; i < xchg->n_peers; ++i) { a.peer_slot[i] = xchg->peer_slot[i]; a.peer_flag[i] = xchg->peer_flag[i]; }
  }

  int dev = 0;
  HB_CUDA(cudaGetDevice(&dev));
  // ---- launch geometry of an earlier launch of the same shape?  (the search below queries occupancy and sets function
  //      attributes: not something to repeat per launch of a 100 µs query)
  const uint64_t total_hint = params->total_rows_hint;
  int hint_bucket = 0;
  while ((total_hint >> hint_bucket) > 1) ++hint_bucket;
  GeoKey key{plan_signature(p), uint64_t(p.entry_count), a.num_fragments, uint32_t(hint_bucket), uint32_t(dev), baseline ? 1u : 0u,
             uint32_t(g_debug.force_generic), uint32_t(g_debug.force_strategy + 1), ko ? ko->gridDimX : 0u};
  {
    std::lock_guard<std::mutex> lock(g_geo_mutex);
    for (const GeoEntry& ge : g_geo_cache)
      if (ge.key == key && !geo_env()) {
        a.consumer_threads = ge.a.consumer_threads; a.n_stages = ge.a.n_stages; a.tile_rows = ge.a.tile_rows; a.off_bins = ge.a.off_bins;
        a.off_stages = ge.a.off_stages; a.off_tile_prefix = ge.a.off_tile_prefix; a.stage_bytes = ge.a.stage_bytes; a.full_iters = ge.a.full_iters;
        memcpy(a.acc_bin_off, ge.a.acc_bin_off, sizeof(a.acc_bin_off));
        memcpy(a.col_region_off, ge.a.col_region_off, sizeof(a.col_region_off));
        ge.kern<<<ge.grid, ge.block, ge.smem_bytes, stream>>>(a);
        HB_LAUNCH_CHECK();
        if (info) {
          info->variant = ge.variant; info->strategy = ge.strategy; info->grid = ge.grid; info->block = ge.block;
          info->smem_bytes = int(ge.smem_bytes); info->n_accumulators = p.n_acc; info->tile_rows = int(ge.a.tile_rows);
        }
        return HDK_B200_OK;
      }
  }
  int max_smem = 0;
  HB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

  // ---- shared-memory map and launch geometry ---------------------------------------------------
  // fixed part: barriers (128 B) + stage headers + tile prefix
  size_t off = 128 + align_up(sizeof(StageHeader) * kStages, 16);
  a.off_tile_prefix = uint32_t(off);
  off += align_up(size_t(a.num_fragments + 1) * 4, 16);
  const size_t fixed_bytes = align_up(off, 16);
  const size_t E = p.entry_count;
  const size_t row_bytes = std::max<size_t>(lw.stage_row_bytes, 1);
  bool counters_only = true;
  for (int i = 0; i < p.n_acc; ++i) counters_only = counters_only && p.accs[i].bytes == 4;

  // pre-compiled shape for this plan?  (its iteration granularity shapes the tile size)
  const StaticEntry* stat = nullptr;
  if (!g_debug.force_generic) {  // (hdk_b200_debug_set("force_generic", 1): tests run the interpreter on the benchmark shapes)
    const uint64_t sig = plan_signature(p);
    for (int i = 0; kStaticShapes[i].name; ++i)
      if (kStaticShapes[i].sig == sig) { stat = &kStaticShapes[i]; break; }
  }
  const int iter_rows = stat ? stat->iter_rows : 1;

  struct Geo {
    int strategy, nct, ctas, stages;
    uint32_t tile_rows;
    size_t off_stages;
  };
  int sm_smem = 0;
  HB_CUDA(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  // Largest tile for (strategy, consumer threads, CTAs per SM, ring depth); false if nothing fits.
  auto fit = [&](int strategy, int nct, int ctas, int stages, Geo* g) -> bool {
    if (ctas * (nct + 32) > 2048) return false;
    size_t bins = 0;
    for (int i = 0; i < p.n_acc; ++i) {
      bins = align_up(bins, 16);
      if (strategy == HDK_B200_STRATEGY_THREAD_PRIVATE) bins += size_t(p.accs[i].bytes) * E * nct;
      else if (strategy == HDK_B200_STRATEGY_CTA_SHARED) bins += size_t(p.accs[i].bytes) * E;
    }
    if (strategy == HDK_B200_STRATEGY_REGISTER) bins = size_t(kRegGroups) * kMaxAcc * 4;   // NULL-row counters of COUNT(arg)
    const size_t off_stages = align_up(fixed_bytes + bins, 128);
    // each resident CTA costs its dynamic shared memory + 1 KB reserved by the driver
    const size_t budget = std::min<size_t>(size_t(sm_smem) / ctas - 1024, size_t(max_smem) - 1024);
    if (off_stages >= budget) return false;
    const size_t per_stage = std::min<size_t>((budget - off_stages) / stages, 72 * 1024);
    // a whole number of full-tile iterations (nct * iter_rows rows each); every column slice is then a multiple of 16 bytes
    const uint32_t quantum = uint32_t(nct) * uint32_t(iter_rows);
    auto fits = [&](uint32_t rows) { return align_up(size_t(rows) * row_bytes + 48 * size_t(p.n_cols), 128) <= per_stage; };
    // small inputs: keep >= 8 tiles per resident CTA when the caller told us the row count
    uint32_t cap = 32768;
    if (params->total_rows_hint) {
      const uint64_t want = params->total_rows_hint / (uint64_t(sm_count()) * ctas * 8);
      cap = uint32_t(std::min<uint64_t>(cap, std::max<uint64_t>(want, quantum)));
    }
    uint32_t tr = 0;
    for (uint32_t cand = quantum; cand <= cap && fits(cand); cand += quantum) tr = cand;
    if (tr == 0)   // tiles smaller than one iteration quantum: scalar path only
      for (uint32_t cand = 32; cand < quantum && fits(cand); cand *= 2) tr = cand;
    if (tr == 0) return false;
    *g = Geo{strategy, nct, ctas, stages, tr, off_stages};
    return true;
  };
  // Score fitted to tools/sweep_geo.py runs on the taxi shapes (profiles/r1_geometry_sweep.md): ~768 consumer
  // threads per SM is the sweet spot (three 256-thread CTAs, each with its own ring, beat two 512-thread CTAs by
  // 5-15 %), more rows per thread per tile amortise the per-tile barrier work (~2 rows' worth), and a 2-deep ring
  // only pays off when its tiles are large enough to cover the refill latency (~0.7 rows per thread).
  auto best = [&](int strategy, Geo* out) -> bool {
    static const int ncts[] = {256, 384, 192, 512, 128, 64, 32};
    double best_score = -1.0;
    for (int nct : ncts)
      for (int ctas = 1; ctas <= 6; ++ctas)
        for (int stages = 2; stages <= 3; ++stages) {
          Geo g;
          // REGISTER kernels are compiled for <= 256 consumer threads and a large register file share
          if (strategy == HDK_B200_STRATEGY_REGISTER && (nct + 32 > kRegThreads || ctas > 2)) continue;
          if (!fit(strategy, nct, ctas, stages, &g)) continue;
          if (strategy == HDK_B200_STRATEGY_REGISTER) {   // the register file decides how many of these CTAs are resident
            ScanKernelFn k = stat->reg_fn[E <= 2 ? 0 : E <= 4 ? 1 : E <= 6 ? 2 : 3];
            const size_t smem_need = g.off_stages + size_t(stages) * align_up(size_t(g.tile_rows) * row_bytes + 48 * size_t(p.n_cols), 128);
            int occ = 0;
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 1024) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, nct + 32, smem_need) != cudaSuccess || occ < ctas) {
              cudaGetLastError();
              continue;
            }
          }
          const double T = double(nct) * ctas;
          // plans that gather through a join table wait on L2 / DRAM latency per row: they want more resident warps
          // (config 5: 4.8 ms at 768 threads per SM, 2.7 ms at 1152)
          const double t_best = p.n_joins ? 1152.0 : 768.0;
          const double teff = T <= t_best ? T : t_best - 0.5 * (T - t_best);
          const double rpt = double(g.tile_rows) / nct;
          const bool full = g.tile_rows % (uint32_t(nct) * uint32_t(iter_rows)) == 0;
          // (gathering plans also want the L1 the ring would take: config 5 with 3 x 384 threads, 2.7 ms with a 2-deep
          //  ring, 4.2 ms with a 3-deep one)
          const double ring = p.n_joins ? (stages == 2 ? 1.0 : 0.7) : (stages == 2 ? rpt / (rpt + 0.7) : 1.0);
          // (REGISTER kernels: one big CTA per SM measured 10 % faster than two small ones on TPC-H Q1 — fewer flushes)
          const double one_cta = (strategy == HDK_B200_STRATEGY_REGISTER && ctas == 1) ? 1.05 : 1.0;
          const double score = teff * (rpt / (rpt + 2.0)) * (full ? 1.0 : 0.5) * ring * one_cta;
          if (score > best_score) { best_score = score; *out = g; }
        }
    return best_score >= 0.0;
  };
  Geo geo{};
  bool have = false;
  int forced = -1;
  if (!baseline && g_debug.force_strategy >= 0 && g_debug.force_strategy <= HDK_B200_STRATEGY_REGISTER &&
      g_debug.force_strategy != HDK_B200_STRATEGY_BASELINE)
    forced = g_debug.force_strategy;   // hdk_b200_debug_set("force_strategy", s): tests / tuning
  if (baseline) have = best(HDK_B200_STRATEGY_BASELINE, &geo);
  else if (forced == HDK_B200_STRATEGY_REGISTER && !(stat && stat->fn[forced] && E <= size_t(kRegGroups))) have = false;
  else if (forced >= 0) have = best(forced, &geo);
  else {
    Geo g;
    // few groups, several wide aggregates, pre-compiled shape: accumulators live in registers
    if (stat && stat->fn[HDK_B200_STRATEGY_REGISTER] && E <= size_t(kRegGroups) && best(HDK_B200_STRATEGY_REGISTER, &g)) { geo = g; have = true; }
    // counters only and enough groups to spread the native shared atomics: one table per CTA
    if (!have && counters_only && E >= 32 && best(HDK_B200_STRATEGY_CTA_SHARED, &g) && g.ctas * g.nct >= 512) { geo = g; have = true; }
    // private bins: no atomics at all.  With few groups a shared table would serialise on its hot bins, so take
    // private bins even when only a few warps fit.
    if (!have && best(HDK_B200_STRATEGY_THREAD_PRIVATE, &g) && (g.ctas * g.nct >= 256 || E < 64)) { geo = g; have = true; }
    if (!have && best(HDK_B200_STRATEGY_CTA_SHARED, &g)) { geo = g; have = true; }
    if (!have) have = best(HDK_B200_STRATEGY_GLOBAL, &geo);
  }
  // tuning hook (tools/sweep_geo.py): HDK_B200_GEO="strategy,consumer_threads,ctas_per_sm,stages,tile_rows" overrides the search
  if (const char* env = geo_env()) {
    int es = 0, en = 0, ec = 0, est = 0, et = 0;
    if (sscanf(env, "%d,%d,%d,%d,%d", &es, &en, &ec, &est, &et) == 5 && en >= 32 && en <= kConsumerWarps * 32 && en % 32 == 0 &&
        est >= 1 && est <= kStages && et >= 32 && es >= 0 && es <= HDK_B200_STRATEGY_REGISTER &&
        (es == HDK_B200_STRATEGY_BASELINE) == baseline) {
      size_t bins = 0;
      for (int i = 0; i < p.n_acc; ++i) {
        bins = align_up(bins, 16);
        if (es == HDK_B200_STRATEGY_THREAD_PRIVATE) bins += size_t(p.accs[i].bytes) * E * en;
        else if (es == HDK_B200_STRATEGY_CTA_SHARED) bins += size_t(p.accs[i].bytes) * E;
      }
      if (es == HDK_B200_STRATEGY_REGISTER) bins = size_t(kRegGroups) * kMaxAcc * 4;
      geo = Geo{es, en, ec, est, uint32_t(et), align_up(fixed_bytes + bins, 128)};
      have = true;
    }
  }
  if (!have) { set_error(forced >= 0 ? "forced strategy does not fit in shared memory" : "stage ring does not fit in shared memory"); return HDK_B200_E_UNSUPPORTED; }
  const int strategy = geo.strategy;
  a.consumer_threads = uint32_t(geo.nct);
  a.n_stages = uint32_t(geo.stages);
  a.tile_rows = geo.tile_rows;
  a.off_bins = uint32_t(fixed_bytes);
  a.off_stages = uint32_t(geo.off_stages);
  {
    size_t bins = 0;
    for (int i = 0; i < p.n_acc; ++i) {
      bins = align_up(bins, 16);
      a.acc_bin_off[i] = uint32_t(bins);
      if (strategy == HDK_B200_STRATEGY_THREAD_PRIVATE) bins += size_t(p.accs[i].bytes) * E * geo.nct;
      else if (strategy == HDK_B200_STRATEGY_CTA_SHARED) bins += size_t(p.accs[i].bytes) * E;
    }
  }
  size_t so = 0;
  for (int c = 0; c < p.n_cols; ++c) {
    a.col_region_off[c] = uint32_t(so);
    so += align_up(size_t(geo.tile_rows) * p.col_width[c] + 32, 16);  // +16 for the alignment shift, +16 slack
  }
  a.stage_bytes = uint32_t(align_up(so, 128));
  const size_t smem_bytes = a.off_stages + size_t(a.stage_bytes) * geo.stages;
  if (smem_bytes > size_t(max_smem) - 1024) { set_error("stage ring does not fit in shared memory (%zu > %d)", smem_bytes, max_smem - 1024); return HDK_B200_E_UNSUPPORTED; }
  const int block = geo.nct + 32;
  int grid = sm_count() * geo.ctas;
  if (ko && ko->gridDimX) grid = int(ko->gridDimX);

  if (strategy == HDK_B200_STRATEGY_REGISTER && !(stat && stat->fn[strategy])) { set_error("REGISTER strategy needs a pre-compiled shape"); return HDK_B200_E_UNSUPPORTED; }
  ScanKernelFn kern = strategy == HDK_B200_STRATEGY_THREAD_PRIVATE ? scan_kernel<HDK_B200_STRATEGY_THREAD_PRIVATE, GenericShape>
                      : strategy == HDK_B200_STRATEGY_CTA_SHARED   ? scan_kernel<HDK_B200_STRATEGY_CTA_SHARED, GenericShape>
                      : strategy == HDK_B200_STRATEGY_GLOBAL       ? scan_kernel<HDK_B200_STRATEGY_GLOBAL, GenericShape>
                                                                   : scan_kernel<HDK_B200_STRATEGY_BASELINE, GenericShape>;
  int variant = 0;
  if (stat && stat->fn[strategy]) {
    kern = stat->fn[strategy];
    if (strategy == HDK_B200_STRATEGY_REGISTER) kern = stat->reg_fn[E <= 2 ? 0 : E <= 4 ? 1 : E <= 6 ? 2 : 3];
    variant = int(stat - kStaticShapes) + 1;
    const uint32_t per_iter = uint32_t(geo.nct) * uint32_t(stat->iter_rows);
    a.full_iters = (geo.tile_rows % per_iter == 0 && geo.tile_rows % 16 == 0) ? geo.tile_rows / per_iter : 0;
  }
  HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 1024));   // (a cap, never lowered: cached geometries of the same kernel stay launchable)
  if (!(ko && ko->gridDimX)) {   // persistent grid = what is actually resident (registers can allow fewer CTAs than planned)
    int occ = 0;
    HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem_bytes));
    grid = sm_count() * std::max(1, std::min(occ, geo.ctas));
  }
  if (!geo_env()) {
    std::lock_guard<std::mutex> lock(g_geo_mutex);
    if (g_geo_cache.size() >= 256) g_geo_cache.clear();
    GeoEntry ge{};
    ge.key = key; ge.a = a; ge.kern = kern; ge.grid = grid; ge.block = block; ge.smem_bytes = smem_bytes; ge.variant = variant; ge.strategy = strategy;
    g_geo_cache.push_back(ge);
  }
  kern<<<grid, block, smem_bytes, stream>>>(a);
  HB_LAUNCH_CHECK();
  if (info) {
    info->variant = variant;
    info->strategy = strategy;
    info->grid = grid;
    info->block = block;
    info->smem_bytes = int(smem_bytes);
    info->n_accumulators = p.n_acc;
    info->tile_rows = int(geo.tile_rows);
  }
  return HDK_B200_OK;
}

int launch_scan(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                int64_t* work_table, cudaStream_t stream, hdk_b200_launch_info* info) {
  return launch_scan_impl(lw, ko, params, work_table, false, stream, info);
}
int launch_scan_exchange(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                         int64_t* work_table, const ExchangeTargets& x, cudaStream_t stream, hdk_b200_launch_info* info) {
  return launch_scan_impl(lw, ko, params, work_table, false, stream, info, &x);
}
int launch_baseline_scan(const Lowered& lw, const hdk_b200_kernel_options* ko, const hdk_b200_kernel_params* params,
                         int64_t* work_table, cudaStream_t stream, hdk_b200_launch_info* info, const int* run_if) {
  return launch_scan_impl(lw, ko, params, work_table, true, stream, info, nullptr, run_if);
}

}  // namespace hb
