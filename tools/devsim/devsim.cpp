// devsim.cpp -- host replay of the per-thread device algorithm (csrc/drg_core.cuh) for development without a GPU.
// DEVELOPMENT / TEST AID ONLY: nothing in the product links or loads this.
// Build: g++ -O2 -std=c++17 -shared -fPIC -o tools/devsim/libdevsim.so tools/devsim/devsim.cpp
//
// sim_count : count_moves per board + the number of DFS descents (the thread-per-board walk)
// sim_split : the split walk (budgeted walks, records handed to the next round, tree sums, emitting re-walk) run
//             round by round on one CPU thread exactly as the kernels schedule it; returns counts and records so that
//             tests/test_devsim.py can compare them with the oracle, plus per-round statistics.
#include "cuda_shim.h"

#include <vector>

static thread_local uint64_t g_descents = 0;
#define DRG_TRACE_DESCENT() (++g_descents)

#include "../../py-draughts_b200/csrc/drg_records.cuh"

using namespace drg;

static int g_explode = kWalkExplode;
extern "C" void sim_set_explode(int k) { g_explode = k; }

namespace {

template <int S>
const Tables<S>& tables() {
  static Tables<S> t;
  static bool done = false;
  if (!done) {
    build_nb<S>(t.nb);
    build_ray_jmp<S>(t.ray, t.jmp, t.nb);
    done = true;
  }
  return t;
}

template <class R>
Pos<R> make_pos(uint64_t wm, uint64_t wk, uint64_t bm, uint64_t bk, int turn) {
  using BB = typename Pos<R>::BB;
  Pos<R> p;
  p.wm = BB(wm) & Geo<R::S>::MASK; p.wk = BB(wk) & Geo<R::S>::MASK; p.bm = BB(bm) & Geo<R::S>::MASK; p.bk = BB(bk) & Geo<R::S>::MASK;
  p.turn = turn;
  return p;
}

template <class R>
void count_one(uint64_t wm, uint64_t wk, uint64_t bm, uint64_t bk, int turn, uint32_t* count, uint32_t* descents) {
  const Pos<R> p = make_pos<R>(wm, wk, bm, bk, turn);
  Frame stack[kMaxLevels];
  uint32_t ovf = 0, stat = 0;
  g_descents = 0;
  *count = (uint32_t)count_moves<R>(p, stack, 1, tables<R::S>(), ovf, stat);
  *descents = (uint32_t)g_descents;
}

struct SimEmitSink {  // what DeepSink does, without the mask
  DrgMove* out;
  uint32_t n, room;
  void quiet(int from, int to, bool king) { if (n < room) { Rec r; rec_quiet(r, from, to, king); rec_store(out + n, r); } n++; }
  template <class BB>
  void capture(int nsq, BB capt, bool promo, bool root_king, int cur, const Frame* stack, int stride) {
    if (n < room) { Rec r; rec_capture(r, nsq, capt, promo, root_king, cur, stack, stride); rec_store(out + n, r); }
    n++;
  }
  void jump(int from, int victim, int land) { if (n < room) { Rec r; rec_jump(r, from, victim, land); rec_store(out + n, r); } n++; }
};

struct SimAlloc {
  std::vector<WalkRec>* pool;
  size_t cap;
  long long reserve(int n) {
    if (pool->size() + (size_t)n > cap) return -1;
    const long long at = (long long)pool->size();
    pool->resize(pool->size() + n);
    return at;
  }
  WalkRec* at(long long i) { return pool->data() + i; }
};

struct SplitStats {  // per round: walks, total descents, max descents of one walk
  uint64_t walks[32], descents[32], max_desc[32];
};

template <class R>
int split_run(int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm, const uint64_t* bk, const uint8_t* turn,
              const int* budgets, int nrounds, int64_t pool_cap, uint32_t* counts, const uint64_t* offsets, DrgMove* moves,
              uint64_t* stats_out) {
  const Tables<R::S>& T = tables<R::S>();
  NullSink ns;
  std::vector<uint32_t> items;              // the tree-search list
  std::vector<uint32_t> quiet(n, 0);
  // k_count phases A + B
  for (int64_t i = 0; i < n; i++) {
    const Pos<R> p = make_pos<R>(wm[i], wk[i], bm[i], bk[i], turn[i] & 1);
    MenJumps<R> mj;
    const auto c = find_capturers<R>(p, T, mj);
    uint32_t q = 0;
    if (R::kOptionalCapture || !c.any()) q = (uint32_t)gen_quiet<R, true>(p, ns, T);
    quiet[i] = q;
    counts[i] = q;
    if (!c.any()) continue;
    if (!c.kings && only_single_jumps<R>(p, mj)) counts[i] = q + (uint32_t)single_jumps<R, false>(p, mj, ns, T);
    else items.push_back((uint32_t)i);
  }
  const size_t nd = items.size();
  std::vector<WalkRec> pool;
  pool.reserve(1 << 16);
  SimAlloc alloc{&pool, (size_t)pool_cap};
  std::vector<WalkRes> rres(nd);            // root walks
  std::vector<WalkRes> wres;                // records
  std::vector<uint32_t> bword(nd, 0);       // per board: max (best << 1 | king attains it)
  std::vector<size_t> region = {0};         // region[r] .. region[r+1]: records walked in round r + 1
  SplitStats S{};
  Frame stack[kMaxLevels];
  uint32_t ovf = 0;
  // round 0: root walks (k_walk_root); kept-leaves pool of 16 records per board as in the product
  std::vector<DrgMove> keep_pool(nd * kKeepRecords + 1);
  std::vector<uint32_t> keep_stat(nd, 0);
  for (size_t k = 0; k < nd; k++) {
    const int64_t j = items[k];
    const Pos<R> p = make_pos<R>(wm[j], wk[j], bm[j], bk[j], turn[j] & 1);
    MenJumps<R> mj;
    const auto c = find_capturers<R>(p, T, mj);
    WalkLane<R, 2> lane;
    KeepSink keep;
    keep.out = keep_pool.data() + k * kKeepRecords;
    keep.n = 0;
    keep.room = kKeepRecords;
    lane.start_root(p, c, (uint32_t)k, budgets[0], T);
    lane.explode = g_explode;
    for (;;) {  // k_walk_root: a board that exceeds its budget starts again as one record per capturing piece
      const int s = lane.advance(keep, stack, 1, T, ovf);
      if (s == kWalkDone) break;
      if (s == kWalkWantsSplit && lane.abandon(c, alloc)) { keep.n = 0; break; }
    }
    rres[k] = lane.result();
    keep_stat[k] = pack_keep_stat(lane.st, keep.n, keep.n <= keep.room);
    if (walk_board_word<R>(lane.st) > bword[k]) bword[k] = walk_board_word<R>(lane.st);
    S.walks[0]++; S.descents[0] += lane.descents; if ((uint64_t)lane.descents > S.max_desc[0]) S.max_desc[0] = lane.descents;
  }
  region.push_back(pool.size());
  // rounds 1 .. : records (k_walk_recs)
  for (int r = 1; r < nrounds; r++) {
    const size_t lo = region[r - 1], hi = region[r];
    wres.resize(hi);
    for (size_t w = lo; w < hi; w++) {
      const WalkRec rec = pool[w];  // copy: the pool may grow
      const int64_t j = items[rec.item];
      const Pos<R> p = make_pos<R>(wm[j], wk[j], bm[j], bk[j], turn[j] & 1);
      WalkLane<R, 0> lane;
      if (rec.pad != r) return -5;  // records carry the round in which they are walked
      lane.start_rec(p, rec, r == nrounds - 1 ? kNoBudget : budgets[r], stack, 1, T);
      lane.explode = g_explode;
      lane.next_round = r + 1;
      lane.run(ns, stack, 1, T, ovf, alloc);
      wres[w] = lane.result();
      if (walk_board_word<R>(lane.st) > bword[rec.item]) bword[rec.item] = walk_board_word<R>(lane.st);
      S.walks[r]++; S.descents[r] += lane.descents; if ((uint64_t)lane.descents > S.max_desc[r]) S.max_desc[r] = lane.descents;
    }
    region.push_back(pool.size());
  }
  if (region.back() != region[region.size() - 2]) return -2;  // the last round may not hand over
  const size_t nw = pool.size();
  wres.resize(nw);
  // k_walk_sums: bottom-up over the rounds
  std::vector<uint32_t> sub(nw, 0), base(nw, 0);
  for (int r = nrounds - 1; r >= 1; r--) {
    for (size_t w = region[r - 1]; w < region[r]; w++) {
      uint32_t s = walk_kept<R>(wres[w], bword[pool[w].item]);
      const uint32_t nc = wres[w].info >> 24;
      for (uint32_t c = 0; c < nc; c++) s += sub[wres[w].child0 + c];
      sub[w] = s;
    }
  }
  // k_count_finish: roots
  for (size_t k = 0; k < nd; k++) {
    uint32_t run = walk_kept<R>(rres[k], bword[k]);
    const uint32_t nc = rres[k].info >> 24;
    for (uint32_t c = 0; c < nc; c++) { base[rres[k].child0 + c] = run; run += sub[rres[k].child0 + c]; }
    counts[items[k]] += run;
  }
  // k_walk_bases: top-down
  for (int r = 1; r < nrounds; r++) {
    for (size_t w = region[r - 1]; w < region[r]; w++) {
      uint32_t run = base[w] + walk_kept<R>(wres[w], bword[pool[w].item]);
      const uint32_t nc = wres[w].info >> 24;
      for (uint32_t c = 0; c < nc; c++) { base[wres[w].child0 + c] = run; run += sub[wres[w].child0 + c]; }
    }
  }
  for (int r = 0; r < 32; r++) { stats_out[3 * r] = S.walks[r]; stats_out[3 * r + 1] = S.descents[r]; stats_out[3 * r + 2] = S.max_desc[r]; }
  stats_out[96] = nd;
  stats_out[97] = nw;
  if (!moves) return ovf ? 1 : 0;
  // ---- emission: everything but the tree-search captures (k_emit), then roots (k_emit_deep), then records (k_emit_recs)
  for (int64_t i = 0; i < n; i++) {
    const Pos<R> p = make_pos<R>(wm[i], wk[i], bm[i], bk[i], turn[i] & 1);
    SimEmitSink sink{moves + offsets[i], 0, (uint32_t)(offsets[i + 1] - offsets[i])};
    MenJumps<R> mj;
    const auto c = find_capturers<R>(p, T, mj);
    if (R::kOptionalCapture || !c.any()) gen_quiet<R, false>(p, sink, T);
    if (c.any() && !c.kings && only_single_jumps<R>(p, mj)) single_jumps<R, true>(p, mj, sink, T);
  }
  for (size_t k = 0; k < nd; k++) {
    const int64_t j = items[k];
    const Pos<R> p = make_pos<R>(wm[j], wk[j], bm[j], bk[j], turn[j] & 1);
    SimEmitSink sink{moves + offsets[j], quiet[j], (uint32_t)(offsets[j + 1] - offsets[j])};
    const uint32_t kept = walk_kept<R>(rres[k], bword[k]);
    if (!kept) continue;
    const uint32_t w = keep_stat[k];
    if (w & 0x40000000u) {  // copy the kept leaves
      const int nk = (int)((w >> 20) & 0xFFu);
      const bool kings_only = R::kFilter == FILTER_VALUE && (bword[k] & 1u);
      for (int r = 0; r < nk; r++) {
        const DrgMove& m = keep_pool[k * kKeepRecords + r];
        if (kings_only && !(m.flags & DRG_MF_KINGMOVE)) continue;
        if (sink.n < sink.room) sink.out[sink.n] = m;
        sink.n++;
      }
    } else {
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      WalkLane<R, 1> lane;
      lane.start_root(p, c, (uint32_t)k, (rres[k].info & kWalkSplit) ? budgets[0] : kNoBudget, T);
      lane.st.best = (int)(bword[k] >> 1);
      lane.st.n_best_king = (int)(bword[k] & 1u);
      lane.run(sink, stack, 1, T, ovf, alloc);
    }
    if (sink.n - quiet[j] != kept) return -3;
  }
  for (int r = 1; r < nrounds; r++) {
    for (size_t w = region[r - 1]; w < region[r]; w++) {
      const WalkRec rec = pool[w];
      const uint32_t kept = walk_kept<R>(wres[w], bword[rec.item]);
      if (!kept) continue;
      const int64_t j = items[rec.item];
      const Pos<R> p = make_pos<R>(wm[j], wk[j], bm[j], bk[j], turn[j] & 1);
      SimEmitSink sink{moves + offsets[j], quiet[j] + base[w], (uint32_t)(offsets[j + 1] - offsets[j])};
      WalkLane<R, 1> lane;
      const int bud = rec.pad == nrounds - 1 ? kNoBudget : budgets[rec.pad];
      lane.start_rec(p, rec, (wres[w].info & kWalkSplit) ? bud : kNoBudget, stack, 1, T);
      lane.st.best = (int)(bword[rec.item] >> 1);
      lane.st.n_best_king = (int)(bword[rec.item] & 1u);
      lane.run(sink, stack, 1, T, ovf, alloc);
      if (sink.n - quiet[j] - base[w] != kept) return -4;
    }
  }
  return ovf ? 1 : 0;
}

}  // namespace

extern "C" int sim_count(int family, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                         const uint64_t* bk, const uint8_t* turn, uint32_t* counts, uint32_t* descents) {
  for (int64_t i = 0; i < n; i++) {
    switch (family) {
      case 0: count_one<RulesStandard>(wm[i], wk[i], bm[i], bk[i], turn[i] & 1, counts + i, descents + i); break;
      case 1: count_one<RulesAmerican>(wm[i], wk[i], bm[i], bk[i], turn[i] & 1, counts + i, descents + i); break;
      case 2: count_one<RulesFrisian>(wm[i], wk[i], bm[i], bk[i], turn[i] & 1, counts + i, descents + i); break;
      case 3: count_one<RulesRussian>(wm[i], wk[i], bm[i], bk[i], turn[i] & 1, counts + i, descents + i); break;
      case 4: count_one<RulesBrazilian>(wm[i], wk[i], bm[i], bk[i], turn[i] & 1, counts + i, descents + i); break;
      default: return -1;
    }
  }
  return 0;
}

// moves == NULL: counting only.  stats_out: uint64[98] = 32 x (walks, descents, max descents of one walk) per round,
// [96] boards in the tree-search list, [97] records handed over.
extern "C" int sim_split(int family, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                         const uint64_t* bk, const uint8_t* turn, const int* budgets, int nrounds, int64_t pool_cap,
                         uint32_t* counts, const uint64_t* offsets, DrgMove* moves, uint64_t* stats_out) {
  switch (family) {
    case 0: return split_run<RulesStandard>(n, wm, wk, bm, bk, turn, budgets, nrounds, pool_cap, counts, offsets, moves, stats_out);
    case 1: return split_run<RulesAmerican>(n, wm, wk, bm, bk, turn, budgets, nrounds, pool_cap, counts, offsets, moves, stats_out);
    case 2: return split_run<RulesFrisian>(n, wm, wk, bm, bk, turn, budgets, nrounds, pool_cap, counts, offsets, moves, stats_out);
    case 3: return split_run<RulesRussian>(n, wm, wk, bm, bk, turn, budgets, nrounds, pool_cap, counts, offsets, moves, stats_out);
    case 4: return split_run<RulesBrazilian>(n, wm, wk, bm, bk, turn, budgets, nrounds, pool_cap, counts, offsets, moves, stats_out);
  }
  return -1;
}
