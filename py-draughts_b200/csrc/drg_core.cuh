// drg_core.cuh -- per-board move generation / make-move device code (sm_100a).
//
// What it computes is what the reference computes (file:line are relative to the reference
// root), how it computes it is new:
//   * quiet moves  (standard.py:107-162, american.py:98-149)   -> bit-parallel shifts + popc/ctz
//   * capture trees (standard.py:184-274, american.py:173-243, russian.py:221-340,
//     brazilian.py:41-77, frisian.py:361-609)                   -> ONE iterative DFS per piece with
//     an explicit stack of 4-byte frames in shared memory (no recursion, no board mutation: the
//     occupancy of a DFS node is static_occ & ~captured | bit(cur)).
//   * the reference's local pruning inside king searches (standard.py:265-270,
//     frisian.py:539-542) is result-neutral once the top-level filter is applied (every node
//     keeps exactly the leaves of its best metric, and the final filter keeps the global best), so
//     the filter is done with a statistics pass (best metric + leaf counts) and, only when
//     moves must be materialised, a second emitting pass.
//   * push (base.py:215-293) -> apply_move(), copy-make.
//
// Canonical order: captures of men (ascending square) then kings (ascending square),
// each in DFS pre-order over directions 0..3 (then 4..7 for Frisian), flying-king landings near
// to far; quiet moves: four man passes (UR-even, UR-odd, UL-even, UL-odd for white; DR-even,
// DR-odd, DL-even, DL-odd for black) each by ascending target, then kings ascending, rays 0..3
// near to far.  A single thread walks one board in exactly that order, so order needs no sort.
#pragma once
#include <stdint.h>

#include "../../include/drg.h"
#include "drg_geom.h"

namespace drg {

// ---------------------------------------------------------------------------------------------
// rule sets (variant rule matrix)
// ---------------------------------------------------------------------------------------------
enum Filter { FILTER_NONE = 0, FILTER_MAXLEN = 1, FILTER_VALUE = 2 };

struct RulesStandard {  // standard.py, antidraughts.py, breakthrough.py
  static constexpr int S = 50;
  static constexpr bool kOrtho = false, kFlying = true, kMenBackward = true, kMidPromo = false;
  static constexpr bool kOptionalCapture = false;
  static constexpr int kFilter = FILTER_MAXLEN;
};
struct RulesFrisian {  // frisian.py, frysk.py
  static constexpr int S = 50;
  static constexpr bool kOrtho = true, kFlying = true, kMenBackward = true, kMidPromo = false;
  static constexpr bool kOptionalCapture = false;
  static constexpr int kFilter = FILTER_VALUE;
};
struct RulesAmerican {  // american.py
  static constexpr int S = 32;
  static constexpr bool kOrtho = false, kFlying = false, kMenBackward = false, kMidPromo = false;
  static constexpr bool kOptionalCapture = true;
  static constexpr int kFilter = FILTER_NONE;
};
struct RulesRussian {  // russian.py
  static constexpr int S = 32;
  static constexpr bool kOrtho = false, kFlying = true, kMenBackward = true, kMidPromo = true;
  static constexpr bool kOptionalCapture = false;
  static constexpr int kFilter = FILTER_NONE;
};
struct RulesBrazilian {  // brazilian.py
  static constexpr int S = 32;
  static constexpr bool kOrtho = false, kFlying = true, kMenBackward = true, kMidPromo = false;
  static constexpr bool kOptionalCapture = false;
  static constexpr int kFilter = FILTER_MAXLEN;
};

// ---------------------------------------------------------------------------------------------
// bit helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int bb_ctz(uint32_t x) { return __ffs((int)x) - 1; }
__device__ __forceinline__ int bb_ctz(uint64_t x) { return __ffsll((long long)x) - 1; }
__device__ __forceinline__ int bb_popc(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int bb_popc(uint64_t x) { return __popcll(x); }
template <class BB>
__device__ __forceinline__ BB bb_bit(int sq) { return BB(1) << sq; }

// DFS depth: frames 0..kMaxLevels-1 are stored, so a path has at most kMaxLevels+1 = DRG_MAX_PATH squares.
constexpr int kMaxLevels = DRG_MAX_PATH - 1;

// 4-byte DFS frame (shared memory, one column per thread: frame(l) = stack[l * stride])
struct __align__(4) Frame {
  uint8_t sq;       // square of the mover at this level (= path[level])
  uint8_t dirflag;  // bits 0-3 next direction to try, bit 4 has_child, bit 5 mover is a king at this level
  uint8_t landcur;  // flying king: last landing square handed out on direction `dir` (kNone = none yet)
  uint8_t victim;   // victim of the child currently being explored
};

template <class R>
struct Pos {
  using BB = typename Geo<R::S>::BB;
  BB wm, wk, bm, bk;
  int turn;  // 0 white, 1 black
  __device__ __forceinline__ BB own_men() const { return turn ? bm : wm; }
  __device__ __forceinline__ BB own_kings() const { return turn ? bk : wk; }
  __device__ __forceinline__ BB opp_men() const { return turn ? wm : bm; }
  __device__ __forceinline__ BB opp_kings() const { return turn ? wk : bk; }
  __device__ __forceinline__ BB occ() const { return wm | wk | bm | bk; }
};

struct CapStat {
  int best;         // best metric over capture leaves (captures for MAXLEN, value for VALUE)
  int n_best_man;   // leaves with the best metric started by a man
  int n_best_king;  // ... started by a king
  int n_all;        // all leaves (PASS 0)
  int n_emit;       // leaves handed to the sink (PASS 1)
  __device__ __forceinline__ void reset() { best = 0; n_best_man = n_best_king = n_all = n_emit = 0; }
};

// ---------------------------------------------------------------------------------------------
// capture DFS for the piece standing on `origin`
//   PASS 0: statistics only (st updated);  PASS 1: leaves passing the filter go to sink.capture().
// ---------------------------------------------------------------------------------------------
template <class R, int PASS, class Sink>
__device__ __forceinline__ void capture_dfs(const Pos<R>& p, int origin, bool root_king, CapStat& st, Sink& sink,
                                            Frame* stack, int stride, const uint8_t* __restrict__ nb,
                                            uint32_t& overflow) {
  using BB = typename Pos<R>::BB;
  using G = Geo<R::S>;
  const BB enemy0 = p.opp_men() | p.opp_kings();
  const BB obit = bb_bit<BB>(origin);
  // board minus the mover (the mover's own bitboard loses the origin bit only, base semantics of
  // `self.white_men = (wm & ~src_bit) | land_bit`, standard.py:201-204)
  const BB static_occ = root_king ? BB(p.own_men() | (p.own_kings() & ~obit) | enemy0)
                                  : BB((p.own_men() & ~obit) | p.own_kings() | enemy0);
  const BB promo_rank = p.turn ? G::ROW_LAST : G::ROW_FIRST;
  // direction window of a jumping mover: men of American checkers capture forward only (american.py:179)
  const int man_d0 = R::kMenBackward ? 0 : (p.turn ? 2 : 0);
  const int man_d1 = R::kMenBackward ? (R::kOrtho ? 8 : 4) : (p.turn ? 4 : 2);
  const int king_d1 = R::kOrtho ? 8 : 4;

  int level = 0, cur = origin;
  bool king = root_king, has_child = false;
  int dir = king ? 0 : man_d0;
  int landcur = kNone, fvictim = kNone;
  BB capt = 0, forb = 0;
  int value = 0;

  for (;;) {
    const BB allp = BB((static_occ & ~capt) | bb_bit<BB>(cur));
    bool found = false;
    int victim = kNone, land = kNone;
    if (!(king && R::kFlying)) {
      // man (any variant) or short king (american.py:210-243): jump over an adjacent enemy
      const int dend = king ? 4 : man_d1;
      while (dir < dend) {
        const int d = dir++;
        const int mid = nb[cur * 8 + d];
        if (mid == kNone) continue;
        const int l2 = nb[mid * 8 + d];
        if (l2 == kNone) continue;
        const BB mbit = bb_bit<BB>(mid);
        if ((capt & mbit) || !(enemy0 & mbit)) continue;
        if (allp & bb_bit<BB>(l2)) continue;
        victim = mid;
        land = l2;
        found = true;
        break;
      }
    } else {
      // flying king: first piece on the ray must be an uncaptured enemy; every free square behind it is a landing
      while (dir < king_d1) {
        if (landcur == kNone) {
          int t = nb[cur * 8 + dir];
          victim = kNone;
          while (t != kNone) {
            const BB tb = bb_bit<BB>(t);
            if (forb & tb) break;
            if (allp & tb) {
              if (!(capt & tb) && (enemy0 & tb)) victim = t;
              break;
            }
            t = nb[t * 8 + dir];
          }
          if (victim == kNone) { dir++; continue; }
          land = nb[victim * 8 + dir];
        } else {
          victim = fvictim;
          land = nb[landcur * 8 + dir];
        }
        if (land == kNone || ((forb | allp) & bb_bit<BB>(land))) { landcur = kNone; dir++; continue; }
        landcur = land;
        found = true;
        break;
      }
    }
    if (found && level >= kMaxLevels) {  // path would not fit a record: report, do not descend
      overflow = 1;
      found = false;
      continue;
    }
    if (found) {
      Frame f;
      f.sq = (uint8_t)cur;
      f.dirflag = (uint8_t)(dir | 0x10 | (king ? 0x20 : 0));
      f.landcur = (uint8_t)landcur;
      f.victim = (uint8_t)victim;
      stack[level * stride] = f;
      const BB vb = bb_bit<BB>(victim);
      capt |= vb;
      if (king && R::kFlying) forb |= vb;  // kings may not cross a captured piece (standard.py:253-254)
      if (R::kFilter == FILTER_VALUE) {
        // cap_piece priority bm, bk, wm, wk (frisian.py:392-393)
        const bool vk = (p.bm & vb) ? false : ((p.bk & vb) ? true : ((p.wm & vb) ? false : true));
        value += vk ? 199 : 100;
      }
      // mid-capture promotion (russian.py:262-271): the man continues as a flying king; its
      // forbidden set starts empty, earlier victims are already off the board
      if (R::kMidPromo && !king && (bb_bit<BB>(land) & promo_rank)) king = true;
      level++;
      cur = land;
      dir = king ? 0 : man_d0;
      landcur = kNone;
      has_child = false;
      continue;
    }
    // no (more) children here
    if (!has_child && level > 0) {
      const int metric = R::kFilter == FILTER_VALUE ? value : level;
      if (PASS == 0) {
        st.n_all++;
        if (metric > st.best) { st.best = metric; st.n_best_man = 0; st.n_best_king = 0; }
        if (metric == st.best) { if (root_king) st.n_best_king++; else st.n_best_man++; }
      } else {
        bool keep = true;
        if (R::kFilter == FILTER_MAXLEN) keep = metric == st.best;
        if (R::kFilter == FILTER_VALUE) keep = metric == st.best && (root_king || st.n_best_king == 0);
        if (keep) {
          st.n_emit++;
          sink.capture(level + 1, capt, king && !root_king, root_king, cur, stack, stride);
        }
      }
    }
    if (level == 0) break;
    level--;
    const Frame f = stack[level * stride];
    cur = f.sq;
    dir = f.dirflag & 0x0F;
    has_child = true;
    king = (f.dirflag & 0x20) != 0;
    landcur = f.landcur;
    fvictim = f.victim;
    const BB vb = bb_bit<BB>(fvictim);
    capt &= ~vb;
    if (king && R::kFlying) forb &= ~vb;
    if (R::kFilter == FILTER_VALUE) {
      const bool vk = (p.bm & vb) ? false : ((p.bk & vb) ? true : ((p.wm & vb) ? false : true));
      value -= vk ? 199 : 100;
    }
  }
}

// men that can make at least a first jump (bit-parallel root test of standard.py:188-197)
template <class R, int D>
__device__ __forceinline__ typename Pos<R>::BB jumpers_dir(typename Pos<R>::BB men, typename Pos<R>::BB enemy,
                                                           typename Pos<R>::BB empty) {
  using BB = typename Pos<R>::BB;
  const BB over = step<R::S, D>(men) & enemy;
  const BB land = step<R::S, D>(over) & empty;
  return step<R::S, opposite(D)>(step<R::S, opposite(D)>(land));
}

template <class R>
__device__ __forceinline__ typename Pos<R>::BB men_jumpers(const Pos<R>& p) {
  using BB = typename Pos<R>::BB;
  const BB men = p.own_men(), enemy = p.opp_men() | p.opp_kings();
  const BB empty = BB(~p.occ() & Geo<R::S>::MASK);
  BB j = 0;
  if (R::kMenBackward || p.turn == 0) {
    j |= jumpers_dir<R, 0>(men, enemy, empty);
    j |= jumpers_dir<R, 1>(men, enemy, empty);
  }
  if (R::kMenBackward || p.turn == 1) {
    j |= jumpers_dir<R, 2>(men, enemy, empty);
    j |= jumpers_dir<R, 3>(men, enemy, empty);
  }
  if (R::kOrtho) {
    j |= jumpers_dir<R, 4>(men, enemy, empty);
    j |= jumpers_dir<R, 5>(men, enemy, empty);
    j |= jumpers_dir<R, 6>(men, enemy, empty);
    j |= jumpers_dir<R, 7>(men, enemy, empty);
  }
  return j & men;
}

// all captures of the side to move, men first then kings (standard.py:164-182)
template <class R, int PASS, class Sink>
__device__ __forceinline__ void gen_captures(const Pos<R>& p, CapStat& st, Sink& sink, Frame* stack, int stride,
                                             const uint8_t* __restrict__ nb, uint32_t& overflow) {
  using BB = typename Pos<R>::BB;
  if (!(p.opp_men() | p.opp_kings())) return;  // standard.py:167-169
  BB bb = men_jumpers<R>(p);
  while (bb) {
    const int sq = bb_ctz(bb);
    bb &= bb - 1;
    capture_dfs<R, PASS>(p, sq, false, st, sink, stack, stride, nb, overflow);
  }
  bb = p.own_kings();
  while (bb) {
    const int sq = bb_ctz(bb);
    bb &= bb - 1;
    capture_dfs<R, PASS>(p, sq, true, st, sink, stack, stride, nb, overflow);
  }
}

// quiet moves (standard.py:107-162 / american.py:98-149); COUNT_ONLY returns the count without visiting
template <class R, bool COUNT_ONLY, class Sink>
__device__ __forceinline__ int gen_quiet(const Pos<R>& p, Sink& sink, const uint8_t* __restrict__ nb) {
  using BB = typename Pos<R>::BB;
  using G = Geo<R::S>;
  constexpr int W = G::W;
  const BB empty = BB(~p.occ() & G::MASK);
  const BB men = p.own_men();
  BB t[4];
  int delta[4];  // from = target + delta
  if (p.turn == 0) {
    const BB even = men & G::EVEN & ~G::ROW_FIRST, odd = men & G::ODD;
    t[0] = BB((even & ~G::EVEN_RIGHT) >> (W - 1)); delta[0] = W - 1;
    t[1] = BB(odd >> W);                          delta[1] = W;
    t[2] = BB(even >> W);                         delta[2] = W;
    t[3] = BB((odd & ~G::ODD_LEFT) >> (W + 1));   delta[3] = W + 1;
  } else {
    const BB even = men & G::EVEN, odd = men & G::ODD & ~G::ROW_LAST;
    t[0] = BB((even & ~G::EVEN_RIGHT) << (W + 1)); delta[0] = -(W + 1);
    t[1] = BB(odd << W);                           delta[1] = -W;
    t[2] = BB(even << W);                          delta[2] = -W;
    t[3] = BB((odd & ~G::ODD_LEFT) << (W - 1));    delta[3] = -(W - 1);
  }
  int n = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    BB bb = BB(t[k] & empty & G::MASK);
    if (COUNT_ONLY) {
      n += bb_popc(bb);
    } else {
      while (bb) {
        const int to = bb_ctz(bb);
        bb &= bb - 1;
        sink.quiet(to + delta[k], to, false);
        n++;
      }
    }
  }
  BB kings = p.own_kings();
  while (kings) {
    const int sq = bb_ctz(kings);
    kings &= kings - 1;
#pragma unroll 1
    for (int d = 0; d < 4; d++) {
      int to = nb[sq * 8 + d];
      if (R::kFlying) {
        while (to != kNone && (empty & bb_bit<BB>(to))) {
          if (!COUNT_ONLY) sink.quiet(sq, to, true);
          n++;
          to = nb[to * 8 + d];
        }
      } else if (to != kNone && (empty & bb_bit<BB>(to))) {
        if (!COUNT_ONLY) sink.quiet(sq, to, true);
        n++;
      }
    }
  }
  return n;
}

// A sink that ignores moves (statistics / counting passes)
struct NullSink {
  __device__ __forceinline__ void quiet(int, int, bool) {}
  template <class BB>
  __device__ __forceinline__ void capture(int, BB, bool, bool, int, const Frame*, int) {}
};

// len(board.legal_moves)
template <class R>
__device__ __forceinline__ int count_moves(const Pos<R>& p, Frame* stack, int stride, const uint8_t* __restrict__ nb,
                                           uint32_t& overflow) {
  NullSink ns;
  CapStat st;
  st.reset();
  gen_captures<R, 0>(p, st, ns, stack, stride, nb, overflow);
  int ncap;
  if (R::kFilter == FILTER_NONE) ncap = st.n_all;
  else if (R::kFilter == FILTER_MAXLEN) ncap = st.n_best_man + st.n_best_king;
  else ncap = st.n_best_king ? st.n_best_king : st.n_best_man;  // frisian.py:264-268
  if (R::kOptionalCapture) return gen_quiet<R, true>(p, ns, nb) + ncap;  // american.py:96
  if (ncap) return ncap;
  return gen_quiet<R, true>(p, ns, nb);
}

// board.legal_moves in canonical order into `sink`; returns the number of moves
template <class R, class Sink>
__device__ __forceinline__ int generate_moves(const Pos<R>& p, Sink& sink, Frame* stack, int stride,
                                              const uint8_t* __restrict__ nb, uint32_t& overflow) {
  CapStat st;
  st.reset();
  if (R::kFilter == FILTER_NONE) {
    // every capture leaf is legal: one emitting pass (russian.py:126-129, american.py:96)
    int nq = 0;
    if (R::kOptionalCapture) nq = gen_quiet<R, false>(p, sink, nb);
    gen_captures<R, 1>(p, st, sink, stack, stride, nb, overflow);
    if (R::kOptionalCapture) return nq + st.n_emit;
    if (st.n_emit) return st.n_emit;
    return gen_quiet<R, false>(p, sink, nb);
  }
  gen_captures<R, 0>(p, st, sink, stack, stride, nb, overflow);
  if (st.n_all) {
    gen_captures<R, 1>(p, st, sink, stack, stride, nb, overflow);
    return st.n_emit;
  }
  return gen_quiet<R, false>(p, sink, nb);
}

// ---------------------------------------------------------------------------------------------
// make-move: BaseBoard.push(move, is_finished=True)  base.py:215-293
// returns false where the reference raises ValueError (base.py:241-245); the board is then unchanged
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ bool apply_move(Pos<R>& p, int& clock, int from, int to, typename Pos<R>::BB captured,
                                           bool promo_flag) {
  using BB = typename Pos<R>::BB;
  using G = Geo<R::S>;
  const BB sb = bb_bit<BB>(from), tb = bb_bit<BB>(to);
  // BaseBoard._get priority: wm, wk, bm, bk (base.py:152-163)
  int piece = 0;
  if (p.wm & sb) piece = -1;
  else if (p.wk & sb) piece = -2;
  else if (p.bm & sb) piece = 1;
  else if (p.bk & sb) piece = 2;
  if (piece == 0 || (piece < 0) != (p.turn == 0)) return false;
  if (piece == -1) p.wm = BB((p.wm & ~sb) | tb);
  else if (piece == -2) p.wk = BB((p.wk & ~sb) | tb);
  else if (piece == 1) p.bm = BB((p.bm & ~sb) | tb);
  else p.bk = BB((p.bk & ~sb) | tb);
  bool promoted = false;
  if (piece == -1 && ((G::ROW_FIRST & tb) || promo_flag)) { p.wm &= ~tb; p.wk |= tb; promoted = true; }
  else if (piece == 1 && ((G::ROW_LAST & tb) || promo_flag)) { p.bm &= ~tb; p.bk |= tb; promoted = true; }
  if (!promoted && (piece == 2 || piece == -2) && captured == 0) clock += 1;
  else clock = 0;
  const BB rem = BB(~(captured & ~tb));  // `if cap_sq != tgt` (base.py:284)
  p.wm &= rem; p.wk &= rem; p.bm &= rem; p.bk &= rem;
  p.turn ^= 1;
  return true;
}

}  // namespace drg
