// drg_core.cuh -- per-board move generation / make-move device code (sm_100a).
//
// What it computes is what the reference computes (file:line are relative to the reference
// root), how it computes it is new:
//   * quiet moves  (standard.py:107-162, american.py:98-149)   -> bit-parallel shifts + popc/ctz for
//     men, ray masks + clz/ctz for flying kings (no per-square walking).
//   * capture pre-check: which men / kings can make a FIRST jump is decided bit-parallel (men) or
//     with one ray-mask probe per direction (kings); it is exact, so boards without captures never
//     enter the tree search and boards with captures can be compacted onto dense warps.
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

// hook of tools/devsim (host replay of the per-thread algorithm): counts DFS descents; empty in the product build
#ifndef DRG_TRACE_DESCENT
#define DRG_TRACE_DESCENT()
#endif

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
__device__ __forceinline__ int bb_msb(uint32_t x) { return 31 - __clz((int)x); }
__device__ __forceinline__ int bb_msb(uint64_t x) { return 63 - __clzll((long long)x); }
__device__ __forceinline__ int bb_popc(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int bb_popc(uint64_t x) { return __popcll(x); }
template <class BB>
__device__ __forceinline__ BB bb_bit(int sq) { return BB(1) << sq; }

// directions whose squares have LOWER indices than the origin: up-right, up-left, up, left
__device__ __forceinline__ bool dir_is_up(int d) { return d < 2 || d == 4 || d == 7; }

// first square of `blk` (non-zero, all on one ray from the origin) met when walking direction d
template <class BB>
__device__ __forceinline__ int nearest_on_ray(BB blk, int d) { return dir_is_up(d) ? bb_msb(blk) : bb_ctz(blk); }

// squares of `ray` strictly before the nearest blocker (all of `ray` if there is none)
template <class BB>
__device__ __forceinline__ BB ray_until_blocker(BB ray, BB blk, int d) {
  if (dir_is_up(d)) return blk ? BB(ray & ~BB((BB(2) << bb_msb(blk)) - 1)) : ray;
  return BB(ray & BB((blk & (0 - blk)) - 1));
}

// DFS depth: frames 0..kMaxLevels-1 are stored, so a path has at most kMaxLevels+1 = DRG_MAX_PATH squares.
constexpr int kMaxLevels = DRG_MAX_PATH - 1;

// 4-byte DFS frame (shared memory, one column per thread: frame(l) = stack[l * stride])
struct __align__(4) Frame {
  uint8_t sq;       // square of the mover at this level (= path[level])
  uint8_t dirflag;  // directions of this node that still have an unvisited child (bit d = direction d)
  uint8_t landcur;  // flying king: last landing square handed out on the lowest direction of dirflag (kNone = none yet)
  uint8_t victim;   // bits 0-6 victim of the child being explored, bit 7 the mover is a king at this level
};

// geometry tables of one board size, resident in shared memory (built on the host, drg_geom.h)
template <int S>
struct Tables {
  using BB = typename Geo<S>::BB;
  // S rows only (4 000 B for S = 50): with 64 the mask-row kernel's block is 460 bytes too large for five blocks per SM
  BB ray[S * 8];        // ray[sq*8+d]: every square met walking direction d from sq (KING_RAYS / ORTHO_RAYS as masks)
  uint8_t nb[S * 8];    // nb[sq*8+d]: neighbour (MOVE_TGT / first ORTHO ray square), kNone off board
  uint8_t jmp[S * 8];   // jmp[sq*8+d] = nb[nb[sq][d]][d] (JUMP_TGT / ORTHO_JUMP landing)
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

// pieces that can start a capture (exact root test of the generators)
template <class BB>
struct Capturers {
  BB men, kings;
  __device__ __forceinline__ bool any() const { return (men | kings) != 0; }
};

// ---------------------------------------------------------------------------------------------
// capture DFS for the piece standing on `origin`
//   PASS 0: statistics only (st updated);  PASS 1: leaves passing the filter go to sink.capture();
//   PASS 2: statistics, and the leaves of the best metric SO FAR go to the sink (sink.restart() when it improves).
// On entering a node ALL directions are probed at once (independent shared-memory lookups in flight together)
// and the result is an 8-bit mask `vd` of directions that have a child; the children loop then only visits
// real children (the 4-probes-per-node walk it replaces was a chain of dependent lookups, profiles/README.md).
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ uint32_t jump_dirs(const Tables<R::S>& T, int cur, typename Pos<R>::BB capt,
                                              typename Pos<R>::BB enemy0, typename Pos<R>::BB allp, uint32_t allowed) {
  using BB = typename Pos<R>::BB;
  uint32_t vd = 0;
#pragma unroll
  for (int d = 0; d < (R::kOrtho ? 8 : 4); d++) {
    const int mid = T.nb[cur * 8 + d];
    const int l2 = T.jmp[cur * 8 + d];  // kNone whenever mid is
    constexpr int kBits = (int)sizeof(BB) * 8 - 1;  // off-board (kNone) indices are masked; l2 != kNone guards them
    const BB mbit = bb_bit<BB>(mid & kBits), lbit = bb_bit<BB>(l2 & kBits);
    const bool ok = l2 != kNone && !(capt & mbit) && (enemy0 & mbit) && !(allp & lbit);
    vd |= ok ? (1u << d) : 0u;
  }
  return vd & allowed;
}

template <class R>
__device__ __forceinline__ uint32_t fly_dirs(const Tables<R::S>& T, int cur, typename Pos<R>::BB capt,
                                             typename Pos<R>::BB forb, typename Pos<R>::BB enemy0,
                                             typename Pos<R>::BB allp) {
  using BB = typename Pos<R>::BB;
  uint32_t vd = 0;
  const BB block = BB(allp | forb);
#pragma unroll
  for (int d = 0; d < (R::kOrtho ? 8 : 4); d++) {
    const BB blk = BB(block & T.ray[cur * 8 + d]);
    if (blk) {
      const int t = nearest_on_ray(blk, d);
      const BB tb = bb_bit<BB>(t);
      if (!(forb & tb) && !(capt & tb) && (enemy0 & tb)) {
        const int land = T.nb[t * 8 + d];
        if (land != kNone && !(block & bb_bit<BB>(land))) vd |= 1u << d;
      }
    }
  }
  return vd;
}

template <class R, int PASS, class Sink>
__device__ __forceinline__ void capture_dfs(const Pos<R>& p, int origin, bool root_king, CapStat& st, Sink& sink,
                                            Frame* stack, int stride, const Tables<R::S>& T, uint32_t& overflow) {
  using BB = typename Pos<R>::BB;
  using G = Geo<R::S>;
  const BB enemy0 = p.opp_men() | p.opp_kings();
  const BB obit = bb_bit<BB>(origin);
  // board minus the mover (the mover's own bitboard loses the origin bit only, base semantics of
  // `self.white_men = (wm & ~src_bit) | land_bit`, standard.py:201-204)
  const BB static_occ = root_king ? BB(p.own_men() | (p.own_kings() & ~obit) | enemy0)
                                  : BB((p.own_men() & ~obit) | p.own_kings() | enemy0);
  const BB promo_rank = p.turn ? G::ROW_LAST : G::ROW_FIRST;
  // directions a jumping man may use: men of American checkers capture forward only (american.py:179)
  const uint32_t man_dirs = R::kMenBackward ? (R::kOrtho ? 0xFFu : 0x0Fu) : (p.turn ? 0x0Cu : 0x03u);

  int level = 0, cur = origin;
  bool king = root_king, has_child = false;
  int landcur = kNone, fvictim = kNone;
  BB capt = 0, forb = 0;
  int value = 0;
  BB allp = BB(static_occ | obit);
  uint32_t vd = (king && R::kFlying) ? fly_dirs<R>(T, cur, capt, forb, enemy0, allp)
                                     : jump_dirs<R>(T, cur, capt, enemy0, allp, king ? 0x0Fu : man_dirs);

  for (;;) {
    bool found = false;
    int victim = kNone, land = kNone;
    if (!(king && R::kFlying)) {
      // man (any variant) or short king (american.py:210-243): jump over an adjacent enemy
      if (vd) {
        const int d = __ffs((int)vd) - 1;
        vd &= vd - 1;
        victim = T.nb[cur * 8 + d];
        land = T.jmp[cur * 8 + d];
        found = true;
      }
    } else {
      // flying king: the first piece on the ray is an uncaptured enemy (checked by fly_dirs; a captured one
      // blocks, standard.py:228-233); every free square behind it, up to the next blocker, is a landing
      while (vd) {
        const int d = __ffs((int)vd) - 1;
        if (landcur == kNone) {
          victim = nearest_on_ray(BB((allp | forb) & T.ray[cur * 8 + d]), d);
          land = T.nb[victim * 8 + d];  // free: fly_dirs checked it
        } else {
          victim = fvictim;
          land = T.nb[landcur * 8 + d];
          if (land == kNone || ((forb | allp) & bb_bit<BB>(land))) { landcur = kNone; vd &= vd - 1; continue; }
        }
        landcur = land;
        found = true;
        break;
      }
    }
    if (found && level >= kMaxLevels) {  // path would not fit a record: report, do not descend
      overflow = 1;
      continue;  // jump mover: the direction is already consumed; flying king: the next landing comes next
    }
    if (found) {
      Frame f;
      f.sq = (uint8_t)cur;
      f.dirflag = (uint8_t)vd;
      f.landcur = (uint8_t)landcur;
      f.victim = (uint8_t)(victim | (king ? 0x80 : 0));
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
      DRG_TRACE_DESCENT();
      cur = land;
      landcur = kNone;
      has_child = false;
      allp = BB((static_occ & ~capt) | bb_bit<BB>(cur));
      vd = (king && R::kFlying) ? fly_dirs<R>(T, cur, capt, forb, enemy0, allp)
                                : jump_dirs<R>(T, cur, capt, enemy0, allp, king ? 0x0Fu : man_dirs);
      continue;
    }
    // no (more) children here
    if (!has_child && level > 0) {
      const int metric = R::kFilter == FILTER_VALUE ? value : level;
      if (PASS == 0 || PASS == 2) {
        st.n_all++;
        if (metric > st.best) {
          st.best = metric;
          st.n_best_man = 0;
          st.n_best_king = 0;
          if constexpr (PASS == 2) {
            if (R::kFilter != FILTER_NONE) sink.restart();  // the leaves kept so far are no longer the best
          }
        }
        if (metric == st.best) { if (root_king) st.n_best_king++; else st.n_best_man++; }
        // PASS 2: statistics AND the leaves of the running best go to the sink (a small per-board pool), so that the
        // emitting step can copy them instead of walking the tree again
        if constexpr (PASS == 2) {
          if (R::kFilter == FILTER_NONE || metric == st.best)
            sink.capture(level + 1, capt, king && !root_king, root_king, cur, stack, stride);
        }
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
    vd = f.dirflag;
    has_child = true;
    king = (f.victim & 0x80) != 0;
    landcur = f.landcur;
    fvictim = f.victim & 0x7F;
    const BB vb = bb_bit<BB>(fvictim);
    capt &= ~vb;
    if (king && R::kFlying) forb &= ~vb;
    if (R::kFilter == FILTER_VALUE) {
      const bool vk = (p.bm & vb) ? false : ((p.bk & vb) ? true : ((p.wm & vb) ? false : true));
      value -= vk ? 199 : 100;
    }
    allp = BB((static_occ & ~capt) | bb_bit<BB>(cur));
  }
}

// ---------------------------------------------------------------------------------------------
// Split walk: the same capture search as capture_dfs, as a resumable state machine with a descent budget.
//
// Why: one thread walking one board's whole tree is fine for the usual 3-descent tree, but a Frisian king endgame
// can have a tree of 500 000 nodes (43 000 legal captures), and a batch waits for its largest tree.  A walker that
// has spent its budget hands over everything it has not explored yet -- the open frames of its stack, deepest
// first, and the pieces it has not started -- as WalkRec records to the next round, where each record is walked by
// its own thread (again with a budget).  The records of one walker are written in canonical order, and a walk's own
// leaves precede those of its records, so the leaves of a board are ordered by a pre-order of the walk tree: the
// emitting step places them with two small tree sums (drg_kernels.cuh, k_walk_*), no sort.  The traversal never
// depends on anything but the record and the budget, so the emitting re-walk of a record visits exactly the nodes
// the counting walk visited and stops where it stopped.
// ---------------------------------------------------------------------------------------------
enum : uint8_t { WF_ROOT_KING = 1, WF_KING = 2, WF_HAS_CHILD = 4, WF_FRESH = 8, WF_PROBE = 16 };

struct __align__(16) WalkRec {
  uint64_t capt;               // victims taken on the way to the resume node
  uint64_t forb;               // squares a flying king may not cross (victims it took as a king)
  uint32_t item;               // index of the board in the tree-search list
  uint8_t level;               // captures made so far = index of the resume node in path
  uint8_t dirflag;             // directions of the resume node that may still have an unvisited child
  uint8_t landcur;             // flying king: last landing handed out on the lowest open direction (kNone: none yet)
  uint8_t fvictim;             // ... and the victim it jumps there
  uint8_t flags;               // WF_*
  uint8_t path[DRG_MAX_PATH];  // squares path[0 .. level]
  uint8_t pad;                 // round in which the record is walked
};
static_assert(sizeof(WalkRec) == 48, "WalkRec is stored as three 16-byte words");

// a record, assembled in registers and stored as three 16-byte words
__device__ __forceinline__ void store_rec(WalkRec& r, uint64_t capt, uint64_t forb, uint32_t item, int level, uint32_t dirs,
                                          int landcur, int fvictim, uint32_t flags, const uint8_t* path, int round) {
  uint32_t w[6];
  w[0] = flags & 0xFFu;
  w[1] = w[2] = w[3] = w[4] = w[5] = 0;
  for (int i = 0; i < DRG_MAX_PATH; i++) w[(i + 1) >> 2] |= (uint32_t)path[i] << (8 * ((i + 1) & 3));
  w[5] |= ((uint32_t)round & 0xFFu) << 24;
  uint4* d = reinterpret_cast<uint4*>(&r);
  d[0] = make_uint4((uint32_t)capt, (uint32_t)(capt >> 32), (uint32_t)forb, (uint32_t)(forb >> 32));
  d[1] = make_uint4(item, ((uint32_t)level & 0xFFu) | ((dirs & 0xFFu) << 8) | (((uint32_t)landcur & 0xFFu) << 16) | (((uint32_t)fvictim & 0xFFu) << 24),
                    w[0], w[1]);
  d[2] = make_uint4(w[2], w[3], w[4], w[5]);
}

template <class R>
struct Walker {
  using BB = typename Pos<R>::BB;
  using G = Geo<R::S>;
  Pos<R> p;
  BB enemy0, static_occ, capt, forb, allp;
  uint32_t vd, man_dirs;
  int level, level0, cur, landcur, fvictim, value;
  bool king, has_child, root_king;

  __device__ __forceinline__ void bind(const Pos<R>& pos) {
    p = pos;
    enemy0 = pos.opp_men() | pos.opp_kings();
    man_dirs = R::kMenBackward ? (R::kOrtho ? 0xFFu : 0x0Fu) : (pos.turn ? 0x0Cu : 0x03u);  // american.py:179
  }
  __device__ __forceinline__ BB promo_rank() const { return p.turn ? G::ROW_LAST : G::ROW_FIRST; }
  __device__ __forceinline__ void set_piece(int origin, bool rk) {
    const BB obit = bb_bit<BB>(origin);
    root_king = rk;
    static_occ = rk ? BB(p.own_men() | (p.own_kings() & ~obit) | enemy0) : BB((p.own_men() & ~obit) | p.own_kings() | enemy0);
  }
  __device__ __forceinline__ uint32_t probe(const Tables<R::S>& T) const {
    return (king && R::kFlying) ? fly_dirs<R>(T, cur, capt, forb, enemy0, allp)
                                : jump_dirs<R>(T, cur, capt, enemy0, allp, king ? 0x0Fu : man_dirs);
  }
  // the search of the piece on `origin`, from its root
  __device__ __forceinline__ void begin_piece(int origin, bool rk, const Tables<R::S>& T) {
    set_piece(origin, rk);
    level = level0 = 0;
    cur = origin;
    king = rk;
    has_child = false;
    landcur = fvictim = kNone;
    capt = forb = 0;
    value = 0;
    allp = BB(static_occ | bb_bit<BB>(origin));
    vd = probe(T);
  }
  // continue where a record says; the squares of the path go into the stack (records read them from there)
  __device__ __forceinline__ void resume(const WalkRec& r, Frame* stack, int stride, const Tables<R::S>& T) {
    const bool rk = (r.flags & WF_ROOT_KING) != 0;
    if (r.flags & WF_FRESH) { begin_piece(r.path[0], rk, T); return; }
    set_piece(r.path[0], rk);
    level = level0 = r.level;
    for (int l = 0; l < level; l++) {
      Frame f;
      f.sq = r.path[l];
      f.dirflag = 0; f.landcur = (uint8_t)kNone; f.victim = 0;
      stack[l * stride] = f;
    }
    cur = r.path[level];
    king = (r.flags & WF_KING) != 0;
    has_child = (r.flags & WF_HAS_CHILD) != 0;
    landcur = r.landcur;
    fvictim = r.fvictim;
    capt = BB(r.capt);
    forb = BB(r.forb);
    value = 0;
    if (R::kFilter == FILTER_VALUE) {
      const BB kv = king_victim_squares_of(p);
      value = 100 * bb_popc(BB(capt & ~kv)) + 199 * bb_popc(BB(capt & kv));
    }
    allp = BB((static_occ & ~capt) | bb_bit<BB>(cur));
    vd = (r.flags & WF_PROBE) ? probe(T) : (uint32_t)r.dirflag;  // a child handed over on its own: nothing probed yet
  }
  // cap_piece priority bm, bk, wm, wk (frisian.py:392-393): squares whose occupant counts as a king victim
  __device__ static __forceinline__ BB king_victim_squares_of(const Pos<R>& q) { return BB(~q.bm & (q.bk | ~q.wm)); }

  // the next unvisited child of the current node (its victim and landing square); the node's cursor moves past it
  __device__ __forceinline__ bool next_child(const Tables<R::S>& T, int& victim, int& land) {
    if (!(king && R::kFlying)) {
      // man (any variant) or short king (american.py:210-243): jump over an adjacent enemy
      if (!vd) return false;
      const int d = __ffs((int)vd) - 1;
      vd &= vd - 1;
      victim = T.nb[cur * 8 + d];
      land = T.jmp[cur * 8 + d];
      return true;
    }
    // flying king: the first piece on the ray is an uncaptured enemy (probe() checked it); every free square behind
    // it, up to the next blocker, is a landing (standard.py:228-254)
    while (vd) {
      const int d = __ffs((int)vd) - 1;
      if (landcur == kNone) {
        victim = nearest_on_ray(BB((allp | forb) & T.ray[cur * 8 + d]), d);
        land = T.nb[victim * 8 + d];
      } else {
        victim = fvictim;
        land = T.nb[landcur * 8 + d];
        if (land == kNone || ((forb | allp) & bb_bit<BB>(land))) { landcur = kNone; vd &= vd - 1; continue; }
      }
      landcur = land;
      fvictim = victim;
      return true;
    }
    return false;
  }
  // mover status after a jump that lands on `land` (mid-capture promotion, russian.py:262-271)
  __device__ __forceinline__ bool king_after(int land) const {
    return king || (R::kMidPromo && (bb_bit<BB>(land) & promo_rank()) != 0);
  }

  // one iteration of capture_dfs's loop; true when the subtree of the resume node is finished
  template <int PASS, class Sink>
  __device__ __forceinline__ bool step(CapStat& st, Sink& sink, Frame* stack, int stride, const Tables<R::S>& T,
                                       uint32_t& overflow, int& descents) {
    int victim = kNone, land = kNone;
    const bool found = next_child(T, victim, land);
    if (found && level >= kMaxLevels) {  // path would not fit a record: report, do not descend
      overflow = 1;
      return false;
    }
    if (found) {
      Frame f;
      f.sq = (uint8_t)cur;
      f.dirflag = (uint8_t)vd;
      f.landcur = (uint8_t)landcur;
      f.victim = (uint8_t)(victim | (king ? 0x80 : 0));
      stack[level * stride] = f;
      const BB vb = bb_bit<BB>(victim);
      capt |= vb;
      if (king && R::kFlying) forb |= vb;
      if (R::kFilter == FILTER_VALUE) value += (king_victim_squares_of(p) & vb) ? 199 : 100;
      king = king_after(land);
      level++;
      descents++;
      DRG_TRACE_DESCENT();
      cur = land;
      landcur = kNone;
      has_child = false;
      allp = BB((static_occ & ~capt) | bb_bit<BB>(cur));
      vd = probe(T);
      return false;
    }
    if (!has_child && level > 0) {  // a leaf
      const int metric = R::kFilter == FILTER_VALUE ? value : level;
      if (PASS == 0 || PASS == 2) {
        st.n_all++;
        if (metric > st.best) {
          st.best = metric;
          st.n_best_man = 0;
          st.n_best_king = 0;
          if constexpr (PASS == 2) {
            if (R::kFilter != FILTER_NONE) sink.restart();
          }
        }
        if (metric == st.best) { if (root_king) st.n_best_king++; else st.n_best_man++; }
        if constexpr (PASS == 2) {
          if (R::kFilter == FILTER_NONE || metric == st.best)
            sink.capture(level + 1, capt, king && !root_king, root_king, cur, stack, stride);
        }
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
    if (level == level0) return true;
    level--;
    const Frame f = stack[level * stride];
    cur = f.sq;
    vd = f.dirflag;
    has_child = true;
    king = (f.victim & 0x80) != 0;
    landcur = f.landcur;
    fvictim = f.victim & 0x7F;
    const BB vb = bb_bit<BB>(fvictim);
    capt &= ~vb;
    if (king && R::kFlying) forb &= ~vb;
    if (R::kFilter == FILTER_VALUE) value -= (king_victim_squares_of(p) & vb) ? 199 : 100;
    allp = BB((static_occ & ~capt) | bb_bit<BB>(cur));
    return false;
  }

  __device__ __forceinline__ void fill_rec(WalkRec& r, uint32_t item, int l, int sq, uint32_t dirs, int lc, int fv, uint32_t fl,
                                           BB c, BB fb, const Frame* stack, int stride, int round) const {
    uint8_t path[DRG_MAX_PATH];
    for (int i = 0; i < DRG_MAX_PATH; i++) path[i] = i < l ? stack[i * stride].sq : (i == l ? (uint8_t)sq : (uint8_t)0);
    store_rec(r, (uint64_t)c, (uint64_t)fb, item, l, dirs, lc, fv, (root_king ? WF_ROOT_KING : 0u) | fl, path, round);
  }
  // Everything this walker has not explored, as records in canonical order: the open frames of its stack, deepest
  // first (the current node included).  The `explode` shallowest open frames -- whose remaining children root the
  // largest unexplored subtrees -- are handed over child by child, so that the heavy part of a tree fans out by the
  // branching factor per round instead of by one record per level.  out == nullptr: only count.
  __device__ __forceinline__ int handover(const Frame* stack, int stride, uint32_t item, WalkRec* out, int explode,
                                          const Tables<R::S>& T, int round) const {
    int open = vd ? 1 : 0;
    for (int l = level0; l < level; l++) open += stack[l * stride].dirflag ? 1 : 0;
    int n = 0, seen = 0;
    BB c = capt, fb = forb;
    for (int l = level; l >= level0; l--) {
      int sq, lc, fv;
      uint32_t dirs;
      bool kl, hc;
      if (l == level) {
        sq = cur; dirs = vd; lc = landcur; fv = fvictim; kl = king; hc = has_child;
      } else {
        const Frame f = stack[l * stride];
        sq = f.sq; dirs = f.dirflag; lc = f.landcur; fv = f.victim & 0x7F; kl = (f.victim & 0x80) != 0; hc = true;
        const BB vb = bb_bit<BB>(fv);  // state of level l: without the victim of the child that was being explored
        c &= ~vb;
        if (kl && R::kFlying) fb &= ~vb;
      }
      if (!dirs) continue;
      seen++;
      if (seen <= open - explode || l >= kMaxLevels) {  // one record for the frame
        if (out) fill_rec(out[n], item, l, sq, dirs, lc, fv, (kl ? WF_KING : 0u) | (hc ? WF_HAS_CHILD : 0u), c, fb, stack, stride, round);
        n++;
        continue;
      }
      // one record per remaining child: a scratch walker standing on the frame enumerates them
      Walker x = *this;
      x.level = l; x.cur = sq; x.vd = dirs; x.landcur = lc; x.fvictim = fv; x.king = kl; x.capt = c; x.forb = fb;
      x.allp = BB((static_occ & ~c) | bb_bit<BB>(sq));
      int victim = kNone, land = kNone;
      while (x.next_child(T, victim, land)) {
        if (out) {
          // the child as a node of its own: nothing probed yet; below l the stack holds the path (level l itself only
          // while l < level, so its square is passed along)
          const BB vb = bb_bit<BB>(victim);
          const bool k2 = x.king_after(land);
          uint8_t path[DRG_MAX_PATH];
          for (int i = 0; i < DRG_MAX_PATH; i++)
            path[i] = i < l ? stack[i * stride].sq : (i == l ? (uint8_t)sq : (i == l + 1 ? (uint8_t)land : (uint8_t)0));
          store_rec(out[n], (uint64_t)BB(c | vb), (uint64_t)((kl && R::kFlying) ? BB(fb | vb) : fb), item, l + 1, 0u, kNone, kNone,
                    (root_king ? WF_ROOT_KING : 0u) | (k2 ? WF_KING : 0u) | WF_PROBE, path, round);
        }
        n++;
      }
    }
    return n;
  }
};

// a piece that has not been started, as a record
__device__ __forceinline__ void fresh_rec(WalkRec& r, uint32_t item, int origin, bool root_king, int round) {
  uint8_t path[DRG_MAX_PATH];
  for (int i = 0; i < DRG_MAX_PATH; i++) path[i] = i == 0 ? (uint8_t)origin : (uint8_t)0;
  store_rec(r, 0ull, 0ull, item, 0, 0u, kNone, kNone, WF_FRESH | (root_king ? (WF_ROOT_KING | WF_KING) : 0u), path, round);
}

// result of one walk (a board's root walk, or a handed-over record)
struct __align__(16) WalkRes {
  uint32_t info;    // bits 0-15 best metric of the walk's own leaves, bit 16 the walk split (its records exist),
                    // bits 24-31 number of records it handed over
  uint32_t n_a;     // FILTER_NONE: all own leaves; otherwise own leaves of the best metric started by a man
  uint32_t n_b;     // own leaves of the best metric started by a king
  uint32_t child0;  // index of its first record (they are consecutive, in canonical order)
};
constexpr uint32_t kWalkSplit = 1u << 16;
constexpr int kNoBudget = 0x7FFFFFFF;
constexpr int kWalkExplode = 1;
enum { kWalkRunning = 0, kWalkDone = 1, kWalkWantsSplit = 2 };

// what a board keeps of all its walks: max over the walks of (best metric << 1 | a king-started leaf attains it)
template <class R>
__device__ __forceinline__ uint32_t walk_board_word(const CapStat& st) {
  return ((uint32_t)st.best << 1) | (st.n_best_king ? 1u : 0u);
}
// leaves of a walk that are legal moves, given the board's word (the variant's filter, frisian.py:264-268)
template <class R>
__device__ __forceinline__ uint32_t walk_kept(const WalkRes& r, uint32_t board_word) {
  if (R::kFilter == FILTER_NONE) return r.n_a;
  if ((r.info & 0xFFFFu) != (board_word >> 1)) return 0u;
  if (R::kFilter == FILTER_MAXLEN) return r.n_a + r.n_b;
  return (board_word & 1u) ? r.n_b : r.n_a;
}

// One lane's walk: a board's root walk (all its capturing pieces, men first) or one record.  advance() is one
// iteration; the kernels run it in a loop in which idle lanes fetch the next walk, the host replay (tools/devsim)
// runs it to completion.  Alloc::reserve(n) returns the index of n consecutive free records or -1 (then the walk
// simply goes on without a budget).
template <class R, int PASS>
struct WalkLane {
  using BB = typename Pos<R>::BB;
  Walker<R> w;
  CapStat st;
  BB men, kings;       // pieces of a root walk that have not been started
  int descents, budget;
  int explode;         // shallowest open frames handed over child by child (Walker::handover)
  int next_round;      // round in which this walk's records will be walked (stored in them: it decides their budget)
  uint32_t item;
  bool in_piece;
  // outcome
  bool split;
  uint32_t child0, nchild;

  __device__ __forceinline__ void reset(uint32_t item_, int budget_) {
    item = item_;
    budget = budget_;
    descents = 0;
    split = false;
    child0 = nchild = 0;
    explode = kWalkExplode;
    next_round = 1;
    st.reset();  // an emitting walk (PASS 1) gets the board's best metric / king preference from its caller afterwards
  }
  __device__ __forceinline__ void start_root(const Pos<R>& p, const Capturers<BB>& c, uint32_t item_, int budget_,
                                             const Tables<R::S>& T) {
    reset(item_, budget_);
    w.bind(p);
    men = c.men;
    kings = c.kings;
    in_piece = false;
    next_piece(T);
  }
  __device__ __forceinline__ void start_rec(const Pos<R>& p, const WalkRec& r, int budget_, Frame* stack, int stride,
                                            const Tables<R::S>& T) {
    reset(r.item, budget_);
    w.bind(p);
    men = kings = 0;
    w.resume(r, stack, stride, T);
    in_piece = true;
  }
  __device__ __forceinline__ void next_piece(const Tables<R::S>& T) {
    const bool rk = men == 0;
    const BB src = rk ? kings : men;
    const int sq = bb_ctz(src);
    if (rk) kings &= kings - 1;
    else men &= men - 1;
    w.begin_piece(sq, rk, T);
    in_piece = true;
  }
  // hand everything unexplored to the next round; false when the pool is full (the walk then continues unbudgeted)
  template <class Alloc>
  __device__ __forceinline__ bool hand_over(const Frame* stack, int stride, Alloc& alloc, const Tables<R::S>& T) {
    if (PASS == 1) { split = true; return true; }  // the emitting re-walk stops where the counting walk stopped
    const int nf = in_piece ? w.handover(stack, stride, item, nullptr, explode, T, next_round) : 0;
    const int np = bb_popc(men) + bb_popc(kings);
    const long long at = alloc.reserve(nf + np);
    if (at < 0) { budget = kNoBudget; return false; }
    WalkRec* out = alloc.at(at);
    if (nf) w.handover(stack, stride, item, out, explode, T, next_round);
    int n = nf;
    BB m = men, k = kings;
    while (m | k) {
      const bool rk = m == 0;
      const BB src = rk ? k : m;
      const int sq = bb_ctz(src);
      if (rk) k &= k - 1;
      else m &= m - 1;
      fresh_rec(out[n++], item, sq, rk, next_round);
    }
    split = true;
    child0 = (uint32_t)at;
    nchild = (uint32_t)n;
    return true;
  }
  // Round 0 only: give the whole board up.  Every capturing piece becomes a fresh record of round 1 and what this
  // walk found so far is dropped.  It keeps the hand-over code out of the kernel that walks EVERY tree-search board
  // (with it the kernel was 1.5x slower on ordinary batches); a board that exceeds the large round-0 budget is a
  // monster and loses a few dozen descents out of thousands.  false: the pool is full, the walk goes on unbudgeted.
  template <class Alloc>
  __device__ __forceinline__ bool abandon(const Capturers<BB>& c, Alloc& alloc) {
    const int np = bb_popc(c.men) + bb_popc(c.kings);
    const long long at = alloc.reserve(np);
    if (at < 0) { budget = kNoBudget; return false; }
    WalkRec* out = alloc.at(at);
    int n = 0;
    BB m = c.men, k = c.kings;
    while (m | k) {
      const bool rk = m == 0;
      const BB src = rk ? k : m;
      const int sq = bb_ctz(src);
      if (rk) k &= k - 1;
      else m &= m - 1;
      fresh_rec(out[n++], item, sq, rk, next_round);
    }
    split = true;
    child0 = (uint32_t)at;
    nchild = (uint32_t)n;
    st.reset();
    return true;
  }
  // one iteration: kWalkRunning, kWalkDone, or kWalkWantsSplit -- the budget is spent and something is unexplored: the
  // caller calls hand_over() (the kernels wait until several lanes of the warp want to, so that they do it together),
  // which completes the walk, or lifts its budget when the pool is full
  template <class Sink>
  __device__ __forceinline__ int advance(Sink& sink, Frame* stack, int stride, const Tables<R::S>& T, uint32_t& overflow) {
    if (in_piece) {
      if (descents >= budget && w.vd) return kWalkWantsSplit;
      if (!w.template step<PASS>(st, sink, stack, stride, T, overflow, descents)) return kWalkRunning;
      in_piece = false;
    }
    if (!(men | kings)) return kWalkDone;
    if (descents >= budget) return kWalkWantsSplit;
    next_piece(T);
    return kWalkRunning;
  }
  // a whole walk (emitting walks, host replay)
  template <class Sink, class Alloc>
  __device__ __forceinline__ void run(Sink& sink, Frame* stack, int stride, const Tables<R::S>& T, uint32_t& overflow,
                                      Alloc& alloc) {
    for (;;) {
      const int s = advance(sink, stack, stride, T, overflow);
      if (s == kWalkDone || (s == kWalkWantsSplit && hand_over(stack, stride, alloc, T))) return;
    }
  }
  __device__ __forceinline__ WalkRes result() const {
    WalkRes r;
    r.info = ((uint32_t)st.best & 0xFFFFu) | (split ? kWalkSplit : 0u) | (nchild << 24);
    r.n_a = R::kFilter == FILTER_NONE ? (uint32_t)st.n_all : (uint32_t)st.n_best_man;
    r.n_b = (uint32_t)st.n_best_king;
    r.child0 = child0;
    return r;
  }
};

// pieces of `movers` that can jump in direction D (bit-parallel root test of standard.py:188-197)
template <class R, int D>
__device__ __forceinline__ typename Pos<R>::BB jumpers_dir(typename Pos<R>::BB movers, typename Pos<R>::BB enemy,
                                                           typename Pos<R>::BB empty) {
  using BB = typename Pos<R>::BB;
  const BB over = step<R::S, D>(movers) & enemy;
  const BB land = step<R::S, D>(over) & empty;
  return step<R::S, opposite(D)>(step<R::S, opposite(D)>(land));
}

// ---------------------------------------------------------------------------------------------
// single-jump fast path.  When no king can capture and no man can jump a second time, every capture
// leaf is one jump of one man; the whole result is then computed bit-parallel, without the DFS.
// A second jump from the landing square l = origin + 2*d1 in direction d2 != opposite(d1) never touches
// the squares the first jump changed (origin, first victim, l), so its existence is decided on the
// ORIGINAL enemy / empty sets; d2 == opposite(d1) would re-capture the first victim and is excluded.
// ---------------------------------------------------------------------------------------------
template <class R>
struct MenJumps {
  using BB = typename Pos<R>::BB;
  BB j[R::kOrtho ? 8 : 4];  // men that can jump in direction d (0 for directions the men may not use)
};

template <class R, int D>
__device__ __forceinline__ bool man_dir_allowed(int turn) {
  if (D >= 4) return R::kOrtho;
  return R::kMenBackward || (D < 2 ? turn == 0 : turn == 1);  // american.py:179
}

template <class R>
__device__ __forceinline__ MenJumps<R> men_jumps(const Pos<R>& p) {
  using BB = typename Pos<R>::BB;
  const BB men = p.own_men(), enemy = p.opp_men() | p.opp_kings();
  const BB empty = BB(~p.occ() & Geo<R::S>::MASK);
  MenJumps<R> m;
  m.j[0] = man_dir_allowed<R, 0>(p.turn) ? BB(jumpers_dir<R, 0>(men, enemy, empty) & men) : BB(0);
  m.j[1] = man_dir_allowed<R, 1>(p.turn) ? BB(jumpers_dir<R, 1>(men, enemy, empty) & men) : BB(0);
  m.j[2] = man_dir_allowed<R, 2>(p.turn) ? BB(jumpers_dir<R, 2>(men, enemy, empty) & men) : BB(0);
  m.j[3] = man_dir_allowed<R, 3>(p.turn) ? BB(jumpers_dir<R, 3>(men, enemy, empty) & men) : BB(0);
  if (R::kOrtho) {
    m.j[4] = BB(jumpers_dir<R, 4>(men, enemy, empty) & men);
    m.j[5] = BB(jumpers_dir<R, 5>(men, enemy, empty) & men);
    m.j[6] = BB(jumpers_dir<R, 6>(men, enemy, empty) & men);
    m.j[7] = BB(jumpers_dir<R, 7>(men, enemy, empty) & men);
  }
  return m;
}

template <class R, int D1, int D2>
__device__ __forceinline__ bool second_jump(const Pos<R>& p, typename Pos<R>::BB land, typename Pos<R>::BB enemy,
                                            typename Pos<R>::BB empty) {
  if (D2 == opposite(D1) || !man_dir_allowed<R, D2>(p.turn)) return false;
  return jumpers_dir<R, D2>(land, enemy, empty) != 0;
}

template <class R, int D1>
__device__ __forceinline__ bool continues_after(const Pos<R>& p, typename Pos<R>::BB jumpers,
                                                typename Pos<R>::BB enemy, typename Pos<R>::BB empty) {
  using BB = typename Pos<R>::BB;
  if (!jumpers) return false;
  const BB land = step<R::S, D1>(step<R::S, D1>(jumpers));
  if (R::kMidPromo && (land & (p.turn ? Geo<R::S>::ROW_LAST : Geo<R::S>::ROW_FIRST))) return true;  // goes on as a king
  bool c = second_jump<R, D1, 0>(p, land, enemy, empty) || second_jump<R, D1, 1>(p, land, enemy, empty) ||
           second_jump<R, D1, 2>(p, land, enemy, empty) || second_jump<R, D1, 3>(p, land, enemy, empty);
  if (R::kOrtho)
    c = c || second_jump<R, D1, 4>(p, land, enemy, empty) || second_jump<R, D1, 5>(p, land, enemy, empty) ||
        second_jump<R, D1, 6>(p, land, enemy, empty) || second_jump<R, D1, 7>(p, land, enemy, empty);
  return c;
}

// true when every capture of this board is a single man jump (requires: no king can capture)
template <class R>
__device__ __forceinline__ bool only_single_jumps(const Pos<R>& p, const MenJumps<R>& m) {
  using BB = typename Pos<R>::BB;
  const BB enemy = p.opp_men() | p.opp_kings();
  const BB empty = BB(~p.occ() & Geo<R::S>::MASK);
  bool deeper = continues_after<R, 0>(p, m.j[0], enemy, empty) || continues_after<R, 1>(p, m.j[1], enemy, empty) ||
                continues_after<R, 2>(p, m.j[2], enemy, empty) || continues_after<R, 3>(p, m.j[3], enemy, empty);
  if (R::kOrtho)
    deeper = deeper || continues_after<R, 4>(p, m.j[4], enemy, empty) ||
             continues_after<R, 5>(p, m.j[5], enemy, empty) || continues_after<R, 6>(p, m.j[6], enemy, empty) ||
             continues_after<R, 7>(p, m.j[7], enemy, empty);
  return !deeper;
}

// squares whose occupant counts as a KING victim (cap_piece priority bm, bk, wm, wk; frisian.py:392-393)
template <class R>
__device__ __forceinline__ typename Pos<R>::BB king_victim_squares(const Pos<R>& p) {
  using BB = typename Pos<R>::BB;
  return BB(~p.bm & (p.bk | ~p.wm));
}

// legal single-jump captures: count, and (EMIT) hand them to the sink in canonical order
template <class R, bool EMIT, class Sink>
__device__ __forceinline__ int single_jumps(const Pos<R>& p, const MenJumps<R>& m, Sink& sink, const Tables<R::S>& T) {
  using BB = typename Pos<R>::BB;
  constexpr int ND = R::kOrtho ? 8 : 4;
  BB keep[ND];
  bool any_king_victim = false;
  if (R::kFilter == FILTER_VALUE) {  // frisian.py:259-262: only the most valuable captures (a king is worth 199, a man 100)
    const BB kv = king_victim_squares<R>(p);
    BB hit[ND];
    hit[0] = step<R::S, opposite(0)>(step<R::S, 0>(m.j[0]) & kv);
    hit[1] = step<R::S, opposite(1)>(step<R::S, 1>(m.j[1]) & kv);
    hit[2] = step<R::S, opposite(2)>(step<R::S, 2>(m.j[2]) & kv);
    hit[3] = step<R::S, opposite(3)>(step<R::S, 3>(m.j[3]) & kv);
    if (ND == 8) {
      hit[4] = step<R::S, opposite(4)>(step<R::S, 4>(m.j[4]) & kv);
      hit[5] = step<R::S, opposite(5)>(step<R::S, 5>(m.j[5]) & kv);
      hit[6] = step<R::S, opposite(6)>(step<R::S, 6>(m.j[6]) & kv);
      hit[7] = step<R::S, opposite(7)>(step<R::S, 7>(m.j[7]) & kv);
    }
#pragma unroll
    for (int d = 0; d < ND; d++) any_king_victim = any_king_victim || hit[d] != 0;
#pragma unroll
    for (int d = 0; d < ND; d++) keep[d] = any_king_victim ? hit[d] : m.j[d];
  } else {
#pragma unroll
    for (int d = 0; d < ND; d++) keep[d] = m.j[d];
  }
  int n = 0;
  BB all = 0;
#pragma unroll
  for (int d = 0; d < ND; d++) {
    n += bb_popc(keep[d]);
    all |= keep[d];
  }
  if (EMIT) {
    while (all) {  // men ascending, directions ascending (DFS pre-order of standard.py:177-181,188)
      const int sq = bb_ctz(all);
      const BB b = bb_bit<BB>(sq);
      all &= all - 1;
#pragma unroll
      for (int d = 0; d < ND; d++)
        if (keep[d] & b) sink.jump(sq, T.nb[sq * 8 + d], T.jmp[sq * 8 + d]);
    }
  }
  return n;
}

// the men and kings that can make a first jump -- exactly the pieces whose DFS yields >= 1 leaf;
// mj receives the per-direction man jumpers
template <class R>
__device__ __forceinline__ Capturers<typename Pos<R>::BB> find_capturers(const Pos<R>& p, const Tables<R::S>& T,
                                                                         MenJumps<R>& mj) {
  using BB = typename Pos<R>::BB;
  Capturers<BB> c;
  c.men = c.kings = 0;
  const BB enemy = p.opp_men() | p.opp_kings();
#pragma unroll
  for (int d = 0; d < (R::kOrtho ? 8 : 4); d++) mj.j[d] = 0;
  if (!enemy) return c;  // standard.py:167-169
  mj = men_jumps<R>(p);
#pragma unroll
  for (int d = 0; d < (R::kOrtho ? 8 : 4); d++) c.men |= mj.j[d];
  const BB occ = p.occ();
  const BB empty = BB(~occ & Geo<R::S>::MASK);
  const BB kings = p.own_kings();
  if (!kings) return c;
  if (!R::kFlying) {  // short kings jump like men, in all four directions (american.py:214-219)
    BB k = jumpers_dir<R, 0>(kings, enemy, empty) | jumpers_dir<R, 1>(kings, enemy, empty) |
           jumpers_dir<R, 2>(kings, enemy, empty) | jumpers_dir<R, 3>(kings, enemy, empty);
    c.kings = k & kings;
    return c;
  }
  BB bb = kings;
  while (bb) {
    const int sq = bb_ctz(bb);
    bb &= bb - 1;
    bool can = false;
#pragma unroll
    for (int d = 0; d < (R::kOrtho ? 8 : 4); d++) {
      const BB blk = BB(occ & T.ray[sq * 8 + d]);
      if (blk) {
        const int t = nearest_on_ray(blk, d);
        if (enemy & bb_bit<BB>(t)) {
          const int land = T.nb[t * 8 + d];
          if (land != kNone && !(occ & bb_bit<BB>(land))) can = true;
        }
      }
    }
    if (can) c.kings |= bb_bit<BB>(sq);
  }
  return c;
}

// all captures of the side to move, men first then kings (standard.py:164-182)
template <class R, int PASS, class Sink>
__device__ __forceinline__ void gen_captures(const Pos<R>& p, const Capturers<typename Pos<R>::BB>& c, CapStat& st,
                                             Sink& sink, Frame* stack, int stride, const Tables<R::S>& T,
                                             uint32_t& overflow) {
  using BB = typename Pos<R>::BB;
  // one loop, one inlined copy of the DFS (the kernels are instruction-cache sensitive): men first, then kings
  BB men = c.men, kings = c.kings;
  while (men | kings) {
    const bool root_king = men == 0;
    const BB src = root_king ? kings : men;
    const int sq = bb_ctz(src);
    if (root_king) kings &= kings - 1;
    else men &= men - 1;
    capture_dfs<R, PASS>(p, sq, root_king, st, sink, stack, stride, T, overflow);
  }
}

// quiet moves (standard.py:107-162 / american.py:98-149); COUNT_ONLY returns the count without visiting
template <class R, bool COUNT_ONLY, class Sink>
__device__ __forceinline__ int gen_quiet(const Pos<R>& p, Sink& sink, const Tables<R::S>& T) {
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
  const BB occ = p.occ();
  while (kings) {
    const int sq = bb_ctz(kings);
    kings &= kings - 1;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      if (R::kFlying) {
        const BB ray = T.ray[sq * 8 + d];
        BB reach = ray_until_blocker<BB>(ray, BB(occ & ray), d);
        if (COUNT_ONLY) {
          n += bb_popc(reach);
        } else {
          while (reach) {  // near -> far
            const int to = dir_is_up(d) ? bb_msb(reach) : bb_ctz(reach);
            reach &= ~bb_bit<BB>(to);
            sink.quiet(sq, to, true);
            n++;
          }
        }
      } else {
        const int to = T.nb[sq * 8 + d];
        if (to != kNone && (empty & bb_bit<BB>(to))) {
          if (!COUNT_ONLY) sink.quiet(sq, to, true);
          n++;
        }
      }
    }
  }
  return n;
}

// A sink that ignores moves (statistics / counting passes)
struct NullSink {
  __device__ __forceinline__ void quiet(int, int, bool) {}
  __device__ __forceinline__ void jump(int, int, int) {}
  template <class BB>
  __device__ __forceinline__ void capture(int, BB, bool, bool, int, const Frame*, int) {}
};

// what the emitting pass needs to know from the statistics pass, in one word (handed from the count step to the
// emit step of the same batch so that the tree is walked twice, not three times): best metric | king-preference bit
__device__ __forceinline__ uint32_t pack_cap_stat(const CapStat& st) {
  return (uint32_t)st.best | (st.n_best_king ? 0x80000000u : 0u);
}
__device__ __forceinline__ void unpack_cap_stat(uint32_t w, CapStat& st) {
  st.reset();
  st.best = (int)(w & 0x7FFFFFFFu);
  st.n_best_king = (w >> 31) ? 1 : 0;
}

// number of capture moves that are legal (after the variant's filter); c.any() must hold
template <class R>
__device__ __forceinline__ int count_captures(const Pos<R>& p, const Capturers<typename Pos<R>::BB>& c, Frame* stack,
                                              int stride, const Tables<R::S>& T, uint32_t& overflow,
                                              uint32_t* stat_word = nullptr) {
  NullSink ns;
  CapStat st;
  st.reset();
  gen_captures<R, 0>(p, c, st, ns, stack, stride, T, overflow);
  if (stat_word) *stat_word = pack_cap_stat(st);
  if (R::kFilter == FILTER_NONE) return st.n_all;
  if (R::kFilter == FILTER_MAXLEN) return st.n_best_man + st.n_best_king;
  return st.n_best_king ? st.n_best_king : st.n_best_man;  // frisian.py:264-268
}

// Statistics word of a tree search whose leaves were kept (count_captures_keep): best metric in bits 0-19, number of
// kept records in bits 20-27, bit 30 = the kept records are complete (<= the pool's room), bit 31 = king preference.
constexpr int kKeepRecords = 16;  // pool records per board
__device__ __forceinline__ uint32_t pack_keep_stat(const CapStat& st, uint32_t kept, bool complete) {
  return ((uint32_t)st.best & 0xFFFFFu) | ((kept & 0xFFu) << 20) | (complete ? 0x40000000u : 0u) |
         (st.n_best_king ? 0x80000000u : 0u);
}
__device__ __forceinline__ void unpack_keep_stat(uint32_t w, CapStat& st) {
  st.reset();
  st.best = (int)(w & 0xFFFFFu);
  st.n_best_king = (w >> 31) ? 1 : 0;
}

// count_captures that also keeps the leaves of the best metric in `sink` (PASS 2); KeepSink: see drg_kernels.cuh
template <class R, class KeepSink>
__device__ __forceinline__ int count_captures_keep(const Pos<R>& p, const Capturers<typename Pos<R>::BB>& c,
                                                   KeepSink& sink, Frame* stack, int stride, const Tables<R::S>& T,
                                                   uint32_t& overflow, uint32_t* stat_word) {
  CapStat st;
  st.reset();
  gen_captures<R, 2>(p, c, st, sink, stack, stride, T, overflow);
  *stat_word = pack_keep_stat(st, sink.n, sink.n <= sink.room);
  if (R::kFilter == FILTER_NONE) return st.n_all;
  if (R::kFilter == FILTER_MAXLEN) return st.n_best_man + st.n_best_king;
  return st.n_best_king ? st.n_best_king : st.n_best_man;  // frisian.py:264-268
}

// the legal capture moves, canonical order, into `sink`; returns their number; c.any() must hold
template <class R, class Sink>
__device__ __forceinline__ int emit_captures(const Pos<R>& p, const Capturers<typename Pos<R>::BB>& c, Sink& sink,
                                             Frame* stack, int stride, const Tables<R::S>& T, uint32_t& overflow) {
  CapStat st;
  st.reset();
  if (R::kFilter != FILTER_NONE) gen_captures<R, 0>(p, c, st, sink, stack, stride, T, overflow);
  gen_captures<R, 1>(p, c, st, sink, stack, stride, T, overflow);
  return st.n_emit;
}

// same, with the statistics pass already done (stat_word from count_captures)
template <class R, class Sink>
__device__ __forceinline__ int emit_captures_with_stat(const Pos<R>& p, const Capturers<typename Pos<R>::BB>& c,
                                                       Sink& sink, Frame* stack, int stride, const Tables<R::S>& T,
                                                       uint32_t& overflow, uint32_t stat_word) {
  CapStat st;
  unpack_cap_stat(stat_word, st);
  gen_captures<R, 1>(p, c, st, sink, stack, stride, T, overflow);
  return st.n_emit;
}

// legal capture moves of a board with c.any(): number, and (EMIT) the moves into the sink
// stat: the counting call stores the statistics word of a tree search there (kNoStat for the bit-parallel path),
// the emitting call reads it and skips its own statistics pass
constexpr uint32_t kNoStat = 0xFFFFFFFFu;
template <class R, bool EMIT, class Sink>
__device__ __forceinline__ int captures_of(const Pos<R>& p, const Capturers<typename Pos<R>::BB>& c,
                                           const MenJumps<R>& mj, Sink& sink, Frame* stack, int stride,
                                           const Tables<R::S>& T, uint32_t& overflow, uint32_t* stat) {
  if (!c.kings && only_single_jumps<R>(p, mj)) {
    if (!EMIT) *stat = kNoStat;
    return single_jumps<R, EMIT>(p, mj, sink, T);
  }
  if (EMIT) return emit_captures_with_stat<R>(p, c, sink, stack, stride, T, overflow, *stat);
  return count_captures<R>(p, c, stack, stride, T, overflow, stat);
}

// len(board.legal_moves) -- single-thread form (rollouts); the batched kernels split it into the
// quiet part and the compacted capture parts
template <class R>
__device__ __forceinline__ int count_moves(const Pos<R>& p, Frame* stack, int stride, const Tables<R::S>& T,
                                           uint32_t& overflow, uint32_t& stat) {
  NullSink ns;
  MenJumps<R> mj;
  const auto c = find_capturers<R>(p, T, mj);
  int n = 0;
  stat = kNoStat;
  if (c.any()) n = captures_of<R, false>(p, c, mj, ns, stack, stride, T, overflow, &stat);
  if (R::kOptionalCapture || !c.any()) n += gen_quiet<R, true>(p, ns, T);  // american.py:96
  return n;
}

// board.legal_moves in canonical order into `sink`; returns the number of moves.  `stat` comes from the
// count_moves call on the same board (the statistics pass of a tree search is not repeated).
template <class R, class Sink>
__device__ __forceinline__ int generate_moves(const Pos<R>& p, Sink& sink, Frame* stack, int stride,
                                              const Tables<R::S>& T, uint32_t& overflow, uint32_t stat) {
  MenJumps<R> mj;
  const auto c = find_capturers<R>(p, T, mj);
  int n = 0;
  if (R::kOptionalCapture || !c.any()) n += gen_quiet<R, false>(p, sink, T);
  if (c.any()) n += captures_of<R, true>(p, c, mj, sink, stack, stride, T, overflow, &stat);
  return n;
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
