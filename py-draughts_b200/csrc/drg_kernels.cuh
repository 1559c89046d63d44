// drg_kernels.cuh -- batched kernels over structure-of-arrays boards (sm_100a).
//
// Layout in HBM: wm/wk/bm/bk uint64[n], turn uint8[n], clock uint8[n]; one thread owns one board,
// so a warp's loads of each array are one contiguous 256-byte (uint64) segment.
//
// Every batched kernel has the same structure (128 boards per block):
//   phase A  thread = its own board: load, bit-parallel quiet moves + exact capture pre-check
//            (find_capturers).  Boards without captures are finished here.  Boards WITH captures are
//            appended to a list in shared memory.
//   phase B  the capture boards are re-dealt onto consecutive threads; boards whose captures are single man
//            jumps are finished bit-parallel (single_jumps).  Boards that need the tree search are appended
//            to a GRID-wide list (one atomicAdd per converged lane group) ...
//   deep     ... which follow-up kernels (k_walk_root + k_walk_rounds / k_emit_deep + k_emit_recs / k_perft_count_deep /
//            k_expand_deep) drain with one board per
//            thread on dense warps and the explicit DFS stack in shared memory.
//   wide outputs (k_emit): every block keeps one BIT per mask byte of its 128 rows in shared memory; move emitters
//            set bits; at the end each warp widens its bit string to bytes and streams the rows with coalesced
//            16-byte stores (no partial-sector byte store ever reaches HBM); tensor planes likewise as float4s.
//   fused call (drg_movegen with a mask, drg_abi.cu SplitMovegenFn): k_emit<rows only> starts at once on the caller's
//            stream; k_count -> k_walk_root -> k_walk_rounds (+ offsets scan) -> k_emit<records only> -> k_emit_deep -> k_emit_recs run beside it on a
//            high-priority stream; mask bytes of tree-search boards go through a list (MaskFix) that k_mask_fixup
//            stores after the join.
//   small batches (n <= 128): k_small, one block, one launch, mapped pinned memory in and out.
#pragma once
#include <cuda_runtime.h>

#include <stdint.h>

#include "drg_core.cuh"
#include "drg_records.cuh"

namespace drg {

constexpr int kBlock = 128;      // threads per block = boards per block
constexpr int kWarps = kBlock / 32;

// geometry tables in global memory (filled by drg_init), copied to shared memory by every block
__device__ Tables<50> g_tab50;
__device__ Tables<32> g_tab32;

template <int S>
__device__ __forceinline__ void load_tables(Tables<S>* dst) {
  static_assert(sizeof(Tables<S>) % 16 == 0, "tables are copied as uint4");
  const uint4* src = reinterpret_cast<const uint4*>(S == 50 ? (const void*)&g_tab50 : (const void*)&g_tab32);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(Tables<S>) / 16); i += blockDim.x) d[i] = src[i];
}

template <class R>
__device__ __forceinline__ Pos<R> load_pos(const uint64_t* __restrict__ wm, const uint64_t* __restrict__ wk,
                                           const uint64_t* __restrict__ bm, const uint64_t* __restrict__ bk,
                                           int turn, int64_t i) {
  using BB = typename Pos<R>::BB;
  Pos<R> p;
  p.wm = BB(wm[i]) & Geo<R::S>::MASK;
  p.wk = BB(wk[i]) & Geo<R::S>::MASK;
  p.bm = BB(bm[i]) & Geo<R::S>::MASK;
  p.bk = BB(bk[i]) & Geo<R::S>::MASK;
  p.turn = turn;
  return p;
}

// shared state of the phase hand-overs: boards with captures (A -> B), boards that need the tree search (B -> C)
struct CapList {
  int n, n_deep;
  uint8_t item[kBlock];
  uint8_t deep[kBlock];
};

// grid-wide list of boards that need the capture tree search (filled by the count kernels, drained by the walk kernels)
struct DeepList {
  uint32_t* items;  // board indices
  unsigned* n;      // number of items (device counter, zeroed before the launch)
};

// append the calling lanes' items to a grid-wide list: one atomicAdd per converged lane group
__device__ __forceinline__ void list_append(DeepList list, uint32_t item) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(list.n, (unsigned)__popc(m));
  base = __shfl_sync(m, base, leader);
  list.items[base + __popc(m & ((1u << lane) - 1))] = item;
}

// ---------------------------------------------------------------------------------------------
// sinks
// ---------------------------------------------------------------------------------------------
template <int S>
struct EmitSink {  // writes records (at most `room`) and sets the board's mask bytes
  DrgMove* out;
  uint8_t* mrow;  // this board's legal_moves_mask row (base.py:833-837) or NULL
  uint32_t n, room;
  __device__ __forceinline__ void quiet(int from, int to, bool king) {
    if (n < room) { Rec r; rec_quiet(r, from, to, king); rec_store(out + n, r); }
    if (mrow) mrow[from * S + to] = 1;
    n++;
  }
  template <class BB>
  __device__ __forceinline__ void capture(int nsq, BB capt, bool promo, bool root_king, int cur, const Frame* stack,
                                          int stride) {
    if (n < room) { Rec r; rec_capture(r, nsq, capt, promo, root_king, cur, stack, stride); rec_store(out + n, r); }
    if (mrow) mrow[(int)stack[0].sq * S + cur] = 1;
    n++;
  }
  __device__ __forceinline__ void jump(int from, int victim, int land) {  // single man jump
    if (n < room) { Rec r; rec_jump(r, from, victim, land); rec_store(out + n, r); }
    if (mrow) mrow[from * S + land] = 1;
    n++;
  }
};

struct SelectSink {  // keeps move number `want`
  Rec rec;
  uint32_t n, want;
  __device__ __forceinline__ void quiet(int from, int to, bool king) {
    if (n == want) rec_quiet(rec, from, to, king);
    n++;
  }
  template <class BB>
  __device__ __forceinline__ void capture(int nsq, BB capt, bool promo, bool root_king, int cur, const Frame* stack,
                                          int stride) {
    if (n == want) rec_capture(rec, nsq, capt, promo, root_king, cur, stack, stride);
    n++;
  }
  __device__ __forceinline__ void jump(int from, int victim, int land) {
    if (n == want) rec_jump(rec, from, victim, land);
    n++;
  }
};

// perft frontier expansion: every move becomes a child board in the next frontier.  A board knows how many
// children it has BEFORE it writes them (popc for quiet moves, bit-parallel single-jump count, statistics pass of
// the tree search), so slots are reserved once per warp: warp prefix sum of the counts + ONE atomicAdd per warp
// (per-move reservation made every child wait for an atomic round trip: 45 % of the samples, profiles/README.md).
template <class R>
struct ExpandSink {
  using BB = typename Pos<R>::BB;
  Pos<R> parent;
  uint64_t *owm, *owk, *obm, *obk;
  uint64_t slot;  // next slot of this parent
  uint64_t cap;
  uint32_t illegal;
  __device__ __forceinline__ void child(int from, int to, BB capt, bool promo) {
    Pos<R> c = parent;
    int clock = 0;
    if (!apply_move<R>(c, clock, from, to, capt, promo)) {
      // the reference raises ValueError here (two pieces on one square): the reserved slot gets an
      // empty board, which has no moves and adds nothing to any count
      illegal++;
      c.wm = c.wk = c.bm = c.bk = 0;
    }
    if (slot < cap) { owm[slot] = c.wm; owk[slot] = c.wk; obm[slot] = c.bm; obk[slot] = c.bk; }
    slot++;
  }
  __device__ __forceinline__ void quiet(int from, int to, bool) { child(from, to, BB(0), false); }
  __device__ __forceinline__ void capture(int nsq, BB capt, bool promo, bool, int cur, const Frame* stack, int) {
    (void)nsq;
    child(stack[0].sq, cur, capt, promo);
  }
  __device__ __forceinline__ void jump(int from, int victim, int land) { child(from, land, bb_bit<BB>(victim), false); }
};

// PASS 2 sink of k_expand_deep: what a child needs (from, to, captured set, promotion, root-king flag), 16 bytes
struct KeepChildSink {
  uint4* out;
  uint32_t n, room;
  __device__ __forceinline__ void restart() { n = 0; }
  __device__ __forceinline__ void quiet(int, int, bool) {}
  __device__ __forceinline__ void jump(int, int, int) {}
  template <class BB>
  __device__ __forceinline__ void capture(int, BB capt, bool promo, bool root_king, int cur, const Frame* stack, int) {
    if (n < room) {
      const uint64_t c64 = (uint64_t)capt;
      out[n] = make_uint4((uint32_t)c64, (uint32_t)(c64 >> 32),
                          (uint32_t)stack[0].sq | ((uint32_t)cur << 8) | (promo ? 1u << 16 : 0u) | (root_king ? 1u << 24 : 0u), 0u);
    }
    n++;
  }
};

// reserve `n` consecutive slots for every lane of the (fully converged) warp: one atomicAdd per warp
__device__ __forceinline__ uint64_t warp_reserve(unsigned long long* cursor, uint32_t n) {
  const int lane = threadIdx.x & 31;
  uint32_t inc = n;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
  unsigned long long base = 0;
  if (lane == 0 && total) base = atomicAdd(cursor, (unsigned long long)total);
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  return base + inc - n;
}

__device__ __forceinline__ uint64_t splitmix64_hash(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------

// len(legal_moves) per board.  Blocks loop over 128-board tiles (grid-stride): the fused call launches it with a
// small grid so that all its blocks are resident at once and the mask-row kernel can fill the rest of the SMs (the
// block scheduler serves one grid's pending blocks at a time, drg_abi.cu SplitMovegenFn).
template <class R>
__global__ void __launch_bounds__(kBlock) k_count(int64_t n, const uint64_t* __restrict__ wm,
                                                  const uint64_t* __restrict__ wk, const uint64_t* __restrict__ bm,
                                                  const uint64_t* __restrict__ bk, const uint8_t* __restrict__ turn,
                                                  uint32_t* __restrict__ counts, DeepList deep) {
  __shared__ __align__(16) Tables<R::S> T;
  __shared__ CapList s_cap;
  __shared__ uint32_t s_quiet[kBlock];
  __shared__ unsigned s_base;
  if (threadIdx.x == 0) s_cap.n = s_cap.n_deep = 0;
  load_tables<R::S>(&T);
  __syncthreads();
  NullSink ns;
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < n; base += (int64_t)gridDim.x * kBlock) {
    const int64_t i = base + threadIdx.x;
    if (i < n) {  // phase A
      const Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn[i] & 1, i);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      uint32_t q = 0;
      if (R::kOptionalCapture || !c.any()) q = (uint32_t)gen_quiet<R, true>(p, ns, T);
      if (c.any()) {
        s_quiet[threadIdx.x] = q;
        s_cap.item[atomicAdd(&s_cap.n, 1)] = (uint8_t)threadIdx.x;
      } else {
        counts[i] = q;
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < s_cap.n; k += kBlock) {  // phase B: single-jump boards, bit-parallel
      const int t = s_cap.item[k];
      const int64_t j = base + t;
      const Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn[j] & 1, j);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      if (!c.kings && only_single_jumps<R>(p, mj)) counts[j] = s_quiet[t] + (uint32_t)single_jumps<R, false>(p, mj, ns, T);
      else s_cap.deep[atomicAdd(&s_cap.n_deep, 1)] = (uint8_t)t;
    }
    __syncthreads();
    // phase C hand-over: boards that need the tree search go to a GRID-wide list, k_walk_root runs them on
    // dense warps (a block rarely has more than a handful, which would leave 3 of its 4 warps idle here)
    if (threadIdx.x == 0 && s_cap.n_deep) s_base = atomicAdd(deep.n, (unsigned)s_cap.n_deep);
    __syncthreads();
    for (int k = threadIdx.x; k < s_cap.n_deep; k += kBlock) {
      const int t = s_cap.deep[k];
      counts[base + t] = s_quiet[t];
      deep.items[s_base + k] = (uint32_t)(base + t);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_cap.n = s_cap.n_deep = 0;
    __syncthreads();
  }
}

// tree-search boards of k_perft_count, one per thread, gathered from the whole grid: sum of their capture counts
template <class R>
__global__ void __launch_bounds__(kBlock) k_perft_count_deep(DeepList deep, const uint64_t* __restrict__ wm,
                                                             const uint64_t* __restrict__ wk,
                                                             const uint64_t* __restrict__ bm,
                                                             const uint64_t* __restrict__ bk, int turn,
                                                             unsigned long long* __restrict__ total,
                                                             unsigned long long* overflow) {
  __shared__ __align__(16) Tables<R::S> T;
  __shared__ Frame s_stack[kMaxLevels * kBlock];
  __shared__ unsigned long long s_sum[kWarps];
  load_tables<R::S>(&T);
  __syncthreads();
  const unsigned nd = *deep.n;
  unsigned long long mine = 0;
  uint32_t ovf = 0;
  for (unsigned k = blockIdx.x * kBlock + threadIdx.x; k < nd; k += gridDim.x * kBlock) {
    const int64_t j = deep.items[k];
    const Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn, j);
    MenJumps<R> mj;
    const auto c = find_capturers<R>(p, T, mj);
    uint32_t sw = 0;
    mine += (uint32_t)count_captures<R>(p, c, s_stack + threadIdx.x, kBlock, T, ovf, &sw);
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < kWarps; w++) s += s_sum[w];
    if (s) atomicAdd(total, s);
  }
  if (ovf && overflow) atomicAdd(overflow, 1ull);
}

// ---------------------------------------------------------------------------------------------
// Split walk of the tree-search boards (drg_core.cuh: Walker / WalkLane).
//
//   k_walk_root    round 0: every listed board is walked, one per lane, with a budget ordinary positions never reach;
//                  the leaves of its best metric are kept in the board's pool slot (16 records) for the emit step; a
//                  board that exceeds the budget is a monster: it starts again as one record per capturing piece.
//   k_walk_rounds  ONE cooperative launch for everything that follows: rounds 1, 2, ... walk the records handed over
//                  in the round before (grid barrier in between, it ends with the first empty round -- at once when
//                  no board exceeded its budget), then the tree sums: bottom-up (legal leaves below every record),
//                  the boards' counts, top-down (position of every record's first leaf in its board's move list).
//   k_emit_deep / k_emit_recs   the emit step: root walks copy their kept leaves (or walk again), records walk again
//                  with the budget they had, and write their leaves where the sums say.
// In the rounds and in k_emit_recs lanes do not own a fixed record: a lane that finishes takes the next item
// (warp-aggregated fetch from a global cursor, whenever a quarter of the warp is idle), and lanes that have spent
// their budget wait for each other and hand over together, so a warp stays populated although the walks differ in
// length by orders of magnitude.
// ---------------------------------------------------------------------------------------------
constexpr int kWalkRounds = 48;      // rounds a tree can be split over (the last one walks without a budget)
constexpr int kWalkRefill = 8;       // idle lanes that trigger a fetch

// Descent budgets.  Round 0 (every tree-search board, one per lane, all resident at once: the kernel is as long as
// its longest walk) gets a budget that ordinary positions never reach -- 1 Mi random-play Standard boards: mean 2.9
// descents, 4 boards above 64, none above 96 -- so that only genuine monsters pay for rounds, barriers and tree
// sums; measured with 16 (about 100 boards per Mi split): k_walk_root 119 us + k_walk_rounds 70 us + k_emit_recs 50 us
// against 65 us for the unsplit walk.  Later rounds: small budgets, so that a big tree fans out quickly.
__host__ __device__ __forceinline__ int walk_budget(int round) {
  if (round >= kWalkRounds - 1) return kNoBudget;
  return round == 0 ? 96 : (round < 12 ? 32 : 64);
}
// Round 0 by rule family.  With orthogonal captures and the value filter (Frisian, Frysk!) a descent costs about twice
// as much and trees above the budget are common in ordinary play, so the root walk gives up earlier: 1 Mi random-play
// Frisian boards 1.33 -> 1.19 ms per fused call with 32, while 48 or 32 make Standard slower (0.56 -> 0.62 / 0.63 ms).
#ifndef DRG_ROOT_BUDGET_ORTHO
#define DRG_ROOT_BUDGET_ORTHO 32
#endif
template <class R>
__host__ __device__ __forceinline__ int root_budget() { return R::kOrtho ? DRG_ROOT_BUDGET_ORTHO : walk_budget(0); }
// Budget of round r >= 1, chosen when the round starts from the number of records it holds.  While there are fewer
// records than lanes, a round is pure latency -- fetch, the longest walk (1-2 us per descent), hand-over, grid barrier
// -- and a tree is finished soonest in short rounds: 1 Mi random-play Frisian boards (637 boards above 32 descents, the
// largest 1 461) kept k_walk_rounds busy for 750 us with budgets of 32, 438 us with 10 / 20, 390 us with 6 / 12 (eight
// rounds; what remains is the per-round overhead, 16 / 24 is slower again).  Once the lanes are saturated (synthetic
// king endgames: unchanged at 11 ms) larger budgets cost less hand-over traffic.  The emitting re-walk reads the
// budget each round had from WalkPool::budget.
__device__ __forceinline__ int round_budget(int round, unsigned n_records, unsigned n_lanes) {
  if (round >= kWalkRounds - 1) return kNoBudget;
#ifndef DRG_LOW_BUDGET
#define DRG_LOW_BUDGET 6
#define DRG_MID_BUDGET 12
#endif
  if (4ull * n_records <= n_lanes) return DRG_LOW_BUDGET;
  if (n_records <= n_lanes) return DRG_MID_BUDGET;
  return round < 12 ? 32 : 64;
}

struct WalkPool {       // device view of the record pool (memory owned by drg_abi.cu)
  WalkRec* recs;        // records, in the order they were handed over: round 1's first, then round 2's, ...
  WalkRes* res;         // result of walking record i
  uint32_t* sub;        // legal leaves of record i and of everything handed over below it
  uint32_t* base;       // index of record i's first legal leaf among its board's capture moves
  unsigned* count;      // count[r] = records reserved for round r (handed over during round r - 1), r >= 1; count[0] = 0
  unsigned* limit;      // limit[r] = ~(index within the round of the first reservation that did not fit the pool); 0: all fit
  unsigned* cursor;     // cursor[r]: next item of round r to fetch (cursor[0]: k_walk_root, cursor[kWalkRounds + e]: emit kernels)
  unsigned* bar;        // grid barrier of k_walk_rounds
  unsigned* budget;     // budget[r]: descent budget of round r (written by k_walk_rounds, read by k_emit_recs)
  unsigned cap;         // records the pool holds
};

struct PoolAlloc {      // WalkLane's allocator: consecutive records of the NEXT round
  WalkPool wp;
  int round;
  unsigned next0;       // index of the next round's first record
  // One atomicAdd per reservation (a compare-and-swap loop on this one word collapsed under 38 000 lanes).  A
  // reservation that does not fit leaves the counter advanced, so every later one fails too: the valid records of a
  // round are a prefix, and limit[] remembers where it ends (walk_round_size).
  __device__ __forceinline__ long long reserve(int n) {
    const unsigned old = atomicAdd(wp.count + round + 1, (unsigned)n);
    if ((unsigned long long)next0 + old + (unsigned)n > wp.cap) {
      atomicMax(wp.limit + round + 1, ~old);
      return -1;
    }
    return (long long)next0 + old;
  }
  __device__ __forceinline__ WalkRec* at(long long i) const { return wp.recs + i; }
};
// records of round r that exist
__device__ __forceinline__ unsigned walk_round_size(const WalkPool& wp, int r) {
  const unsigned c = *reinterpret_cast<volatile unsigned*>(wp.count + r), l = ~*reinterpret_cast<volatile unsigned*>(wp.limit + r);
  return c < l ? c : l;
}
struct NoAlloc {        // emitting walks never allocate
  __device__ __forceinline__ long long reserve(int) { return -1; }
  __device__ __forceinline__ WalkRec* at(long long) const { return nullptr; }
};

// warp-aggregated fetch for the idle lanes of a (converged) warp: index of this lane's item, >= n when there is none
__device__ __forceinline__ unsigned walk_fetch(unsigned* cursor, unsigned idle_mask, bool idle) {
  const int lane = threadIdx.x & 31;
  unsigned base = 0;
  if (lane == 0) base = atomicAdd(cursor, (unsigned)__popc(idle_mask));
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  return idle ? base + __popc(idle_mask & ((1u << lane) - 1)) : 0xFFFFFFFFu;
}

// Lanes of a warp that take walks when `n_items` are dealt over `n_warps` warps.  32 walks in one warp follow 32
// different paths through the same code and serialise: a descent then costs about 3 us instead of 1.  While there are
// fewer items than lanes, the items are therefore spread over ALL warps, a few lanes each.
__device__ __forceinline__ unsigned walk_lane_mask(unsigned n_items, unsigned n_warps) {
  const unsigned per_warp = (n_items + n_warps - 1) / n_warps;
  return per_warp >= 32 ? 0xFFFFFFFFu : ((1u << (per_warp ? per_warp : 1u)) - 1u);
}

__device__ __forceinline__ WalkRec load_rec(const WalkRec* src) {
  WalkRec r;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(&r);
  d4[0] = s4[0]; d4[1] = s4[1]; d4[2] = s4[2];
  return r;
}

template <class R>
__global__ void __launch_bounds__(kBlock) k_walk_root(DeepList deep, const uint64_t* __restrict__ wm,
                                                      const uint64_t* __restrict__ wk, const uint64_t* __restrict__ bm,
                                                      const uint64_t* __restrict__ bk, const uint8_t* __restrict__ turn,
                                                      uint32_t* counts, bool set_counts,
                                                      unsigned long long* overflow, uint32_t* __restrict__ stat,
                                                      DrgMove* __restrict__ keep_pool, unsigned pool_boards,
                                                      WalkRes* __restrict__ rres, uint32_t* __restrict__ bword,
                                                      WalkPool wp) {
  __shared__ __align__(16) Tables<R::S> T;
  __shared__ Frame s_stack[kMaxLevels * kBlock];
  const unsigned nd = *deep.n;
  if ((unsigned long long)blockIdx.x * kBlock >= nd) return;
  load_tables<R::S>(&T);
  __syncthreads();
  PoolAlloc alloc{wp, 0, 0u};
  uint32_t ovf = 0;
  Frame* stack = s_stack + threadIdx.x;
  // one board per lane, all of them resident at once: the kernel is as long as its longest walk, which the budget caps
  for (unsigned k = blockIdx.x * kBlock + threadIdx.x; k < nd; k += gridDim.x * kBlock) {
    const int64_t j = deep.items[k];
    const Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn[j] & 1, j);
    MenJumps<R> mj;
    const auto c = find_capturers<R>(p, T, mj);
    KeepSink keep;
    keep.out = keep_pool + (size_t)k * kKeepRecords;
    keep.n = 0;
    keep.room = k < pool_boards ? (uint32_t)kKeepRecords : 0u;
    WalkLane<R, 2> lane;
    lane.start_root(p, c, k, root_budget<R>(), T);
    for (;;) {
      const int s = lane.advance(keep, stack, kBlock, T, ovf);
      if (s == kWalkDone) break;
      if (s == kWalkWantsSplit && lane.abandon(c, alloc)) {  // a monster: its pieces start again as records of round 1
        keep.n = 0;
        break;
      }
    }
    const WalkRes res = lane.result();
    const uint32_t bw = walk_board_word<R>(lane.st);
    rres[k] = res;
    bword[k] = bw;  // records of this board are walked in later rounds (atomicMax there)
    stat[k] = pack_keep_stat(lane.st, keep.n, keep.n <= keep.room);
    if (counts && !lane.split) {
      // the whole tree was walked here: the board's count is final (boards that gave up get theirs from
      // k_walk_rounds).  set_counts: the list was built by an emit step, no count step wrote the quiet part --
      // optional-capture variants list their quiet moves as well (american.py:96)
      uint32_t q = 0;
      if (set_counts && R::kOptionalCapture) {
        NullSink ns;
        q = (uint32_t)gen_quiet<R, true>(p, ns, T);
      }
      counts[j] = (set_counts ? q : counts[j]) + walk_kept<R>(res, bw);
    }
  }
  if (ovf && overflow) atomicAdd(overflow, 1ull);
}

// barrier over the (co-resident: cooperative launch) grid; `target` counts the arrivals expected so far
// counts -> offsets as the tail of a cooperative kernel (k_walk_rounds): every block sums its contiguous share of the
// counts, one grid barrier, every block adds up the shares before its own and scans its share again into offsets
struct ScanTail {
  int64_t n;
  uint64_t* offsets;              // [n + 1]; NULL: no scan
  unsigned long long* share_sums; // [gridDim.x]
  unsigned* bar;                  // its own barrier word (zeroed with the walk counters)
};

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  __syncthreads();
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  unsigned long long t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_warp[w];
  return t;
}

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target);

__device__ __forceinline__ void scan_tail(const ScanTail& sc, const uint32_t* counts, unsigned long long* s_u64) {
  constexpr int kRow = kBlock * 4;  // four consecutive counts per thread and row
  const int64_t rows = (sc.n + kRow - 1) / kRow;
  const int64_t rows_per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * rows_per * kRow;
  int64_t hi = lo + rows_per * kRow;
  if (hi > sc.n) hi = sc.n;
  unsigned long long mine = 0;
  for (int64_t e = lo + (int64_t)threadIdx.x * 4; e < hi; e += kRow)
    for (int q = 0; q < 4; q++)
      if (e + q < hi) mine += counts[e + q];
  const unsigned long long share = block_sum_u64(mine, s_u64);
  if (threadIdx.x == 0) sc.share_sums[blockIdx.x] = share;
  unsigned target = 0;
  grid_barrier(sc.bar, target);
  unsigned long long before = 0;
  for (unsigned b = threadIdx.x; b < blockIdx.x; b += kBlock) before += sc.share_sums[b];
  unsigned long long carry = block_sum_u64(before, s_u64);
  for (int64_t e0 = lo; e0 < hi; e0 += kRow) {
    const int64_t e = e0 + (int64_t)threadIdx.x * 4;
    uint32_t c[4];
    unsigned long long v = 0;
    for (int q = 0; q < 4; q++) {
      c[q] = e + q < hi ? counts[e + q] : 0u;
      v += c[q];
    }
    // exclusive scan of v over the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_u64[warp] = inc;
    __syncthreads();
    unsigned long long ex = carry + inc - v, row_total = 0;
    for (int w = 0; w < kWarps; w++) {
      const unsigned long long t = s_u64[w];
      if (w < warp) ex += t;
      row_total += t;
    }
    for (int q = 0; q < 4; q++) {
      if (e + q < hi) sc.offsets[e + q] = ex;
      ex += c[q];
    }
    carry += row_total;
  }
  if (hi == sc.n && lo < sc.n && threadIdx.x == 0) sc.offsets[sc.n] = carry;
}

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    // the launch is cooperative, so every block is resident and arrives; the spin limit (seconds) only keeps a logic
    // error from hanging the device -- it then shows as a wrong result in the tests instead
    for (unsigned spins = 0; *reinterpret_cast<volatile unsigned*>(bar) < target && spins < (1u << 24); spins++) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}

template <class R>
__global__ void __launch_bounds__(kBlock) k_walk_rounds(DeepList deep, const uint64_t* __restrict__ wm,
                                                        const uint64_t* __restrict__ wk,
                                                        const uint64_t* __restrict__ bm,
                                                        const uint64_t* __restrict__ bk,
                                                        const uint8_t* __restrict__ turn, uint32_t* counts,
                                                        bool set_counts, unsigned long long* overflow,
                                                        const WalkRes* __restrict__ rres, uint32_t* bword, WalkPool wp,
                                                        ScanTail sc) {
  __shared__ __align__(16) Tables<R::S> T;
  __shared__ Frame s_stack[kMaxLevels * kBlock];
  __shared__ unsigned s_start[kWalkRounds + 1];  // s_start[r]: first record of round r
  const unsigned nd = *deep.n;
  const unsigned tid = blockIdx.x * kBlock + threadIdx.x, nthreads = gridDim.x * kBlock;
  unsigned target = 0;
  int rounds = 1;  // rounds 1 .. rounds-1 have records
  // no board exceeded its budget: k_walk_root has finished every count
  if (walk_round_size(wp, 1) != 0) {
  load_tables<R::S>(&T);
  if (threadIdx.x == 0) s_start[1] = 0;
  __syncthreads();
  {
    NullSink ns;
    uint32_t ovf = 0;
    Frame* stack = s_stack + threadIdx.x;
    for (int r = 1; r < kWalkRounds; r++) {
      const unsigned n_r = walk_round_size(wp, r);  // final since the barrier before
      if (n_r == 0) break;
      rounds = r + 1;
      const unsigned lo = s_start[r];
      if (threadIdx.x == 0) s_start[r + 1] = lo + n_r;
      PoolAlloc alloc{wp, r, lo + n_r};
      const int budget = round_budget(r, n_r, nthreads);
      if (tid == 0) wp.budget[r] = (unsigned)budget;
      // lanes to spare: hand the shallowest open frames over child by child, so that a tree fans out in fewer rounds
      // (what a walker hands over does not change what it walks: the emitting re-walk needs no record of this)
#ifndef DRG_LOW_EXPLODE
#define DRG_LOW_EXPLODE 3
#define DRG_MID_EXPLODE 2
#endif
      const int explode = 4ull * n_r <= nthreads ? DRG_LOW_EXPLODE : (n_r <= nthreads ? DRG_MID_EXPLODE : kWalkExplode);
      WalkLane<R, 0> lane;
      bool active = false, parked = false, more = true;
      unsigned w = 0;
      const unsigned takers = walk_lane_mask(n_r, nthreads / 32);
      const bool taker = (takers >> (threadIdx.x & 31)) & 1u;
      for (;;) {
        __syncwarp();
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active && taker);
        if (more && (idle == takers || __popc(idle) >= kWalkRefill)) {
          const unsigned kk = walk_fetch(wp.cursor + r, idle, !active && taker);
          more = __any_sync(0xFFFFFFFFu, !active && taker && kk < n_r);
          if (!active && taker && kk < n_r) {
            w = lo + kk;
            const WalkRec rec = load_rec(wp.recs + w);
            const int64_t j = deep.items[rec.item];
            const Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn[j] & 1, j);
            lane.start_rec(p, rec, budget, stack, kBlock, T);
            lane.next_round = r + 1;
            lane.explode = explode;
            active = true;
          }
        }
        const unsigned act = __ballot_sync(0xFFFFFFFFu, active), park = __ballot_sync(0xFFFFFFFFu, parked);
        if (!act) break;
        bool done = false;
        if (park && (__popc(park) >= kWalkRefill || park == act)) {  // hand over together
          if (parked) {
            done = lane.hand_over(stack, kBlock, alloc, T);
            parked = false;
          }
        }
        if (active && !parked && !done) {
          const int s = lane.advance(ns, stack, kBlock, T, ovf);
          done = s == kWalkDone;
          parked = s == kWalkWantsSplit;
        }
        if (done) {
          wp.res[w] = lane.result();
          if (lane.st.best) atomicMax(bword + lane.item, walk_board_word<R>(lane.st));
          active = false;
        }
      }
      grid_barrier(wp.bar, target);
    }
    if (ovf && overflow) atomicAdd(overflow, 1ull);
    // bottom-up: legal leaves of every record's subtree (the boards' best metrics are final now)
    for (int r = rounds - 1; r >= 1; r--) {
      const unsigned lo = s_start[r], hi = s_start[r + 1];
      for (unsigned w = lo + tid; w < hi; w += nthreads) {
        const WalkRes res = wp.res[w];
        uint32_t s = walk_kept<R>(res, bword[wp.recs[w].item]);
        const uint32_t nc = res.info >> 24;
        for (uint32_t c = 0; c < nc; c++) s += wp.sub[res.child0 + c];
        wp.sub[w] = s;
      }
      grid_barrier(wp.bar, target);
    }
  }
  // the boards that handed records over: len(legal captures) = root walk's legal leaves + everything below its
  // records; and where the records' leaves start
  for (unsigned k = tid; k < nd; k += nthreads) {
    const WalkRes res = rres[k];
    if (!(res.info & kWalkSplit)) continue;
    uint32_t run = walk_kept<R>(res, bword[k]);
    const uint32_t nc = res.info >> 24;
    for (uint32_t c = 0; c < nc; c++) {
      wp.base[res.child0 + c] = run;
      run += wp.sub[res.child0 + c];
    }
    if (counts) {
      // set_counts: the list was built by an emit step (no count step wrote the quiet part): optional-capture variants
      // list their quiet moves as well (american.py:96)
      const int64_t j = deep.items[k];
      uint32_t q = 0;
      if (set_counts && R::kOptionalCapture) {
        NullSink ns;
        q = (uint32_t)gen_quiet<R, true>(load_pos<R>(wm, wk, bm, bk, turn[j] & 1, j), ns, T);
      }
      counts[j] = (set_counts ? q : counts[j]) + run;
    }
  }
  // top-down: first-leaf positions of the records of records
  for (int r = 1; r < rounds - 1; r++) {
    grid_barrier(wp.bar, target);
    const unsigned lo = s_start[r], hi = s_start[r + 1];
    for (unsigned w = lo + tid; w < hi; w += nthreads) {
      const WalkRes res = wp.res[w];
      const uint32_t nc = res.info >> 24;
      if (!nc) continue;
      uint32_t run = wp.base[w] + walk_kept<R>(res, bword[wp.recs[w].item]);
      for (uint32_t c = 0; c < nc; c++) {
        wp.base[res.child0 + c] = run;
        run += wp.sub[res.child0 + c];
      }
    }
  }
  }
  // every count of the batch is final now: counts -> offsets here, behind one more grid barrier, instead of three more
  // launches on a chain whose kernels are mostly launch latency (the fused call's count -> records chain)
  if (sc.offsets) {
    __shared__ unsigned long long s_u64[kWarps];
    if (walk_round_size(wp, 1) != 0) grid_barrier(wp.bar, target);
    scan_tail(sc, counts, s_u64);
  }
}

// perft leaf level: sum of len(legal_moves) over a same-turn frontier (persistent blocks, 128-board tiles)
template <class R>
__global__ void __launch_bounds__(kBlock) k_perft_count(int64_t n, const uint64_t* __restrict__ wm,
                                                        const uint64_t* __restrict__ wk,
                                                        const uint64_t* __restrict__ bm,
                                                        const uint64_t* __restrict__ bk, int turn,
                                                        unsigned long long* __restrict__ total, DeepList deep) {
  __shared__ __align__(16) Tables<R::S> T;
  __shared__ CapList s_cap;
  __shared__ unsigned long long s_sum[kWarps];
  __shared__ unsigned s_base;
  if (threadIdx.x == 0) s_cap.n = s_cap.n_deep = 0;
  load_tables<R::S>(&T);
  __syncthreads();
  unsigned long long mine = 0;
  NullSink ns;
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < n; base += (int64_t)gridDim.x * kBlock) {
    const int64_t i = base + threadIdx.x;
    if (i < n) {  // phase A
      const Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn, i);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      if (R::kOptionalCapture || !c.any()) mine += (unsigned long long)gen_quiet<R, true>(p, ns, T);
      if (c.any()) s_cap.item[atomicAdd(&s_cap.n, 1)] = (uint8_t)threadIdx.x;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < s_cap.n; k += kBlock) {  // phase B: single-jump boards, bit-parallel
      const int t = s_cap.item[k];
      const Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn, base + t);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      if (!c.kings && only_single_jumps<R>(p, mj)) mine += (unsigned long long)single_jumps<R, false>(p, mj, ns, T);
      else s_cap.deep[atomicAdd(&s_cap.n_deep, 1)] = (uint8_t)t;
    }
    __syncthreads();
    // phase C hand-over: tree-search boards go to the grid-wide list drained by k_perft_count_deep
    if (threadIdx.x == 0 && s_cap.n_deep) s_base = atomicAdd(deep.n, (unsigned)s_cap.n_deep);
    __syncthreads();
    for (int k = threadIdx.x; k < s_cap.n_deep; k += kBlock) deep.items[s_base + k] = (uint32_t)(base + s_cap.deep[k]);
    __syncthreads();
    if (threadIdx.x == 0) s_cap.n = s_cap.n_deep = 0;
    __syncthreads();
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < kWarps; w++) s += s_sum[w];
    if (s) atomicAdd(total, s);
  }
}

// perft over several GPUs: at one ply every rank keeps the boards whose content hash falls to it.  A partition by
// CONTENT is the same on every rank although the order in which k_expand writes children (atomic slot reservation)
// differs from run to run; transpositions land on one rank and are counted there as often as they occur.
__device__ __forceinline__ uint32_t board_rank(uint64_t wm, uint64_t wk, uint64_t bm, uint64_t bk, uint32_t world) {
  const uint64_t h = splitmix64_hash(wm * 0x9E3779B97F4A7C15ull ^ wk * 0xC2B2AE3D27D4EB4Full ^ bm * 0x165667B19E3779F9ull ^
                                     bk * 0x27D4EB2F165667C5ull);
  return (uint32_t)(((h >> 32) * (uint64_t)world) >> 32);
}

__global__ void __launch_bounds__(256) k_perft_shard(int64_t n, const uint64_t* __restrict__ wm,
                                                     const uint64_t* __restrict__ wk, const uint64_t* __restrict__ bm,
                                                     const uint64_t* __restrict__ bk, uint32_t rank, uint32_t world,
                                                     uint64_t* owm, uint64_t* owk, uint64_t* obm, uint64_t* obk,
                                                     unsigned long long* cursor) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const bool mine = i < n && board_rank(wm[i], wk[i], bm[i], bk[i], world) == rank;
  const unsigned m = __ballot_sync(0xFFFFFFFFu, mine);
  if (!m) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (mine) {
    const unsigned long long o = base + __popc(m & ((1u << lane) - 1));
    owm[o] = wm[i]; owk[o] = wk[i]; obm[o] = bm[i]; obk[o] = bk[i];
  }
}

// the driver's device counters as the DRG_C_* vector of include/drg.h, in device memory (for an all-reduce over NCCL
// without a host round trip)
__global__ void k_perft_counters(const unsigned long long* __restrict__ ctr, int k_total, int k_ovf, int k_illegal,
                                 unsigned long long movegen, unsigned long long* __restrict__ out) {
  if (threadIdx.x >= DRG_N_COUNTERS) return;
  unsigned long long v = 0;
  if (threadIdx.x == DRG_C_NODES) v = ctr[k_total];
  if (threadIdx.x == DRG_C_MOVEGEN) v = movegen;
  if (threadIdx.x == DRG_C_OVERFLOW) v = ctr[k_ovf];
  if (threadIdx.x == DRG_C_ILLEGAL) v = ctr[k_illegal];
  out[threadIdx.x] = v;
}

// perft interior level: children of every board written to the next frontier
template <class R>
__global__ void __launch_bounds__(kBlock) k_expand(int64_t n, const uint64_t* __restrict__ wm,
                                                   const uint64_t* __restrict__ wk, const uint64_t* __restrict__ bm,
                                                   const uint64_t* __restrict__ bk, int turn, uint64_t* owm,
                                                   uint64_t* owk, uint64_t* obm, uint64_t* obk, uint64_t cap,
                                                   unsigned long long* cursor, unsigned long long* illegal,
                                                   DeepList deep) {
  __shared__ __align__(16) Tables<R::S> T;
  __shared__ CapList s_cap;
  if (threadIdx.x == 0) s_cap.n = s_cap.n_deep = 0;
  load_tables<R::S>(&T);
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kBlock;
  const int64_t i = base + threadIdx.x;
  ExpandSink<R> sink;
  sink.owm = owm; sink.owk = owk; sink.obm = obm; sink.obk = obk;
  sink.cap = cap;
  sink.illegal = 0;
  NullSink ns;
  {  // phase A: quiet children (every lane takes part in the warp-wide slot reservation)
    const bool valid = i < n;
    bool quiet_here = false, has_cap = false;
    uint32_t nq = 0;
    if (valid) {
      sink.parent = load_pos<R>(wm, wk, bm, bk, turn, i);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(sink.parent, T, mj);
      has_cap = c.any();
      quiet_here = R::kOptionalCapture || !has_cap;
      if (quiet_here) nq = (uint32_t)gen_quiet<R, true>(sink.parent, ns, T);
    }
    __syncwarp();
    sink.slot = warp_reserve(cursor, nq);
    if (quiet_here && nq) gen_quiet<R, false>(sink.parent, sink, T);
    if (has_cap) s_cap.item[atomicAdd(&s_cap.n, 1)] = (uint8_t)threadIdx.x;
  }
  __syncthreads();
  // phase B: capture boards re-dealt onto consecutive threads; single-jump boards are finished here, boards that
  // need the tree search go to the grid-wide list of k_expand_deep
  for (int k0 = 0; k0 < s_cap.n; k0 += kBlock) {
    const int k = k0 + threadIdx.x;
    const bool active = k < s_cap.n;
    bool simple = false;
    uint32_t nj = 0;
    MenJumps<R> mj;
    int64_t j = 0;
    if (active) {
      j = base + s_cap.item[k];
      sink.parent = load_pos<R>(wm, wk, bm, bk, turn, j);
      const auto c = find_capturers<R>(sink.parent, T, mj);
      simple = !c.kings && only_single_jumps<R>(sink.parent, mj);
      if (simple) nj = (uint32_t)single_jumps<R, false>(sink.parent, mj, ns, T);
    }
    __syncwarp();
    sink.slot = warp_reserve(cursor, nj);
    if (simple) single_jumps<R, true>(sink.parent, mj, sink, T);
    else if (active) list_append(deep, (uint32_t)j);
    __syncwarp();
  }
  if (sink.illegal) atomicAdd(illegal, (unsigned long long)sink.illegal);
}

// tree-search boards of k_expand, one per thread: statistics pass (= the number of children), warp-wide slot
// reservation, emitting pass
template <class R>
__global__ void __launch_bounds__(kBlock) k_expand_deep(DeepList deep, const uint64_t* __restrict__ wm,
                                                        const uint64_t* __restrict__ wk,
                                                        const uint64_t* __restrict__ bm,
                                                        const uint64_t* __restrict__ bk, int turn, uint64_t* owm,
                                                        uint64_t* owk, uint64_t* obm, uint64_t* obk, uint64_t cap,
                                                        unsigned long long* cursor, unsigned long long* overflow,
                                                        unsigned long long* illegal, uint4* __restrict__ kept_pool) {
  __shared__ __align__(16) Tables<R::S> T;
  __shared__ Frame s_stack[kMaxLevels * kBlock];
  load_tables<R::S>(&T);
  __syncthreads();
  const unsigned nd = *deep.n;
  ExpandSink<R> sink;
  sink.owm = owm; sink.owk = owk; sink.obm = obm; sink.obk = obk;
  sink.cap = cap;
  sink.illegal = 0;
  uint32_t ovf = 0;
  constexpr bool kKeepChildren = R::kFilter != FILTER_VALUE;
  KeepChildSink keep;
  keep.out = kept_pool + ((size_t)blockIdx.x * kBlock + threadIdx.x) * kKeepRecords;
  keep.n = 0;
  keep.room = kKeepRecords;
  for (unsigned k0 = blockIdx.x * kBlock; k0 < nd; k0 += gridDim.x * kBlock) {
    const unsigned k = k0 + threadIdx.x;
    const bool active = k < nd;
    uint32_t nc = 0, sw = 0;
    Capturers<typename Pos<R>::BB> c;
    c.men = c.kings = 0;
    if (active) {
      sink.parent = load_pos<R>(wm, wk, bm, bk, turn, (int64_t)deep.items[k]);
      MenJumps<R> mj;
      c = find_capturers<R>(sink.parent, T, mj);
      // statistics (= the number of children) AND the children of the best metric, kept in this thread's pool slot:
      // after the slots are reserved they are applied from there, the tree is not walked a second time
      // (not for the Frisian value filter: the best metric changes often there and the kept set keeps restarting --
      // measured 7 % slower; those boards walk twice as before)
      if (kKeepChildren) {
        keep.n = 0;
        nc = (uint32_t)count_captures_keep<R>(sink.parent, c, keep, s_stack + threadIdx.x, kBlock, T, ovf, &sw);
      } else {
        nc = (uint32_t)count_captures<R>(sink.parent, c, s_stack + threadIdx.x, kBlock, T, ovf, &sw);
      }
    }
    __syncwarp();
    sink.slot = warp_reserve(cursor, nc);
    if (active && !kKeepChildren) {
      emit_captures_with_stat<R>(sink.parent, c, sink, s_stack + threadIdx.x, kBlock, T, ovf, sw);
    } else if (active) {
      if (sw & 0x40000000u) {
        const int kept = (int)((sw >> 20) & 0xFFu);
        const bool kings_only = R::kFilter == FILTER_VALUE && (sw >> 31);
        for (int r = 0; r < kept; r++) {
          const uint4 e = keep.out[r];
          if (kings_only && !(e.z >> 24)) continue;
          const uint64_t capt = (uint64_t)e.x | ((uint64_t)e.y << 32);
          sink.child((int)(e.z & 0xFFu), (int)((e.z >> 8) & 0xFFu), typename Pos<R>::BB(capt), ((e.z >> 16) & 1u) != 0);
        }
      } else {  // more children than the pool slot holds: walk again
        CapStat st;
        unpack_keep_stat(sw, st);
        gen_captures<R, 1>(sink.parent, c, st, sink, s_stack + threadIdx.x, kBlock, T, ovf);
      }
    }
    __syncwarp();
  }
  if (ovf) atomicAdd(overflow, 1ull);
  if (sink.illegal) atomicAdd(illegal, (unsigned long long)sink.illegal);
}

// board.push(chosen[i]) per board (base.py:215-293)
template <class R>
__global__ void __launch_bounds__(kBlock) k_apply(int64_t n, uint64_t* wm, uint64_t* wk, uint64_t* bm, uint64_t* bk,
                                                  uint8_t* turn, uint8_t* clock, const DrgMove* __restrict__ chosen,
                                                  uint8_t* status) {
  using BB = typename Pos<R>::BB;
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn[i] & 1, i);
  int clk = clock ? clock[i] : 0;
  const uint4* src = reinterpret_cast<const uint4*>(chosen + i);
  const uint4 a = src[0], b = src[1];
  Rec r;
  r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
  r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
  const uint64_t capt = (uint64_t)r.w[0] | ((uint64_t)r.w[1] << 32);
  int nsq = rec_nsq(r);
  if (nsq > DRG_MAX_PATH) nsq = DRG_MAX_PATH;
  Rec r2 = r;
  r2.w[2] = (r.w[2] & ~0xFFu) | (uint32_t)nsq;
  const bool ok = nsq >= 1 && apply_move<R>(p, clk, rec_from(r), rec_to(r2), BB(capt) & Geo<R::S>::MASK,
                                            ((r.w[2] >> 8) & DRG_MF_PROMO) != 0);
  if (status) status[i] = ok ? 0 : 1;
  if (ok) {
    wm[i] = p.wm; wk[i] = p.wk; bm[i] = p.bm; bk[i] = p.bk;
    turn[i] = (uint8_t)p.turn;
    if (clock) clock[i] = (uint8_t)(clk > 255 ? 255 : clk);
  }
}

// ---------------------------------------------------------------------------------------------
// Search-leaf service: the children of a batch in move-list order, and the reference engine's
// static evaluation of a batch.
// ---------------------------------------------------------------------------------------------
// child j = copy of its parent with moves[j] pushed (board.copy(); board.push(move), base.py:215-293), where the
// parent of j is the board i with offsets[i] <= j < offsets[i+1] -- the layout drg_movegen writes.  One thread per
// child; the parent index comes from a binary search in the offsets (they stay in L2).  HBM: 32 B record + 34 B
// parent (cached across siblings) in, 34 B + 4 B + 1 B out per child.
template <class R>
__global__ void __launch_bounds__(kBlock) k_children(int64_t n, const uint64_t* __restrict__ wm,
                                                     const uint64_t* __restrict__ wk,
                                                     const uint64_t* __restrict__ bm,
                                                     const uint64_t* __restrict__ bk,
                                                     const uint8_t* __restrict__ turn,
                                                     const uint8_t* __restrict__ clock,
                                                     const uint64_t* __restrict__ offsets, int64_t total,
                                                     const DrgMove* __restrict__ moves, uint64_t* owm, uint64_t* owk,
                                                     uint64_t* obm, uint64_t* obk, uint8_t* oturn, uint8_t* oclock,
                                                     uint32_t* parent, uint8_t* status) {
  using BB = typename Pos<R>::BB;
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const uint64_t real = offsets[n] < (uint64_t)total ? offsets[n] : (uint64_t)total;  // `total` slots, offsets[n] children
  const uint64_t jw = (uint64_t)(j - lane);  // first child of this warp
  if (jw >= real) return;                    // the whole warp leaves together
  // parent of the warp's first child, i0 = largest i with offsets[i] <= jw: a 32-ary search, one probe per lane and
  // round (4 rounds for 1 Mi boards; the per-thread binary search it replaces was 240 of the kernel's 630
  // instructions per warp and 20 dependent L2 round trips)
  int64_t lo = 0, len = n;
  while (len > 1) {
    const int64_t step = (len + 31) >> 5;
    const int64_t probe = lo + lane * step;
    const bool le = probe < lo + len && offsets[probe] <= jw;
    const int k = __popc(__ballot_sync(0xFFFFFFFFu, le)) - 1;  // offsets[lo] <= jw always holds
    const int64_t end = lo + len;
    lo += k * step;
    len = end - lo < step ? end - lo : step;
  }
  // the 32 children of the warp mostly share a handful of parents: lane l holds offsets[i0+1+l], and a child's
  // parent is i0 + (number of those that are <= its index), a binary search over the window by shuffles
  const int64_t wi = lo + 1 + lane;
  const uint64_t w = wi <= n ? offsets[wi] : ~0ull;
  int c = 0;
#pragma unroll
  for (int sft = 16; sft; sft >>= 1) {
    const uint64_t v = __shfl_sync(0xFFFFFFFFu, w, c + sft - 1);
    if (v <= (uint64_t)j) c += sft;
  }
  const uint64_t w31 = __shfl_sync(0xFFFFFFFFu, w, 31);
  if ((uint64_t)j >= real) return;
  int64_t i = lo + c;
  if (c == 31 && w31 <= (uint64_t)j) {  // more than 32 parents under this warp (runs of boards without moves)
    int64_t a = lo + 32, b = n;         // offsets[a] <= j < offsets[b]
    while (b - a > 1) {
      const int64_t mid = (a + b) >> 1;
      if (offsets[mid] <= (uint64_t)j) a = mid;
      else b = mid;
    }
    i = a;
  }
  Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn[i] & 1, i);
  int clk = clock ? clock[i] : 0;
  const uint4* src = reinterpret_cast<const uint4*>(moves + j);
  const uint4 a = src[0], b = src[1];
  Rec r;
  r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
  r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
  const uint64_t capt = (uint64_t)r.w[0] | ((uint64_t)r.w[1] << 32);
  int nsq = rec_nsq(r);
  if (nsq > DRG_MAX_PATH) nsq = DRG_MAX_PATH;
  Rec r2 = r;
  r2.w[2] = (r.w[2] & ~0xFFu) | (uint32_t)nsq;
  const bool ok = nsq >= 1 && apply_move<R>(p, clk, rec_from(r), rec_to(r2), BB(capt) & Geo<R::S>::MASK,
                                            ((r.w[2] >> 8) & DRG_MF_PROMO) != 0);
  // where the reference raises ValueError the child is the unchanged parent and status[j] = 1 (as drg_apply)
  owm[j] = p.wm; owk[j] = p.wk; obm[j] = p.bm; obk[j] = p.bk;
  oturn[j] = (uint8_t)p.turn;
  if (oclock) oclock[j] = (uint8_t)(clk > 255 ? 255 : clk);
  if (parent) parent[j] = (uint32_t)i;
  if (status) status[j] = ok ? 0 : 1;
}

// AlphaBetaEngine.evaluate (alpha_beta.py:205-244): material (man 1.0, king 2.5) + piece-square tables, from the
// side to move's point of view, float64.  Bit-exact with the reference, so the additions happen in numpy's order:
// `table[mask].sum()` is a pairwise sum whose base case is a plain loop below 8 elements and eight interleaved
// accumulators above (a side has at most 50 pieces, one block of the pairwise recursion).  __dadd_rn / __dmul_rn
// keep the compiler from contracting a multiply and an add into an FMA.
template <int S>
__device__ __forceinline__ double pst_sum(const double* __restrict__ table, uint64_t bb) {
  const int n = __popcll(bb);
  // squares ascending, on 32-bit halves (the 64-bit find-first-set / clear pair was most of the kernel's instructions)
  uint32_t lo = (uint32_t)bb, hi = (uint32_t)(bb >> 32);
  const double* t = table;
  auto next = [&]() {
    if (S > 32 && lo == 0) {
      lo = hi;
      hi = 0;
      t = table + 32;
    }
    const int sq = __ffs((int)lo) - 1;
    lo &= lo - 1;
    return t[sq];
  };
  if (n < 8) {
    double r = 0.0;
    for (int k = 0; k < n; k++) r = __dadd_rn(r, next());
    return r;
  }
  double r[8];
#pragma unroll
  for (int k = 0; k < 8; k++) r[k] = next();
  for (int i = 8; i < n - (n & 7); i += 8) {
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = __dadd_rn(r[k], next());
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  while (lo | hi) res = __dadd_rn(res, next());
  return res;
}

template <int S>
__global__ void __launch_bounds__(kBlock) k_evaluate(int64_t n, const uint64_t* __restrict__ wm,
                                                     const uint64_t* __restrict__ wk,
                                                     const uint64_t* __restrict__ bm,
                                                     const uint64_t* __restrict__ bk,
                                                     const uint8_t* __restrict__ turn,
                                                     const double* __restrict__ pst, double* __restrict__ score) {
  __shared__ double s_pst[4 * S];  // man black, man white, king black, king white (alpha_beta.py:166-177)
  for (int k = threadIdx.x; k < 4 * S; k += kBlock) s_pst[k] = pst[k];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  // board._pos: one code per square, priority of BaseBoard._get (base.py:152-163)
  const uint64_t all = Geo<S>::MASK;
  const uint64_t c_wm = wm[i] & all, c_wk = wk[i] & all & ~c_wm;
  const uint64_t c_bm = bm[i] & all & ~(c_wm | c_wk), c_bk = bk[i] & all & ~(c_wm | c_wk | c_bm);
  double s = __dmul_rn((double)(__popcll(c_bm) - __popcll(c_wm)), 1.0);
  s = __dadd_rn(s, __dmul_rn((double)(__popcll(c_bk) - __popcll(c_wk)), 2.5));
  s = __dadd_rn(s, pst_sum<S>(s_pst, c_bm));
  s = __dsub_rn(s, pst_sum<S>(s_pst + S, c_wm));
  s = __dadd_rn(s, pst_sum<S>(s_pst + 2 * S, c_bk));
  s = __dsub_rn(s, pst_sum<S>(s_pst + 3 * S, c_wk));
  score[i] = (turn[i] & 1) == 0 ? -s : s;
}

// ---------------------------------------------------------------------------------------------
// legal_moves (+ mask + tensor) for a batch.
//
// What bounds it (tools/micro/emit_pattern.cu, profiles/README.md): the mask rows are 2.6 GB of zeros with ~8 set
// bytes per row.  Streaming full 16-byte stores reaches ~6.8 TB/s on B200, but every single-byte store into the
// rows is a partial-sector write that costs ~40 ns of chip-wide throughput (8 per board doubles the kernel's
// time).  So no byte ever goes to global memory on its own: each warp keeps ONE BIT per mask byte of its 32 rows
// in shared memory (80 000 bits), move emitters set bits there, and at the end the warp widens the bit string to
// bytes in registers and streams the rows out with coalesced 16-byte stores.  Records are stored with one 32-byte
// (v8.b32) store each, a full sector.
//   phase 0  clear the warp's bit string, load tables
//   phase A  thread = own board: quiet moves (records + bits); boards with captures go to a block list
//   phase B  capture boards re-dealt onto consecutive threads: single-jump boards finished bit-parallel
//   phase C  the rest: tree search with the explicit DFS stack in shared memory
//   phase D  warp-cooperative: rows widened from the bit string and streamed out; tensor planes as float4s
// ---------------------------------------------------------------------------------------------
// records a board may write: its slice [off, next) of the record array, clipped to the array's capacity
__device__ __forceinline__ uint64_t emit_room(uint64_t off, uint64_t next, uint64_t cap) {
  if (off >= cap) return 0;
  const uint64_t end = next < cap ? next : cap;
  return end > off ? end - off : 0;
}

struct EmitArgs {
  int64_t n;
  const uint64_t *wm, *wk, *bm, *bk;
  const uint8_t* turn;
  const uint64_t* offsets;
  DrgMove* moves;
  uint64_t moves_cap;
  uint8_t* mask;
  float* tensor;
  uint32_t* counts;
  unsigned long long* overflow;
};

// one full-sector store per record
__device__ __forceinline__ void rec_store32(DrgMove* dst, const Rec& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]),
               "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7])
               : "memory");
}

template <int S>
struct BitSink {  // writes records (at most `room`) and sets the board's mask bits in shared memory
  DrgMove* out;
  uint32_t* bits;    // the block's bit strings (NULL: no mask wanted)
  uint32_t bit0;     // first bit of this board's row inside `bits`
  uint32_t n, room;
  __device__ __forceinline__ void mark(int from, int to) {
    if (bits) {
      const uint32_t b = bit0 + (uint32_t)(from * S + to);
      atomicOr(&bits[b >> 5], 1u << (b & 31));
    }
  }
  __device__ __forceinline__ void quiet(int from, int to, bool king) {
    if (n < room) { Rec r; rec_quiet(r, from, to, king); rec_store32(out + n, r); }
    mark(from, to);
    n++;
  }
  template <class BB>
  __device__ __forceinline__ void capture(int nsq, BB capt, bool promo, bool root_king, int cur, const Frame* stack,
                                          int stride) {
    if (n < room) { Rec r; rec_capture(r, nsq, capt, promo, root_king, cur, stack, stride); rec_store32(out + n, r); }
    mark((int)stack[0].sq, cur);
    n++;
  }
  __device__ __forceinline__ void jump(int from, int victim, int land) {  // single man jump
    if (n < room) { Rec r; rec_jump(r, from, victim, land); rec_store32(out + n, r); }
    mark(from, land);
    n++;
  }
};

template <int S, bool RECORDS = true>
__device__ __forceinline__ BitSink<S> make_bit_sink(const EmitArgs& a, int64_t i, uint32_t* bits, int t) {
  BitSink<S> sink;
  // mask-only launches run before the offsets exist: no slice, no room, no record is stored
  const uint64_t off = RECORDS ? a.offsets[i] : 0;
  const uint64_t room64 = RECORDS ? emit_room(off, a.offsets[i + 1], a.moves_cap) : 0;
  sink.out = a.moves + off;
  sink.bits = bits;
  sink.bit0 = (uint32_t)t * (S * S);  // block-local board t: rows are contiguous, so are their bit strings
  sink.n = 0;
  sink.room = room64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)room64;
  return sink;
}

__device__ __forceinline__ uint32_t bytes_of_nibble(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

template <int S>
__host__ __device__ constexpr int emit_bits_words() { return kBlock * S * S / 32; }  // 10 000 (S=50) / 4 096 (S=32) words per block

// to_tensor(perspective = side to move) (base.py:753-805) of a warp's boards: 4 planes of S floats per board.  Each
// lane packs its board's planes into a 4*S-bit string (8 words of `bits`, 32 * 8 words per warp), then the warp
// streams float4s.
template <class R>
__device__ __forceinline__ void emit_tensor_rows(const Pos<R>& p, uint32_t* bits, float* out, int nboards, int lane) {
  constexpr int S = R::S;
  {
    const uint64_t c0 = p.own_men();
    const uint64_t c1 = p.own_kings() & ~c0;
    const uint64_t c2 = p.opp_men() & ~(c0 | c1);
    const uint64_t c3 = p.opp_kings() & ~(c0 | c1 | c2);
    uint64_t w0, w1, w2, w3;
    if (S == 50) {
      w0 = c0 | (c1 << 50);
      w1 = (c1 >> 14) | (c2 << 36);
      w2 = (c2 >> 28) | (c3 << 22);
      w3 = c3 >> 42;
    } else {
      w0 = c0 | (c1 << 32);
      w1 = c2 | (c3 << 32);
      w2 = w3 = 0;
    }
    uint32_t* b = bits + lane * 8;
    b[0] = (uint32_t)w0; b[1] = (uint32_t)(w0 >> 32); b[2] = (uint32_t)w1; b[3] = (uint32_t)(w1 >> 32);
    b[4] = (uint32_t)w2; b[5] = (uint32_t)(w2 >> 32); b[6] = (uint32_t)w3; b[7] = (uint32_t)(w3 >> 32);
  }
  __syncwarp();
  float4* dst = reinterpret_cast<float4*>(out);
  const int total4 = nboards * S;  // float4 chunks: 4*S floats per board
  for (int q = lane; q < total4; q += 32) {
    const int board = q / S, bit = (q - board * S) * 4;
    const uint32_t nib = (bits[board * 8 + (bit >> 5)] >> (bit & 31)) & 0xFu;
    __stcs(&dst[q], make_float4((nib & 1) ? 1.0f : 0.0f, (nib & 2) ? 1.0f : 0.0f, (nib & 4) ? 1.0f : 0.0f,
                                (nib & 8) ? 1.0f : 0.0f));
  }
}

// HAVE_LIST: the count step of this batch already built the tree-search list (and its statistics), nothing to append
// RECORDS = false: mask rows (and tensor planes) only -- needs neither counts nor offsets, so the fused call starts it
// at once on its own stream while the count -> scan -> records chain runs beside it (drg_abi.cu, drg_movegen)
template <class R, bool WITH_MASK, bool WITH_TENSOR, bool HAVE_LIST, bool RECORDS = true>
__global__ void __launch_bounds__(kBlock) k_emit(EmitArgs a, DeepList deep) {
  constexpr int S = R::S;
  constexpr int kRow = S * S;
  extern __shared__ __align__(16) uint8_t smem[];
  Tables<S>& T = *reinterpret_cast<Tables<S>*>(smem);
  CapList& s_cap = *reinterpret_cast<CapList*>(smem + sizeof(Tables<S>));
  uint32_t* s_n = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(&s_cap) + ((sizeof(CapList) + 15) & ~15));
  uint32_t* s_bits = s_n + kBlock;  // mask bit strings [kBlock * S*S bits] (WITH_MASK) / tensor bit strings
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_tables<S>(&T);
  // blocks loop over 128-board tiles; normally one tile per block, a smaller grid when the kernel has to leave room
  // for another one (the records chain of the fused call)
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < a.n; base += (int64_t)gridDim.x * kBlock) {
  if (threadIdx.x == 0) s_cap.n = s_cap.n_deep = 0;
  if (WITH_MASK) {
    uint4* z = reinterpret_cast<uint4*>(s_bits);
    for (int q = threadIdx.x; q < emit_bits_words<S>() / 4; q += kBlock) z[q] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  const int64_t wb0 = base + warp * 32;  // first board of this warp
  const int nboards = wb0 >= a.n ? 0 : ((a.n - wb0) < 32 ? (int)(a.n - wb0) : 32);
  uint32_t* mbits = WITH_MASK ? s_bits : nullptr;

  const int64_t i = base + threadIdx.x;
  uint32_t ovf = 0;
  Pos<R> p;
  p.wm = p.wk = p.bm = p.bk = 0;
  p.turn = 0;
  if (i < a.n) {  // phase A
    p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[i] & 1, i);
    BitSink<S> sink = make_bit_sink<S, RECORDS>(a, i, mbits, threadIdx.x);
    MenJumps<R> mj;
    const auto c = find_capturers<R>(p, T, mj);
    if (R::kOptionalCapture || !c.any()) gen_quiet<R, false>(p, sink, T);
    if (c.any()) {
      s_n[threadIdx.x] = sink.n;
      s_cap.item[atomicAdd(&s_cap.n, 1)] = (uint8_t)threadIdx.x;
    } else if (RECORDS) {
      if (a.counts) a.counts[i] = sink.n;
      if (sink.n > sink.room) ++ovf;
    }
  }
  __syncthreads();
  // phase B: capture boards re-dealt onto consecutive threads; single man jumps are finished bit-parallel, boards
  // that need the tree search are handed to k_emit_deep through a grid-wide list (32 divergent DFS walks per warp
  // would otherwise keep the whole block, and its 40 KB of bit strings, waiting)
  for (int k = threadIdx.x; k < s_cap.n; k += kBlock) {
    const int t = s_cap.item[k];
    const int64_t j = base + t;
    const Pos<R> q = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[j] & 1, j);
    MenJumps<R> mj;
    const auto c = find_capturers<R>(q, T, mj);
    if (!c.kings && only_single_jumps<R>(q, mj)) {
      BitSink<S> sink = make_bit_sink<S, RECORDS>(a, j, mbits, t);
      sink.n = s_n[t];
      single_jumps<R, true>(q, mj, sink, T);  // the sink counts the moves
      if (RECORDS) {
        if (a.counts) a.counts[j] = sink.n;
        if (sink.n > sink.room) ++ovf;
      }
    } else if (!HAVE_LIST) {
      list_append(deep, (uint32_t)j);
    }
  }
  __syncthreads();
  if (ovf && a.overflow) atomicAdd(a.overflow, (unsigned long long)ovf);  // one per clipped board

  if (WITH_MASK && nboards) {
    // phase D: legal_moves_mask rows (base.py:807-838) of this warp's boards are one contiguous, 16-byte aligned
    // range of the [n, S*S] array (32 * S*S is a multiple of 16).  16 bits of the warp's bit string -> 16 bytes.
    const uint16_t* wbits = reinterpret_cast<const uint16_t*>(s_bits) + warp * (32 * kRow / 16);
    uint8_t* dst = a.mask + (size_t)wb0 * kRow;
    const int bytes = nboards * kRow;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int q = lane; q < bytes / 16; q += 32) {
      const uint32_t h = wbits[q];
      const uint4 v = make_uint4(bytes_of_nibble(h & 15u), bytes_of_nibble((h >> 4) & 15u), bytes_of_nibble((h >> 8) & 15u),
                                 bytes_of_nibble((h >> 12) & 15u));
      // streaming store (evict-first): 2.6 GB of rows per Mi boards pass through the L2 once; with plain stores they evict
      // what the count -> records chain beside this kernel re-reads (fused step 0.563 -> 0.555 ms)
      __stcs(&d4[q], v);
    }
    const uint8_t* wnib = reinterpret_cast<const uint8_t*>(wbits);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(dst);
    for (int q = (bytes / 16) * 4 + lane; q < bytes / 4; q += 32) d1[q] = bytes_of_nibble((wnib[q >> 1] >> ((q & 1) * 4)) & 15u);
  }

  if (WITH_TENSOR && nboards) {
    // to_tensor(perspective = side to move) (base.py:753-805): 4 planes of S floats per board.
    // Each board's planes are packed into a 4*S-bit string (8 words), then the warp streams float4s.
    __syncthreads();  // the mask bit strings are dead from here on: their memory holds the tensor bit strings
    emit_tensor_rows<R>(p, s_bits + warp * 32 * 8, a.tensor + (size_t)wb0 * 4 * S, nboards, lane);
  }
  __syncthreads();  // the tile's shared memory is reused by the next one
  }
}

template <int S, bool WITH_MASK, bool WITH_TENSOR>
constexpr size_t emit_smem_bytes() {
  size_t b = sizeof(Tables<S>) + ((sizeof(CapList) + 15) & ~(size_t)15) + sizeof(uint32_t) * kBlock;
  const size_t mask_b = WITH_MASK ? (size_t)emit_bits_words<S>() * 4 : 0;
  const size_t tens_b = WITH_TENSOR ? (size_t)kBlock * 8 * 4 : 0;
  return b + (mask_b > tens_b ? mask_b : tens_b) + 16;
}

// tree-search boards of k_emit (7 % of random-play Standard positions), one per thread: records, counts and -- as
// single-byte stores into rows k_emit has already written -- the mask bytes of their capture moves
// Mask bytes either go straight into the row (the rows are written: k_emit ran before) or, when this kernel runs
// NEXT TO k_emit on a second stream, their addresses go to a list that k_mask_fixup stores after both have finished.
struct MaskFix {
  uint64_t* at;      // byte offsets into the mask array
  unsigned* n;       // entries appended; the list holds as many as the record buffer holds records, so it can only
                     // overflow for boards whose records were clipped as well (k_mask_fixup reports it)
  unsigned cap;
};

template <int S>
struct DeepSink {
  DrgMove* out;
  uint8_t* mrow;
  uint32_t n, room;
  uint64_t row0;     // offset of the row in the mask array (deferred mode)
  MaskFix fix;       // fix.at == nullptr: direct mode
  __device__ __forceinline__ void mark(int from, int to) {
    if (!mrow) return;
    if (fix.at) {
      // one atomicAdd per converged lane group (all entries go through ONE counter)
      const unsigned m = __activemask();
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      unsigned base = 0;
      if (lane == leader) base = atomicAdd(fix.n, (unsigned)__popc(m));
      base = __shfl_sync(m, base, leader);
      const unsigned k = base + __popc(m & ((1u << lane) - 1));
      if (k < fix.cap) fix.at[k] = row0 + (uint64_t)(from * S + to);
    } else {
      mrow[from * S + to] = 1;
    }
  }
  __device__ __forceinline__ void quiet(int, int, bool) { n++; }  // quiet moves of optional-capture variants: written by k_emit
  template <class BB>
  __device__ __forceinline__ void capture(int nsq, BB capt, bool promo, bool root_king, int cur, const Frame* stack,
                                          int stride) {
    if (n < room) { Rec r; rec_capture(r, nsq, capt, promo, root_king, cur, stack, stride); rec_store32(out + n, r); }
    mark((int)stack[0].sq, cur);
    n++;
  }
  __device__ __forceinline__ void jump(int from, int victim, int land) {
    if (n < room) { Rec r; rec_jump(r, from, victim, land); rec_store32(out + n, r); }
    mark(from, land);
    n++;
  }
};

// the deferred mask bytes of k_emit_deep
__global__ void __launch_bounds__(256) k_mask_fixup(uint8_t* mask, MaskFix fix, unsigned long long* overflow) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && *fix.n > fix.cap && overflow) atomicAdd(overflow, 1ull);
  const unsigned n = *fix.n < fix.cap ? *fix.n : fix.cap;
  for (unsigned k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) mask[fix.at[k]] = 1;
}

// fix.at != nullptr: deferred mask bytes (see MaskFix).
//
// Root walks of the tree-search boards (k_walk_root + k_walk_rounds ran before on the same list): the board's legal
// captures that its root walk found are copied from the pool slot where that walk kept them, or -- more than the slot
// holds -- found again by the same walk (same budget, so it stops where it stopped); what the walk handed over is
// k_emit_recs' part.
template <class R, bool WITH_MASK>
__global__ void __launch_bounds__(kBlock) k_emit_deep(EmitArgs a, DeepList deep, const uint32_t* __restrict__ stat,
                                                      MaskFix fix, const DrgMove* __restrict__ pool, unsigned pool_boards,
                                                      const WalkRes* __restrict__ rres,
                                                      const uint32_t* __restrict__ bword, WalkPool wp) {
  constexpr int S = R::S;
  __shared__ __align__(16) Tables<S> T;
  __shared__ Frame s_stack[kMaxLevels * kBlock];
  const unsigned nd = *deep.n;
  if ((unsigned long long)blockIdx.x * kBlock >= nd) return;
  load_tables<S>(&T);
  __syncthreads();
  NullSink ns;
  NoAlloc noalloc;
  uint32_t ovf = 0, clipped = 0;
  for (unsigned k = blockIdx.x * kBlock + threadIdx.x; k < nd; k += gridDim.x * kBlock) {
    const int64_t j = deep.items[k];
    const WalkRes res = rres[k];
    const uint32_t bw = bword[k];
    const uint32_t kept = walk_kept<R>(res, bw);
    const uint32_t sw = stat[k];
    const bool copy = k < pool_boards && (sw & 0x40000000u);  // the root walk's best leaves are all in the pool slot
    Pos<R> p;
    p.wm = p.wk = p.bm = p.bk = 0;
    p.turn = 0;
    if ((kept && !copy) || R::kOptionalCapture) p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[j] & 1, j);
    DeepSink<S> sink;
    const uint64_t off = a.offsets[j];
    const uint64_t room64 = emit_room(off, a.offsets[j + 1], a.moves_cap);
    sink.out = a.moves + off;
    sink.mrow = WITH_MASK ? a.mask + (size_t)j * (S * S) : nullptr;
    sink.row0 = (uint64_t)j * (S * S);
    sink.fix = fix;
    sink.room = room64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)room64;
    sink.n = R::kOptionalCapture ? (uint32_t)gen_quiet<R, true>(p, ns, T) : 0u;  // quiet moves come first (american.py:96)
    uint32_t total = sink.n + kept;  // all moves of the board: + what is below the root walk's records
    for (uint32_t c = 0, nc = res.info >> 24; c < nc; c++) total += wp.sub[res.child0 + c];
    if (kept && copy) {
      const int n_kept = (int)((sw >> 20) & 0xFFu);
      const bool kings_only = R::kFilter == FILTER_VALUE && (bw & 1u);  // frisian.py:264-268
      const uint4* src = reinterpret_cast<const uint4*>(pool + (size_t)k * kKeepRecords);
      for (int r = 0; r < n_kept; r++) {
        const uint4 lo = src[2 * r], hi = src[2 * r + 1];
        Rec rec;
        rec.w[0] = lo.x; rec.w[1] = lo.y; rec.w[2] = lo.z; rec.w[3] = lo.w;
        rec.w[4] = hi.x; rec.w[5] = hi.y; rec.w[6] = hi.z; rec.w[7] = hi.w;
        if (kings_only && !((rec.w[2] >> 8) & DRG_MF_KINGMOVE)) continue;
        if (sink.n < sink.room) rec_store32(sink.out + sink.n, rec);
        sink.mark(rec_from(rec), rec_to(rec));
        sink.n++;
      }
    } else if (kept) {
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      WalkLane<R, 1> lane;
      lane.start_root(p, c, k, (res.info & kWalkSplit) ? root_budget<R>() : kNoBudget, T);
      lane.st.best = (int)(bw >> 1);
      lane.st.n_best_king = (int)(bw & 1u);
      lane.run(sink, s_stack + threadIdx.x, kBlock, T, ovf, noalloc);
    }
    if (total > sink.room) ++clipped;
  }
  if ((ovf || clipped) && a.overflow) atomicAdd(a.overflow, (unsigned long long)(clipped ? clipped : 1u));
}

// the records' part of the emit step: every record whose walk found legal leaves is walked again (the budget it
// had, so it stops where it stopped) and writes them at base[w] of its board's capture moves
template <class R, bool WITH_MASK>
__global__ void __launch_bounds__(kBlock) k_emit_recs(EmitArgs a, DeepList deep, MaskFix fix,
                                                      const uint32_t* __restrict__ bword, WalkPool wp, unsigned* cursor) {
  constexpr int S = R::S;
  if (walk_round_size(wp, 1) == 0) return;  // nothing was handed over
  __shared__ __align__(16) Tables<S> T;
  __shared__ Frame s_stack[kMaxLevels * kBlock];
  load_tables<S>(&T);
  __syncthreads();
  unsigned nrec = 0;
  for (int r = 1; r < kWalkRounds; r++) {
    const unsigned c = walk_round_size(wp, r);
    if (!c) break;
    nrec += c;
  }
  NullSink ns;
  NoAlloc noalloc;
  uint32_t ovf = 0;
  WalkLane<R, 1> lane;
  DeepSink<S> sink;
  sink.out = nullptr; sink.mrow = nullptr; sink.n = sink.room = 0; sink.row0 = 0; sink.fix = fix;
  bool active = false, more = true;
  Frame* stack = s_stack + threadIdx.x;
  const unsigned takers = walk_lane_mask(nrec, gridDim.x * (kBlock / 32));
  const bool taker = (takers >> (threadIdx.x & 31)) & 1u;
  for (;;) {
    __syncwarp();
    const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active && taker);
    if (more && (idle == takers || __popc(idle) >= kWalkRefill)) {
      const unsigned w = walk_fetch(cursor, idle, !active && taker);
      more = __any_sync(0xFFFFFFFFu, !active && taker && w < nrec);
      if (!active && taker && w < nrec) {
        const WalkRec rec = load_rec(wp.recs + w);
        const WalkRes res = wp.res[w];
        const uint32_t bw = bword[rec.item];
        if (walk_kept<R>(res, bw)) {
          const int64_t j = deep.items[rec.item];
          const Pos<R> p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[j] & 1, j);
          const uint64_t off = a.offsets[j];
          const uint64_t room64 = emit_room(off, a.offsets[j + 1], a.moves_cap);
          sink.out = a.moves + off;
          sink.mrow = WITH_MASK ? a.mask + (size_t)j * (S * S) : nullptr;
          sink.row0 = (uint64_t)j * (S * S);
          sink.room = room64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)room64;
          sink.n = (R::kOptionalCapture ? (uint32_t)gen_quiet<R, true>(p, ns, T) : 0u) + wp.base[w];
          lane.start_rec(p, rec, (res.info & kWalkSplit) ? (int)wp.budget[rec.pad] : kNoBudget, stack, kBlock, T);
          lane.st.best = (int)(bw >> 1);
          lane.st.n_best_king = (int)(bw & 1u);
          active = true;
        }
      }
    }
    if (!more && !__any_sync(0xFFFFFFFFu, active)) break;
    if (active) {
      const int s = lane.advance(sink, stack, kBlock, T, ovf);
      if (s == kWalkDone || (s == kWalkWantsSplit && lane.hand_over(stack, kBlock, noalloc, T))) active = false;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Small batches (n <= kBlock boards; the batch-of-1 behind Board.legal_moves / legal_moves_mask / to_tensor,
// standard.py:99-105, base.py:753-838): count -> scan -> emit -> mask -> tensor in ONE block and ONE launch.  The
// batched pipeline is seven dependent launches, which is all latency at this size.  Inputs are read straight from
// the library's mapped, pinned staging words and results are written back into them (drg_abi.cu small_path).
// ---------------------------------------------------------------------------------------------
struct SmallArgs {
  int n, emit;
  const uint64_t *wm, *wk, *bm, *bk;
  const uint8_t* turn;
  uint32_t* counts;
  uint64_t* offsets;
  DrgMove* moves;
  uint32_t moves_cap;
  uint8_t* mask;               // [n, S*S] or NULL, 16-byte aligned
  float* tensor;               // [n, 4, S] or NULL
  unsigned long long* info;    // [0] boards with a capture path longer than a record, [1] total moves
};

template <class R>
__global__ void __launch_bounds__(kBlock) k_small(SmallArgs a) {
  constexpr int S = R::S;
  __shared__ __align__(16) Tables<S> T;
  __shared__ Frame s_stack[kMaxLevels * kBlock];
  __shared__ uint32_t s_cnt[kBlock];
  load_tables<S>(&T);
  const int i = threadIdx.x;
  const bool active = i < a.n;
  if (i == 0) a.info[0] = 0;
  Pos<R> p;
  p.wm = p.wk = p.bm = p.bk = 0;
  p.turn = 0;
  if (active) p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[i] & 1, i);
  if (a.emit && a.mask) {  // rows start as zeros (whole block, 16-byte stores)
    uint4* z = reinterpret_cast<uint4*>(a.mask);
    const int n16 = (a.n * S * S + 15) / 16;
    for (int k = i; k < n16; k += kBlock) z[k] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  uint32_t ovf = 0, stat = kNoStat;
  const uint32_t cnt = active ? (uint32_t)count_moves<R>(p, s_stack + i, kBlock, T, ovf, stat) : 0u;
  s_cnt[i] = cnt;
  __syncthreads();
  uint32_t off = 0;
  for (int k = 0; k < i; k++) off += s_cnt[k];
  if (active) {
    a.counts[i] = cnt;
    a.offsets[i] = off;
    if (i == a.n - 1) {
      a.offsets[a.n] = off + cnt;
      a.info[1] = off + cnt;
    }
  }
  if (a.emit && active) {
    EmitSink<S> sink;
    sink.out = a.moves + off;
    sink.mrow = a.mask ? a.mask + (size_t)i * (S * S) : nullptr;
    sink.n = 0;
    sink.room = off >= a.moves_cap ? 0u : (a.moves_cap - off < cnt ? a.moves_cap - off : cnt);
    if (cnt) generate_moves<R>(p, sink, s_stack + i, kBlock, T, ovf, stat);
    if (a.tensor) {  // to_tensor(perspective = side to move), base.py:753-805; overlapping cells: first plane wins
      const uint64_t c0 = p.own_men();
      const uint64_t c1 = p.own_kings() & ~c0;
      const uint64_t c2 = p.opp_men() & ~(c0 | c1);
      const uint64_t c3 = p.opp_kings() & ~(c0 | c1 | c2);
      float* t = a.tensor + (size_t)i * 4 * S;
      for (int sq = 0; sq < S; sq++) {
        t[sq] = (c0 >> sq & 1) ? 1.0f : 0.0f;
        t[S + sq] = (c1 >> sq & 1) ? 1.0f : 0.0f;
        t[2 * S + sq] = (c2 >> sq & 1) ? 1.0f : 0.0f;
        t[3 * S + sq] = (c3 >> sq & 1) ? 1.0f : 0.0f;
      }
    }
  }
  if (ovf) atomicAdd(a.info, 1ull);
}

// ---------------------------------------------------------------------------------------------
// random self-play (include/drg.h drg_rollout)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

enum DrawFamily { DRAW_STANDARD = 0, DRAW_AMERICAN = 1, DRAW_FRISIAN = 2, DRAW_RUSSIAN = 3 };

// per-variant is_draw without the threefold term: standard.py:276-311, american.py:245-256,
// russian.py:342-379 (the 3-kings rule needs clock >= 30 as well, so clock >= 30 subsumes it), frisian.py:618-659
template <class R>
__device__ __forceinline__ bool is_draw_no3(const Pos<R>& p, int clock, int fam) {
  const int all = bb_popc(p.occ());
  const int kings = bb_popc(typename Pos<R>::BB(p.wk | p.bk)), men = bb_popc(typename Pos<R>::BB(p.wm | p.bm));
  if (fam == DRAW_STANDARD) {
    if (clock >= 50) return true;
    if (all <= 3 && kings * 2 + men >= 5 && clock >= 10) return true;
    if (clock >= 32 && all <= 4 && kings * 2 + men >= 6) return true;
    return false;
  }
  if (fam == DRAW_AMERICAN) return clock >= 40;
  if (fam == DRAW_RUSSIAN) return clock >= 30;
  if (men == 0 && kings == 2 && bb_popc(p.wk) == 1 && bb_popc(p.bk) == 1 && clock >= 4) return true;
  if (clock >= 14 && men == 0 && kings == 3) return true;
  return clock >= 50;
}

template <class R>
__global__ void __launch_bounds__(kBlock)
k_rollout(int variant, int draw_fam, int64_t n, uint64_t seed, uint64_t game_id0, int max_plies,
          const uint16_t* __restrict__ stop_ply, uint32_t flags, uint64_t* wm, uint64_t* wk, uint64_t* bm,
          uint64_t* bk, uint8_t* turn, uint8_t* clock, uint8_t* result, uint16_t* plies,
          unsigned long long* counters) {
  __shared__ __align__(16) Tables<R::S> T;
  __shared__ Frame s_stack[kMaxLevels * kBlock];
  load_tables<R::S>(&T);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn[i] & 1, i);
  int clk = clock[i];
  const int limit = stop_ply ? (int)stop_ply[i] : max_plies;
  const uint64_t gid = game_id0 + (uint64_t)i;
  // last 9 pushed moves: words 2..7 of the record hold n_sq, flags, path (base.py:355-366 compares square_list)
  uint32_t hist[9][6];
  int nh = 0, ply = 0, res = 0;
  uint32_t ovf = 0, movegen = 0;
  Frame* stack = s_stack + threadIdx.x;
  for (;;) {
    if (flags & DRG_RF_DRAW_RULES) {
      bool draw = is_draw_no3<R>(p, clk, draw_fam);
      if (!draw && nh >= 9) {
        const uint32_t* a = hist[(nh - 1) % 9];
        const uint32_t* b = hist[(nh - 5) % 9];
        const uint32_t* c = hist[(nh - 9) % 9];
        bool same = true;
        for (int k = 0; k < 6; k++) {
          const uint32_t m = k == 0 ? 0xFFFF00FFu : 0xFFFFFFFFu;  // ignore the flags byte
          same = same && ((a[k] ^ b[k]) & m) == 0 && ((a[k] ^ c[k]) & m) == 0;
        }
        draw = same;
      }
      if (variant == DRG_BREAKTHROUGH && (p.wk | p.bk)) { res = p.wk ? 1 : 2; break; }  // breakthrough.py:28-39
      if (draw) { res = 3; break; }
    }
    uint32_t stat;
    const int nm = count_moves<R>(p, stack, kBlock, T, ovf, stat);
    movegen++;
    if (nm == 0) {
      // base.py:391-392; antidraughts.py:35-36 inverts
      if (variant == DRG_ANTIDRAUGHTS) res = p.turn == 0 ? 1 : 2;
      else res = p.turn == 0 ? 2 : 1;
      break;
    }
    if (ply >= limit) { res = 0; break; }
    const uint64_t x = splitmix64(seed + gid * 0x9E3779B97F4A7C15ull + (uint64_t)ply * 0xBF58476D1CE4E5B9ull);
    const uint32_t idx = (uint32_t)(((x >> 32) * (uint64_t)nm) >> 32);
    SelectSink sel;
    sel.n = 0;
    sel.want = idx;
    rec_quiet(sel.rec, 0, 0, false);
    generate_moves<R>(p, sel, stack, kBlock, T, ovf, stat);
    const uint64_t capt = (uint64_t)sel.rec.w[0] | ((uint64_t)sel.rec.w[1] << 32);
    if (!apply_move<R>(p, clk, rec_from(sel.rec), rec_to(sel.rec), typename Pos<R>::BB(capt),
                       ((sel.rec.w[2] >> 8) & DRG_MF_PROMO) != 0)) {
      if (counters) atomicAdd(&counters[DRG_C_ILLEGAL], 1ull);
      res = 0;
      break;
    }
    uint32_t* h = hist[nh % 9];
    for (int k = 0; k < 6; k++) h[k] = sel.rec.w[2 + k];
    nh++;
    ply++;
  }
  wm[i] = p.wm; wk[i] = p.wk; bm[i] = p.bm; bk[i] = p.bk;
  turn[i] = (uint8_t)p.turn;
  clock[i] = (uint8_t)(clk > 255 ? 255 : clk);
  if (result) result[i] = (uint8_t)res;
  if (plies) plies[i] = (uint16_t)ply;
  if (counters) {
    atomicAdd(&counters[DRG_C_NODES], (unsigned long long)ply);
    atomicAdd(&counters[DRG_C_MOVEGEN], (unsigned long long)movegen);
    atomicAdd(&counters[res == 1 ? DRG_C_WHITE_WINS : res == 2 ? DRG_C_BLACK_WINS : res == 3 ? DRG_C_DRAWS
                                                                                                : DRG_C_UNFINISHED],
              1ull);
    if (ovf) atomicAdd(&counters[DRG_C_OVERFLOW], 1ull);
  }
}

// one ply of self-play for every running game, consuming the move lists k_emit just produced for the
// current positions (so tensor / mask exports of the same launch describe exactly the position the move is
// chosen from -- the examples/reinforcement_learning.py:103-112 loop).  Same order of checks and same RNG
// as k_rollout: is_draw / Breakthrough kings -> no legal move -> ply limit -> play legal_moves[idx].
// state: 0 running, 1 '1-0', 2 '0-1', 3 '1/2-1/2', 4 stopped at the ply limit.
// hist : uint32[9*6][n], words 2..7 of the last nine pushed records (base.py:355-366).
template <class R>
__global__ void __launch_bounds__(kBlock)
k_selfplay_step(int variant, int draw_fam, int64_t n, uint64_t seed, uint64_t game_id0, int max_plies,
                const uint16_t* __restrict__ stop_ply, uint32_t flags, uint64_t* wm, uint64_t* wk, uint64_t* bm,
                uint64_t* bk, uint8_t* turn, uint8_t* clock, uint16_t* plies, uint8_t* state, uint32_t* hist,
                const uint32_t* __restrict__ counts, const uint64_t* __restrict__ offsets,
                const DrgMove* __restrict__ moves, uint64_t moves_cap, const int32_t* __restrict__ choice,
                const uint64_t* __restrict__ game_ids, unsigned long long* running, unsigned long long* illegal) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n || state[i] != 0) return;
  Pos<R> p = load_pos<R>(wm, wk, bm, bk, turn[i] & 1, i);
  int clk = clock[i];
  const int ply = plies[i];
  if (flags & DRG_RF_DRAW_RULES) {
    bool draw = is_draw_no3<R>(p, clk, draw_fam);
    if (!draw && ply >= 9) {
      bool same = true;
      const int a = (ply - 1) % 9, b = (ply - 5) % 9, c = (ply - 9) % 9;
      for (int k = 0; k < 6; k++) {
        const uint32_t m = k == 0 ? 0xFFFF00FFu : 0xFFFFFFFFu;  // ignore the flags byte
        const uint32_t va = hist[(size_t)(a * 6 + k) * n + i];
        same = same && ((va ^ hist[(size_t)(b * 6 + k) * n + i]) & m) == 0 &&
               ((va ^ hist[(size_t)(c * 6 + k) * n + i]) & m) == 0;
      }
      draw = same;
    }
    if (variant == DRG_BREAKTHROUGH && (p.wk | p.bk)) { state[i] = p.wk ? 1 : 2; return; }
    if (draw) { state[i] = 3; return; }
  }
  const uint32_t nm = counts[i];
  if (nm == 0) {
    if (variant == DRG_ANTIDRAUGHTS) state[i] = p.turn == 0 ? 1 : 2;
    else state[i] = p.turn == 0 ? 2 : 1;
    return;
  }
  const int limit = stop_ply ? (int)stop_ply[i] : max_plies;
  if (ply >= limit) { state[i] = 4; return; }
  uint32_t idx;
  if (choice) {  // caller-chosen move index (policy sampling); out of range falls back to move 0
    idx = (uint32_t)choice[i] < nm ? (uint32_t)choice[i] : 0u;
  } else {
    const uint64_t gid = game_ids ? game_ids[i] : game_id0 + (uint64_t)i;
    const uint64_t x = splitmix64(seed + gid * 0x9E3779B97F4A7C15ull + (uint64_t)ply * 0xBF58476D1CE4E5B9ull);
    idx = (uint32_t)(((x >> 32) * (uint64_t)nm) >> 32);
  }
  if (offsets[i] + idx >= moves_cap) {  // the record was clipped by the buffer's capacity: stop the game, report it
    atomicAdd(illegal, 1ull);
    state[i] = 4;
    return;
  }
  const uint4* src = reinterpret_cast<const uint4*>(moves + offsets[i] + idx);
  const uint4 ra = src[0], rb = src[1];
  Rec r;
  r.w[0] = ra.x; r.w[1] = ra.y; r.w[2] = ra.z; r.w[3] = ra.w;
  r.w[4] = rb.x; r.w[5] = rb.y; r.w[6] = rb.z; r.w[7] = rb.w;
  const uint64_t capt = (uint64_t)r.w[0] | ((uint64_t)r.w[1] << 32);
  if (!apply_move<R>(p, clk, rec_from(r), rec_to(r), typename Pos<R>::BB(capt), ((r.w[2] >> 8) & DRG_MF_PROMO) != 0)) {
    atomicAdd(illegal, 1ull);
    state[i] = 4;
    return;
  }
  wm[i] = p.wm; wk[i] = p.wk; bm[i] = p.bm; bk[i] = p.bk;
  turn[i] = (uint8_t)p.turn;
  clock[i] = (uint8_t)(clk > 255 ? 255 : clk);
  const int slot = ply % 9;
  for (int k = 0; k < 6; k++) hist[(size_t)(slot * 6 + k) * n + i] = r.w[2 + k];
  plies[i] = (uint16_t)(ply + 1);
  atomicAdd(running, 1ull);
}

// ---------------------------------------------------------------------------------------------
// Policy head -> move (the consumer loop of examples/reinforcement_learning.py:103-112: to_tensor -> logits ->
// logits[~mask] = -1e9 -> softmax(logits / temp) -> multinomial -> index_to_move -> push), on the device.
//
// The legal actions of a board are the distinct from*S+to of its move records (exactly the set bytes of its
// legal_moves_mask row, base.py:833-837), so the masked softmax is taken over the records instead of over the S*S-wide
// row: ~8 records of 32 bytes per board instead of 2 500 mask bytes + 10 000 logit bytes.  One warp per board.  An
// action's record is the FIRST one with that (from, to) in canonical order -- what index_to_move returns
// (base.py:861-891); different capture routes between the same two squares are one action.
// mode 0: sample (counter RNG keyed like the rollouts: seed, game id, ply), mode 1: argmax (ties: lowest action).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int rec_action(const DrgMove* rec, int S, uint32_t w2lo_unused = 0) {
  (void)w2lo_unused;
  const uint4* src = reinterpret_cast<const uint4*>(rec);
  const uint4 a = src[0], b = src[1];
  Rec r;
  r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
  r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
  int nsq = rec_nsq(r);
  if (nsq < 1) return -1;
  if (nsq > DRG_MAX_PATH) r.w[2] = (r.w[2] & ~0xFFu) | (uint32_t)DRG_MAX_PATH;
  return rec_from(r) * S + rec_to(r);
}

template <int S>
__global__ void __launch_bounds__(kBlock) k_sample_action(int64_t n, const float* __restrict__ logits,
                                                          const uint32_t* __restrict__ counts,
                                                          const uint64_t* __restrict__ offsets,
                                                          const DrgMove* __restrict__ moves, uint64_t moves_cap,
                                                          float inv_temp, int mode, uint64_t seed, uint64_t game_id0,
                                                          const uint64_t* __restrict__ game_ids,
                                                          const uint16_t* __restrict__ plies, int32_t* __restrict__ action,
                                                          int32_t* __restrict__ choice) {
  constexpr int kWords = (S * S + 31) / 32;
  __shared__ uint32_t s_seen[kWarps][kWords];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kWarps + warp;
  if (i >= n) return;  // the whole warp
  const uint64_t off = offsets[i];
  uint32_t m = counts[i];
  if (off >= moves_cap) m = 0;
  else if (off + m > moves_cap) m = (uint32_t)(moves_cap - off);  // records clipped by the buffer's capacity
  if (m == 0) {
    if (lane == 0) { action[i] = -1; choice[i] = -1; }
    return;
  }
  const float* lg = logits + (size_t)i * (S * S);
  const DrgMove* rec = moves + off;
  uint32_t* seen = s_seen[warp];
  // one pass over the records: every lane learns whether its record is the first with its action
  auto first_of_chunk = [&](uint32_t base, int& a) {
    const uint32_t k = base + lane;
    a = k < m ? rec_action(rec + k, S) : -1;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, a >= 0 ? a : -1 - lane);
    bool first = a >= 0 && lane == __ffs(peers) - 1;
    if (first) first = !(atomicOr(&seen[a >> 5], 1u << (a & 31)) & (1u << (a & 31)));
    __syncwarp();
    return first;
  };
  auto clear_seen = [&]() {
    for (int q = lane; q < kWords; q += 32) seen[q] = 0;
    __syncwarp();
  };
  // pass 1: the largest scaled logit (and, for argmax, where)
  clear_seen();
  float best = -INFINITY;
  int best_a = 0x7FFFFFFF, best_k = -1;
  for (uint32_t base = 0; base < m; base += 32) {
    int a;
    if (first_of_chunk(base, a)) {
      const float l = lg[a] * inv_temp;
      if (l > best || (l == best && a < best_a) || best_k < 0) { best = l; best_a = a; best_k = (int)(base + lane); }
    }
  }
  for (int o = 16; o; o >>= 1) {
    const float ol = __shfl_xor_sync(0xFFFFFFFFu, best, o);
    const int oa = __shfl_xor_sync(0xFFFFFFFFu, best_a, o), ok = __shfl_xor_sync(0xFFFFFFFFu, best_k, o);
    if (ok >= 0 && (best_k < 0 || ol > best || (ol == best && oa < best_a))) { best = ol; best_a = oa; best_k = ok; }
  }
  if (mode == 1) {
    if (lane == 0) { action[i] = best_a; choice[i] = best_k; }
    return;
  }
  // pass 2: sum of exp(l - max) over the actions, in record order
  clear_seen();
  float total = 0.0f;
  for (uint32_t base = 0; base < m; base += 32) {
    int a;
    const bool first = first_of_chunk(base, a);
    float e = first ? __expf(lg[a] * inv_temp - best) : 0.0f;
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xFFFFFFFFu, e, o);
      if (lane >= o) e += t;
    }
    total += __shfl_sync(0xFFFFFFFFu, e, 31);
  }
  // pass 3: the action whose cumulative weight first reaches u * total
  const uint64_t gid = game_ids ? game_ids[i] : game_id0 + (uint64_t)i;
  const uint64_t x = splitmix64_hash((seed ^ 0x5851F42D4C957F2Dull) + gid * 0x9E3779B97F4A7C15ull +
                                     (uint64_t)(plies ? plies[i] : 0) * 0xBF58476D1CE4E5B9ull);
  const float target = (float)(x >> 40) * (1.0f / 16777216.0f) * total;
  clear_seen();
  float carry = 0.0f;
  int pick_a = -1, pick_k = -1, last_a = -1, last_k = -1;
  for (uint32_t base = 0; base < m && pick_k < 0; base += 32) {
    int a;
    const bool first = first_of_chunk(base, a);
    const float w = first ? __expf(lg[a] * inv_temp - best) : 0.0f;
    float e = w;
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xFFFFFFFFu, e, o);
      if (lane >= o) e += t;
    }
    const unsigned hit = __ballot_sync(0xFFFFFFFFu, first && carry + e > target);
    const unsigned firsts = __ballot_sync(0xFFFFFFFFu, first);
    if (hit) {
      const int src = __ffs(hit) - 1;
      pick_a = __shfl_sync(0xFFFFFFFFu, a, src);
      pick_k = (int)base + src;
    }
    if (firsts) {
      const int src = 31 - __clz(firsts);
      last_a = __shfl_sync(0xFFFFFFFFu, a, src);
      last_k = (int)base + src;
    }
    carry += __shfl_sync(0xFFFFFFFFu, e, 31);
  }
  if (pick_k < 0) { pick_a = last_a; pick_k = last_k; }  // rounding left the target just above the last cumulative weight
  if (lane == 0) { action[i] = pick_a; choice[i] = pick_k; }
}

// index_to_move (base.py:861-891) for a batch: choice[i] = index, in board i's move list, of the first record whose
// (from, to) is action[i]; -1 where the reference raises ValueError (no legal move with these squares)
template <int S>
__global__ void __launch_bounds__(kBlock) k_action_to_choice(int64_t n, const int32_t* __restrict__ action,
                                                             const uint32_t* __restrict__ counts,
                                                             const uint64_t* __restrict__ offsets,
                                                             const DrgMove* __restrict__ moves, uint64_t moves_cap,
                                                             int32_t* __restrict__ choice) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const uint64_t off = offsets[i];
  uint32_t m = counts[i];
  if (off >= moves_cap) m = 0;
  else if (off + m > moves_cap) m = (uint32_t)(moves_cap - off);
  const int want = action[i];
  int found = -1;
  for (uint32_t k = 0; k < m && found < 0; k++)
    if (rec_action(moves + off + k, S) == want) found = (int)k;
  choice[i] = found;
}

// ---------------------------------------------------------------------------------------------
// packed transport of mask rows for the host-buffer path: 32 mask bytes (each 0 / 1) -> one uint32 of bits,
// flat over the whole [n, S*S] array (bit j of the stream = byte j of the mask).  drg_host.cpp widens it back.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nibble_of_bytes(uint32_t x) { return ((x * 0x01020408u) >> 24) & 0xFu; }

__global__ void __launch_bounds__(256) k_pack_mask(int64_t nbytes, const uint8_t* __restrict__ mask,
                                                   uint32_t* __restrict__ bits) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b0 = t * 32;
  if (b0 >= nbytes) return;
  uint32_t r = 0;
  if (b0 + 32 <= nbytes) {
    const uint4* p = reinterpret_cast<const uint4*>(mask + b0);
    const uint4 a = p[0], b = p[1];
    r = nibble_of_bytes(a.x) | (nibble_of_bytes(a.y) << 4) | (nibble_of_bytes(a.z) << 8) | (nibble_of_bytes(a.w) << 12) |
        (nibble_of_bytes(b.x) << 16) | (nibble_of_bytes(b.y) << 20) | (nibble_of_bytes(b.z) << 24) |
        (nibble_of_bytes(b.w) << 28);
  } else {
    for (int k = 0; b0 + k < nbytes; k++) r |= (uint32_t)(mask[b0 + k] & 1) << k;
  }
  bits[t] = r;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of counts (uint32) into offsets (uint64), offsets[n] = total
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Self-play working set: finished games out to the result arrays, running games packed to the front (drg_selfplay_compact)
// ---------------------------------------------------------------------------------------------
struct GameSet {
  uint64_t *wm, *wk, *bm, *bk;
  uint8_t *turn, *clock;
  uint16_t* plies;
  uint8_t* state;
  uint32_t* hist;
  int64_t* ids;
  uint64_t* game_ids;
};

__global__ void __launch_bounds__(256) k_compact_flags(int64_t n, const uint8_t* __restrict__ state, uint32_t* __restrict__ keep) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) keep[i] = state[i] == 0 ? 1u : 0u;
}

// pos: exclusive scan of the flags, pos[n] = running games; pack == false: every game goes to `fin`
__global__ void __launch_bounds__(256) k_compact_games(int64_t n, GameSet src, GameSet dst, GameSet fin, bool pack,
                                                       const uint64_t* __restrict__ pos) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint8_t s = src.state[i];
  const uint64_t wm = src.wm[i], wk = src.wk[i], bm = src.bm[i], bk = src.bk[i];
  const uint8_t turn = src.turn[i], clock = src.clock[i];
  const uint16_t plies = src.plies[i];
  const int64_t id = src.ids[i];
  if (s != 0 || !pack) {
    fin.wm[id] = wm; fin.wk[id] = wk; fin.bm[id] = bm; fin.bk[id] = bk;
    fin.turn[id] = turn; fin.clock[id] = clock; fin.plies[id] = plies; fin.state[id] = s;
  }
  if (s == 0 && pack) {
    const uint64_t j = pos[i], m = pos[n];
    dst.wm[j] = wm; dst.wk[j] = wk; dst.bm[j] = bm; dst.bk[j] = bk;
    dst.turn[j] = turn; dst.clock[j] = clock; dst.plies[j] = plies; dst.state[j] = 0;
    dst.ids[j] = id;
    dst.game_ids[j] = src.game_ids[i];
#pragma unroll 6
    for (int k = 0; k < 54; k++) dst.hist[(size_t)k * m + j] = src.hist[(size_t)k * n + i];
  }
}

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* s_warp,
                                                                  unsigned long long& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
    unsigned long long winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (int)(blockDim.x >> 5)) s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  block_total = s_warp[32];
  const unsigned long long r = s_warp[warp] + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanBlock) k_scan_tiles(int64_t n, const uint32_t* __restrict__ counts,
                                                           unsigned long long* __restrict__ tile_sums) {
  __shared__ unsigned long long s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  unsigned long long v = 0;
  for (int k = 0; k < kScanItems; k++)
    if (base + k < n) v += counts[base + k];
  unsigned long long total;
  block_exclusive_scan(v, s_warp, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock) k_scan_sums(int64_t ntiles, unsigned long long* tile_sums) {
  __shared__ unsigned long long s_warp[33];
  unsigned long long carry = 0;
  for (int64_t base = 0; base < ntiles; base += kScanBlock) {
    const int64_t j = base + threadIdx.x;
    const unsigned long long v = j < ntiles ? tile_sums[j] : 0;
    unsigned long long total;
    const unsigned long long ex = block_exclusive_scan(v, s_warp, total);
    if (j < ntiles) tile_sums[j] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}

__global__ void __launch_bounds__(kScanBlock) k_scan_apply(int64_t n, const uint32_t* __restrict__ counts,
                                                           const unsigned long long* __restrict__ tile_sums,
                                                           uint64_t* __restrict__ offsets, int64_t ntiles) {
  __shared__ unsigned long long s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t c[kScanItems];
  unsigned long long v = 0;
  for (int k = 0; k < kScanItems; k++) {
    c[k] = base + k < n ? counts[base + k] : 0;
    v += c[k];
  }
  unsigned long long total;
  unsigned long long ex = block_exclusive_scan(v, s_warp, total) + tile_sums[blockIdx.x];
  for (int k = 0; k < kScanItems; k++) {
    if (base + k < n) offsets[base + k] = ex;
    ex += c[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) offsets[n] = tile_sums[ntiles];
}

}  // namespace drg
