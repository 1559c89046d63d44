// drg_fused.cuh -- the fused call (drg_movegen: counts, offsets, records, mask rows, tensor planes) as ONE kernel.
//
// Replaces, for ordinary batches, the split pipeline of drg_kernels.cuh (k_emit<rows> beside k_count -> k_walk_root ->
// scan x3 -> k_emit<records> -> k_emit_deep -> k_mask_fixup: the moves were enumerated three times and the chain's wide
// kernels competed with the 2.6 GB row writer for the SMs, one kernel-level dependency after the other).
//
// Blocks are numbered by an atomic ticket b and have four TILE warps and kScoutWarps SCOUT warps:
//
//   scout warps  do everything that is latency, not bytes, well before the bytes are due.
//                step 1, tile b: exact capture pre-check + popc counts (phase A), single-jump boards re-dealt and counted
//                bit-parallel (B), the tile's tree-search boards (7 % of random-play Standard positions) walked on a few
//                lanes of every scout warp (C; statistics pass that keeps the best leaves in a small pool).  Published:
//                counts, board classes, kept leaves, and the tile's total in its look-back word.  Step 1 waits for nobody.
//                step 2, tile b - lag: the records before that tile -- a look-back over the words of the tiles before it
//                (totals, or inclusive prefixes where an earlier step 2 got that far), published as the tile's ready
//                word.  It only waits when a tile before it is still being walked `lag` blocks later.
//   tile warps   write tile b - ahead: wait for its ready word (published ahead - lag blocks earlier), read counts and
//                classes, scan 128 counts, then enumerate every move ONCE, at its final place: records as 32-byte
//                full-sector stores, mask bits into the tile's bit strings in shared memory; tree-search boards copy
//                their kept leaves from the pool.  Each warp then widens its bit string to bytes and streams its 32
//                rows with coalesced 16-byte stores; tensor planes as float4s.
//
// The grid has tiles + ahead blocks; `ahead` is about the number of blocks the GPU holds at once, so a tile is scouted
// one wave before it is written.  The first `ahead` blocks have no tile to write: all their warps scout together (the
// same code, more warps wide).  Every wait in the kernel is for work of a block with a SMALLER ticket, which is running
// or done, and nothing ever waits for tile warps: whatever number of blocks is really resident, nothing can deadlock;
// no second stream, no co-residency assumption.
//
// How it got here (profiles/README.md, round 2 "single kernel"): (1) walks by the tile's own warps: a walk is about 1 us
// of latency per descent on one lane, during which the tile's 40 KB of bit strings write nothing -- 0.79 ms per Mi boards
// against 0.54 without the walks; (2) a second kernel of small scout blocks beside the tile kernel: four tile blocks fix
// the SM's shared-memory carve-out and no scout block becomes resident; (3) a scout warp that also did the look-back
// right after its walks: a wave of blocks reaches that point together, so every scout waited for the longest walk of
// the wave (0.72 ms); hence the lag.
//
// Monsters.  A tree that exceeds the round-0 budget (walk_budget(0): no ordinary position does) is still walked to the
// end by its scout lane -- results never depend on the budget -- but the kernel raises a word in mapped host memory; the
// host side (drg_abi.cu, use_tile_pipeline) then routes the following calls of that rule family through the split-walk
// pipeline, which deals such trees over thousands of lanes (k_walk_root / k_walk_rounds), until they stop appearing.
#pragma once
#include "drg_kernels.cuh"

namespace drg {

// Tile size.  The bit strings of a tile (S*S bits per board) are what limits the blocks an SM holds, and every block
// is one latency chain (wait for the ready word, counts, scan, emit, then the rows).  Measured (1 Mi Standard boards,
// mask + records, ahead = 2 waves): 128-board tiles, 3 scout warps, tables in shared memory 0.646 ms; the same with the
// tables read from global memory through L1 0.779; 64-board tiles (8 chains per SM instead of 4, tables from L1) 0.845;
// 32-board tiles 0.95.  -DDRG_TILE_WARPS / -DDRG_SCOUT_WARPS / -DDRG_FUSED_SMEM_TABLES=0 rebuild those variants.
#ifndef DRG_TILE_WARPS
#define DRG_TILE_WARPS 4
#endif
#ifndef DRG_SCOUT_WARPS
#define DRG_SCOUT_WARPS 3
#endif
#ifndef DRG_FUSED_SMEM_TABLES
#define DRG_FUSED_SMEM_TABLES 1
#endif
constexpr int kTileWarps = DRG_TILE_WARPS;            // tile warps of a block = 32-board quarters of a tile
constexpr int kTileBoards = kTileWarps * 32;
constexpr int kScoutWarps = DRG_SCOUT_WARPS;          // scout warps of a block (warps kTileWarps ..)
constexpr int kFusedWarps = kTileWarps + kScoutWarps;
constexpr int kFusedThreads = kFusedWarps * 32;
constexpr int kScoutWalkers = 8 * kTileWarps;         // columns of the scout's DFS stack = tree-search boards with a pool slot
constexpr int kScoutKeep = 4;                         // best leaves kept per tree-search board
constexpr unsigned long long kTileAgg = 1ull << 62;   // look-back word of a tile: its own total is known (and its counts,
                                                      // classes and kept leaves are in memory)
constexpr unsigned long long kTilePre = 2ull << 62;   // ... and the total of everything before it: value = inclusive prefix
constexpr unsigned long long kTileVal = kTileAgg - 1;
constexpr unsigned long long kTileReady = 1ull << 63; // ready word: value = records before the tile

// board classes (scout -> tile warps): CLS_DEEP + k = tree-search board with pool slot k of its tile
enum : uint8_t { CLS_QUIET = 0, CLS_SINGLE = 1, CLS_DEEP = 2, CLS_DEEP_NOSLOT = 0xFF };

struct FusedArgs {
  int64_t n;
  const uint64_t *wm, *wk, *bm, *bk;
  const uint8_t* turn;
  uint32_t* counts;
  uint64_t* offsets;               // [n + 1]
  DrgMove* moves;
  uint64_t moves_cap;
  uint8_t* mask;                   // [n, S*S] (WITH_MASK)
  float* tensor;                   // [n, 4, S] (WITH_TENSOR)
  unsigned long long* overflow;    // boards whose records were clipped / whose capture path exceeds a record
  uint8_t* cls;                    // [n] scratch: class of every board
  DrgMove* pool;                   // [tiles * kScoutWalkers * kScoutKeep] scratch: kept leaves
  uint32_t* meta;                  // [tiles * kScoutWalkers] scratch: kept leaves (bits 0-7, 0xFF: walk again) | board word << 8
  unsigned long long* tile_look;   // [tiles] scratch, zero before the launch: look-back words
  unsigned long long* tile_ready;  // [tiles] scratch, zero before the launch: what the tile warps wait for
  unsigned* ticket;                // zero before the launch
  unsigned* monster;               // mapped host word: a tree exceeded the round-0 budget
  unsigned tiles, ahead, lag;
  unsigned long long* timing;      // measurement aid (DRG_FUSED_TIMING): clock sums per phase, or NULL
};

__device__ __forceinline__ unsigned long long word_load(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long word_load_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void word_store_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// barrier over `threads` threads of the block (whole warps), hardware barrier `id` (0 is __syncthreads')
__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct ScoutCtl {                // the scouted tile
  int n_cap, n_deep;
  uint32_t wsum[8];              // per-warp totals
  uint8_t cap[kTileBoards];           // boards with captures (phase A -> B)
  uint8_t deep[kTileBoards];          // ... that need the tree search
  uint8_t cls[kTileBoards];
  uint32_t cnt[kTileBoards];          // len(legal_moves)
};

struct TileCtl {                 // the block's own tile
  int n_single, n_deep;
  unsigned tile;
  unsigned long long base;       // records before this tile
  uint32_t warp_sum[kTileWarps];
  uint8_t single[kTileBoards];        // boards whose captures are single man jumps
  uint8_t deep[kTileBoards];          // boards that needed the tree search
};

constexpr int kTileRewalkers = 2 * kTileWarps;  // lanes that can walk a tree again (boards without a complete pool slot)

template <int S>
__host__ __device__ constexpr int fused_bits_words() { return kTileBoards * S * S / 32; }

// geometry tables in global memory (copied to shared memory by every block unless DRG_FUSED_SMEM_TABLES=0)
template <int S> __device__ __forceinline__ const Tables<S>& fused_tables();
template <> __device__ __forceinline__ const Tables<50>& fused_tables<50>() { return g_tab50; }
template <> __device__ __forceinline__ const Tables<32>& fused_tables<32>() { return g_tab32; }

template <int S, bool WITH_MASK, bool WITH_TENSOR>
constexpr size_t fused_smem_bytes() {
  size_t b = (DRG_FUSED_SMEM_TABLES ? sizeof(Tables<S>) : 0) + ((sizeof(TileCtl) + 15) & ~(size_t)15) + ((sizeof(ScoutCtl) + 15) & ~(size_t)15) +
             3 * sizeof(uint32_t) * kTileBoards + sizeof(Frame) * kMaxLevels * (kTileRewalkers + kScoutWalkers);
  b = (b + 15) & ~(size_t)15;
  const size_t mask_b = WITH_MASK ? (size_t)fused_bits_words<S>() * 4 : 0;
  const size_t tens_b = WITH_TENSOR ? (size_t)kTileBoards * 8 * 4 : 0;
  return b + (mask_b > tens_b ? mask_b : tens_b) + 16;
}

// Step 1: everything the tile warps need to know about tile `tile` except where its records start, by `nw` warps
// (kScoutWarps: the scout warps; kFusedWarps: the whole block, first wave).  `slot` = index of the calling thread among
// the nw * 32 participants.
template <class R>
__device__ __forceinline__ void scout_tile(const FusedArgs& a, int64_t tile, int nw, int slot, const Tables<R::S>& T,
                                           ScoutCtl& sc, Frame* s_stack, uint32_t& ovf) {
  const int lane = threadIdx.x & 31;
  const int nthr = nw * 32;
  auto barrier = [&]() {
    if (nw == kScoutWarps) named_barrier(3, kScoutWarps * 32);
    else named_barrier(1, kFusedThreads);
  };
  NullSink ns;
  const int64_t b0 = tile * kTileBoards;
  const int nb = (int)((a.n - b0) < kTileBoards ? (a.n - b0) : kTileBoards);
  const bool tm = a.timing && nw == kScoutWarps && slot == 0;
  long long tc = tm ? clock64() : 0;
  auto lap = [&](int k) {
    if (tm) {
      const long long now = clock64();
      atomicAdd(a.timing + k, (unsigned long long)(now - tc));
      tc = now;
    }
  };
  if (slot == 0) sc.n_cap = sc.n_deep = 0;
  barrier();
  // ---- phase A: capture pre-check, quiet moves counted with popc ----
  for (int j = slot; j < kTileBoards; j += nthr) {
    uint32_t q = 0;
    if (j < nb) {
      const Pos<R> p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[b0 + j] & 1, b0 + j);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      if (R::kOptionalCapture || !c.any()) q = (uint32_t)gen_quiet<R, true>(p, ns, T);
      if (c.any()) sc.cap[atomicAdd(&sc.n_cap, 1)] = (uint8_t)j;
    }
    sc.cnt[j] = q;
    sc.cls[j] = CLS_QUIET;
  }
  barrier();
  lap(4);
  // ---- phase B: capture boards on consecutive threads; single man jumps are counted bit-parallel ----
  for (int k = slot; k < sc.n_cap; k += nthr) {
    const int j = sc.cap[k];
    const Pos<R> p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[b0 + j] & 1, b0 + j);
    MenJumps<R> mj;
    const auto c = find_capturers<R>(p, T, mj);
    if (!c.kings && only_single_jumps<R>(p, mj)) {
      sc.cnt[j] += (uint32_t)single_jumps<R, false>(p, mj, ns, T);
      sc.cls[j] = CLS_SINGLE;
    } else {
      const int d = atomicAdd(&sc.n_deep, 1);
      sc.deep[d] = (uint8_t)j;
      sc.cls[j] = d < kScoutWalkers ? (uint8_t)(CLS_DEEP + d) : (uint8_t)CLS_DEEP_NOSLOT;
    }
  }
  barrier();
  lap(5);
  // ---- phase C: the tile's tree-search boards, walker w on lane w / nw of participant warp w % nw.  Statistics pass
  // that also keeps the leaves of the best metric so far (PASS 2) in the board's pool slot ----
  const int n_deep = sc.n_deep;
  const int lanes_w = kScoutWalkers / nw;        // walking lanes per warp
  const int nwalk = lanes_w * nw;
  const int walker = lane * nw + slot / 32;      // < nwalk for the walking lanes
  Frame* stack = s_stack + walker;
  if (lane < lanes_w) {
    for (int k = walker; k < n_deep; k += nwalk) {
      const int j = sc.deep[k];
      const Pos<R> p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[b0 + j] & 1, b0 + j);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      KeepSink keep;
      keep.out = a.pool + ((size_t)tile * kScoutWalkers + (k < kScoutWalkers ? k : 0)) * kScoutKeep;
      keep.n = 0;
      keep.room = k < kScoutWalkers ? (uint32_t)kScoutKeep : 0u;
      WalkLane<R, 2> wl;
      wl.start_root(p, c, 0u, walk_budget(0), T);
      for (;;) {
        const int s = wl.advance(keep, stack, kScoutWalkers, T, ovf);
        if (s == kWalkDone) break;
        if (s == kWalkWantsSplit) {  // a monster: finished here all the same; the host hears about it
          wl.budget = kNoBudget;
          *reinterpret_cast<volatile unsigned*>(a.monster) = 1u;
        }
      }
      const uint32_t bw = walk_board_word<R>(wl.st);
      if (k < kScoutWalkers)
        a.meta[(size_t)tile * kScoutWalkers + k] = (keep.n <= keep.room ? keep.n : 0xFFu) | (bw << 8);
      sc.cnt[j] += walk_kept<R>(wl.result(), bw);
    }
  }
  barrier();
  lap(6);
  // ---- publish: counts and classes (coalesced), then the tile's total ----
  uint32_t mine = 0;
  for (int j = slot; j < kTileBoards; j += nthr) {
    const uint32_t c = sc.cnt[j];
    mine += c;
    if (j < nb) {
      a.counts[b0 + j] = c;
      a.cls[b0 + j] = sc.cls[j];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
  if (lane == 0) sc.wsum[slot / 32] = mine;
  __threadfence();
  barrier();
  if (slot == 0) {
    unsigned long long total = 0;
    for (int w = 0; w < nw; w++) total += sc.wsum[w];
    // tile 0 has nothing before it: its total is its inclusive prefix
    word_store_release(a.tile_look + tile, (tile == 0 ? kTilePre : kTileAgg) | total);
    if (tile == 0) word_store_release(a.tile_ready + 0, kTileReady | 0ull);
  }
  lap(7);
}

// Step 2: the records before tile `tile` (> 0), by one warp: look-back over the words of the tiles before it.
__device__ __forceinline__ void prefix_tile(const FusedArgs& a, int64_t tile) {
  const int lane = threadIdx.x & 31;
  // every word read here is written by a block with a smaller ticket, which is running or done: it arrives.  A wave of
  // blocks reaches this point almost together, so the nearest finished prefix is often hundreds of tiles back: four
  // windows of 32 words are loaded per round trip.  The spin limits only keep a logic error from hanging the device
  // (it then shows as wrong offsets in the tests).
  unsigned long long own = word_load(a.tile_look + tile);
  for (unsigned spins = 0; (own >> 62) == 0 && spins < (1u << 22); spins++) {
    __nanosleep(40);
    own = word_load(a.tile_look + tile);
  }
  unsigned long long excl = 0;
  bool done = false;
  for (int64_t look = tile - 1; !done; look -= 128) {
    unsigned long long w[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int64_t idx = look - 32 * r - lane;
      w[r] = idx >= 0 ? word_load(a.tile_look + idx) : kTilePre;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (done) break;
      const int64_t idx = look - 32 * r - lane;
      for (unsigned spins = 0; __any_sync(0xFFFFFFFFu, (w[r] >> 62) == 0) && spins < (1u << 22); spins++) {
        if ((w[r] >> 62) == 0) {
          __nanosleep(40);
          w[r] = word_load(a.tile_look + idx);
        }
      }
      const unsigned pre = __ballot_sync(0xFFFFFFFFu, (w[r] >> 62) == 2);
      const int first = pre ? __ffs((int)pre) - 1 : 31;  // nearest tile whose inclusive prefix is known
      unsigned long long part = lane <= first ? (w[r] & kTileVal) : 0ull;
#pragma unroll
      for (int o = 16; o; o >>= 1) part += __shfl_down_sync(0xFFFFFFFFu, part, o);
      excl += __shfl_sync(0xFFFFFFFFu, part, 0);
      done = pre != 0;
    }
  }
  // the words were read relaxed: the fence orders what their writers published (counts, classes, kept leaves of every
  // tile up to this one) before the ready word
  __threadfence();
  if (lane == 0) {
    word_store_release(a.tile_look + tile, kTilePre | (excl + (own & kTileVal)));
    word_store_release(a.tile_ready + tile, kTileReady | excl);
  }
}

template <class R, bool WITH_MASK, bool WITH_TENSOR>
__global__ void __launch_bounds__(kFusedThreads) k_movegen_fused(FusedArgs a) {
  constexpr int S = R::S;
  constexpr int kRow = S * S;
  extern __shared__ __align__(16) uint8_t smem[];
#if DRG_FUSED_SMEM_TABLES
  Tables<S>& T = *reinterpret_cast<Tables<S>*>(smem);
  TileCtl& ctl = *reinterpret_cast<TileCtl*>(smem + sizeof(Tables<S>));
#else
  const Tables<S>& T = fused_tables<S>();
  TileCtl& ctl = *reinterpret_cast<TileCtl*>(smem);
#endif
  ScoutCtl& sc = *reinterpret_cast<ScoutCtl*>(reinterpret_cast<uint8_t*>(&ctl) + ((sizeof(TileCtl) + 15) & ~15));
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(&sc) + ((sizeof(ScoutCtl) + 15) & ~15));
  uint32_t* s_off = s_cnt + kTileBoards;   // first record of the board within the tile
  uint32_t* s_n = s_off + kTileBoards;     // capture boards: records written before the captures (quiet moves, american.py:96)
  Frame* s_restack = reinterpret_cast<Frame*>(s_n + kTileBoards);
  Frame* s_scstack = s_restack + kMaxLevels * kTileRewalkers;
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(
      (reinterpret_cast<uintptr_t>(s_scstack + kMaxLevels * kScoutWalkers) + 15) & ~(uintptr_t)15);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    ctl.tile = atomicAdd(a.ticket, 1u);
    ctl.n_single = ctl.n_deep = 0;
  }
#if DRG_FUSED_SMEM_TABLES
  {
    const uint4* src = reinterpret_cast<const uint4*>(&fused_tables<S>());
    uint4* d = reinterpret_cast<uint4*>(&T);
    for (int q = tid; q < (int)(sizeof(Tables<S>) / 16); q += kFusedThreads) d[q] = src[q];
  }
#endif
  if (WITH_MASK) {
    uint4* z = reinterpret_cast<uint4*>(s_bits);
    for (int q = tid; q < fused_bits_words<S>() / 4; q += kFusedThreads) z[q] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const int64_t ticket = ctl.tile;
  uint32_t ovf = 0;

  // ---- scouting: block b prepares tile b (step 1) and tile b - lag (step 2) -- the first `ahead` blocks with all
  // their warps (they have no tile to write), the others with their scout warps while the tile warps write tile b - ahead
  const long long tq0 = a.timing ? clock64() : 0;
  const bool first_wave = ticket < (int64_t)a.ahead;
  if (first_wave || warp >= kTileWarps) {
    if (ticket < (int64_t)a.tiles) {
      if (first_wave) scout_tile<R>(a, ticket, kFusedWarps, tid, T, sc, s_scstack, ovf);
      else scout_tile<R>(a, ticket, kScoutWarps, tid - kTileBoards, T, sc, s_scstack, ovf);
    }
    const int64_t pt = ticket - (int64_t)a.lag;
    if (warp == kTileWarps && pt > 0 && pt < (int64_t)a.tiles) {
      const long long tp = a.timing ? clock64() : 0;
      prefix_tile(a, pt);
      if (a.timing && lane == 0) atomicAdd(a.timing + 8, (unsigned long long)(clock64() - tp));
    }
    if (ovf && a.overflow) atomicAdd(a.overflow, 1ull);
    if (a.timing && tid == kTileBoards && !first_wave) atomicAdd(a.timing + 0, (unsigned long long)(clock64() - tq0));
    return;
  }

  // ---- tile warps (128 threads, barrier 2) ----
  const int64_t tile = ticket - a.ahead;
  const int64_t base = tile * kTileBoards;
  const int64_t i = base + tid;
  Pos<R> p;
  p.wm = p.wk = p.bm = p.bk = 0;
  p.turn = 0;
  if (i < a.n) p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[i] & 1, i);  // does not depend on the scouts
  if (tid == 0) {
    // published by step 2 of block tile + lag, which holds a smaller ticket than this block (lag < ahead); the spin
    // limit keeps a logic error from hanging the device
    unsigned long long w = word_load_acquire(a.tile_ready + tile);
    for (unsigned spins = 0; !(w & kTileReady) && spins < (1u << 24); spins++) {
      __nanosleep(100);
      w = word_load_acquire(a.tile_ready + tile);
    }
    ctl.base = w & kTileVal;
  }
  named_barrier(2, kTileBoards);
  const long long tq1 = a.timing ? clock64() : 0;
  const uint64_t gbase = ctl.base;

  // ---- offsets of the tile's boards ----
  uint32_t cnt = 0;
  int cls = CLS_QUIET;
  if (i < a.n) {
    cnt = __ldcg(a.counts + i);
    cls = __ldcg(a.cls + i);
  }
  uint32_t inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) ctl.warp_sum[warp] = inc;
  s_cnt[tid] = cnt;
  if (cls == CLS_SINGLE) ctl.single[atomicAdd(&ctl.n_single, 1)] = (uint8_t)tid;
  if (cls >= CLS_DEEP) ctl.deep[atomicAdd(&ctl.n_deep, 1)] = (uint8_t)tid;
  named_barrier(2, kTileBoards);
  uint32_t before = 0;
#pragma unroll
  for (int w = 0; w < kTileWarps; w++) before += w < warp ? ctl.warp_sum[w] : 0u;
  const uint32_t local = before + inc - cnt;
  s_off[tid] = local;
  const uint64_t off = gbase + local;
  if (i < a.n) {
    a.offsets[i] = off;
    if (i == a.n - 1) a.offsets[a.n] = off + cnt;
  }

  // ---- emit: every move once, at its place ----
  uint32_t* mbits = WITH_MASK ? s_bits : nullptr;
  if (i < a.n) {
    BitSink<S> sink;
    sink.out = a.moves + off;
    sink.bits = mbits;
    sink.bit0 = (uint32_t)tid * kRow;
    sink.n = 0;
    sink.room = (uint32_t)emit_room(off, off + cnt, a.moves_cap);
    if (R::kOptionalCapture || cls == CLS_QUIET) gen_quiet<R, false>(p, sink, T);
    // capture boards: the quiet part; tree-search boards: quiet part | class (= pool slot)
    s_n[tid] = cls >= CLS_DEEP ? (sink.n << 8) | (uint32_t)cls : sink.n;
    if (cnt > sink.room) ++ovf;  // one per clipped board
  }
  named_barrier(2, kTileBoards);
  for (int k = tid; k < ctl.n_single; k += kTileBoards) {  // single-jump boards on consecutive threads
    const int t = ctl.single[k];
    const Pos<R> q = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[base + t] & 1, base + t);
    MenJumps<R> mj;
    find_capturers<R>(q, T, mj);
    const uint64_t o2 = gbase + s_off[t];
    BitSink<S> sink;
    sink.out = a.moves + o2;
    sink.bits = mbits;
    sink.bit0 = (uint32_t)t * kRow;
    sink.n = s_n[t];
    sink.room = (uint32_t)emit_room(o2, o2 + s_cnt[t], a.moves_cap);
    single_jumps<R, true>(q, mj, sink, T);
  }
  {
    // tree-search boards: the leaves their scout kept are copied to their place (and give the mask bits); boards with
    // more leaves than a pool slot holds are walked again here, a few lanes at a time
    bool rewalk = false;
    int t = 0;
    uint32_t qn = 0;
    const int k = kTileBoards - 1 - tid;  // the LAST threads take them (the first ones are busy with the single-jump boards)
    if (k < ctl.n_deep) {
      t = ctl.deep[k];
      const uint32_t sn = s_n[t];
      const int slot = (int)(sn & 0xFFu);
      qn = sn >> 8;
      const uint32_t meta =
          slot == CLS_DEEP_NOSLOT ? 0xFFu : __ldcg(a.meta + (size_t)tile * kScoutWalkers + (slot - CLS_DEEP));
      if ((meta & 0xFFu) == 0xFFu) {
        rewalk = true;
      } else {
        const uint64_t o2 = gbase + s_off[t];
        const uint32_t room = (uint32_t)emit_room(o2, o2 + s_cnt[t], a.moves_cap);
        const uint32_t bw = meta >> 8;
        const bool kings_only = R::kFilter == FILTER_VALUE && (bw & 1u);  // frisian.py:264-268
        const uint4* src =
            reinterpret_cast<const uint4*>(a.pool + ((size_t)tile * kScoutWalkers + (slot - CLS_DEEP)) * kScoutKeep);
        uint32_t n = qn;
        for (uint32_t r = 0; r < (meta & 0xFFu); r++) {
          const uint4 lo = __ldcg(src + 2 * r), hi = __ldcg(src + 2 * r + 1);
          Rec rec;
          rec.w[0] = lo.x; rec.w[1] = lo.y; rec.w[2] = lo.z; rec.w[3] = lo.w;
          rec.w[4] = hi.x; rec.w[5] = hi.y; rec.w[6] = hi.z; rec.w[7] = hi.w;
          if (kings_only && !((rec.w[2] >> 8) & DRG_MF_KINGMOVE)) continue;
          if (n < room) rec_store32(a.moves + o2 + n, rec);
          n++;
          if (WITH_MASK) {
            const uint32_t b = (uint32_t)t * kRow + (uint32_t)(rec_from(rec) * S + rec_to(rec));
            atomicOr(&s_bits[b >> 5], 1u << (b & 31));
          }
        }
      }
    }
    unsigned todo = __ballot_sync(0xFFFFFFFFu, rewalk);
    while (todo) {
      const int slot = __popc(todo & ((1u << lane) - 1));
      if (rewalk && slot < kTileRewalkers / kTileWarps) {
        const Pos<R> q = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[base + t] & 1, base + t);
        MenJumps<R> mj;
        const auto c = find_capturers<R>(q, T, mj);
        const uint64_t o2 = gbase + s_off[t];
        BitSink<S> sink;
        sink.out = a.moves + o2;
        sink.bits = mbits;
        sink.bit0 = (uint32_t)t * kRow;
        sink.n = qn;
        sink.room = (uint32_t)emit_room(o2, o2 + s_cnt[t], a.moves_cap);
        emit_captures<R>(q, c, sink, s_restack + warp * (kTileRewalkers / kTileWarps) + slot, kTileRewalkers, T, ovf);
        rewalk = false;
      }
      todo = __ballot_sync(0xFFFFFFFFu, rewalk);
    }
  }
  named_barrier(2, kTileBoards);
  const long long tq2 = a.timing ? clock64() : 0;
  if (ovf && a.overflow) atomicAdd(a.overflow, (unsigned long long)ovf);

  // ---- rows (base.py:807-838): this warp's 32 rows are one contiguous, 16-byte aligned range ----
  const int64_t wb0 = base + warp * 32;
  const int nboards = wb0 >= a.n ? 0 : ((a.n - wb0) < 32 ? (int)(a.n - wb0) : 32);
  if (WITH_MASK && nboards) {
    const uint16_t* wbits = reinterpret_cast<const uint16_t*>(s_bits) + warp * (32 * kRow / 16);
    uint8_t* dst = a.mask + (size_t)wb0 * kRow;
    const int bytes = nboards * kRow;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int q = lane; q < bytes / 16; q += 32) {
      const uint32_t h = wbits[q];
      d4[q] = make_uint4(bytes_of_nibble(h & 15u), bytes_of_nibble((h >> 4) & 15u), bytes_of_nibble((h >> 8) & 15u),
                         bytes_of_nibble((h >> 12) & 15u));
    }
    const uint8_t* wnib = reinterpret_cast<const uint8_t*>(wbits);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(dst);
    for (int q = (bytes / 16) * 4 + lane; q < bytes / 4; q += 32) d1[q] = bytes_of_nibble((wnib[q >> 1] >> ((q & 1) * 4)) & 15u);
  }
  if (WITH_TENSOR) {
    // to_tensor(perspective = side to move) (base.py:753-805): 4 planes of S floats per board, packed into a 4*S-bit
    // string (8 words) per board, then the warp streams float4s
    named_barrier(2, kTileBoards);  // the mask bit strings are dead from here on: their memory holds the tensor bit strings
    if (nboards) {
      uint32_t* bits = s_bits + warp * 32 * 8;
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
      float4* dst = reinterpret_cast<float4*>(a.tensor + (size_t)wb0 * 4 * S);
      const int total4 = nboards * S;
      for (int q = lane; q < total4; q += 32) {
        const int board = q / S, bit = (q - board * S) * 4;
        const uint32_t nib = (bits[board * 8 + (bit >> 5)] >> (bit & 31)) & 0xFu;
        dst[q] = make_float4((nib & 1) ? 1.0f : 0.0f, (nib & 2) ? 1.0f : 0.0f, (nib & 4) ? 1.0f : 0.0f,
                             (nib & 8) ? 1.0f : 0.0f);
      }
    }
  }
  if (a.timing && tid == 0) {
    const long long tq3 = clock64();
    atomicAdd(a.timing + 1, (unsigned long long)(tq1 - tq0));  // wait for the ready word
    atomicAdd(a.timing + 2, (unsigned long long)(tq2 - tq1));  // offsets + emit
    atomicAdd(a.timing + 3, (unsigned long long)(tq3 - tq2));  // rows (+ tensor) of warp 0
  }
}

}  // namespace drg
