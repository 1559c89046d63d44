// drg_fused.cuh -- the fused call (drg_movegen: counts, offsets, records, mask rows, tensor planes) as ONE kernel.
//
// Replaces, for ordinary batches, the split pipeline of drg_kernels.cuh (k_emit<rows> beside k_count -> k_walk_root ->
// scan x3 -> k_emit<records> -> k_emit_deep -> k_mask_fixup: the moves were enumerated three times and the chain's wide
// kernels competed with the 2.6 GB row writer for the SMs, one kernel-level dependency after the other).
//
// Blocks are numbered by an atomic ticket b and have SIX warps; block b prepares tile b and writes tile b - ahead:
//
//   scout warps  (two) for tile b it does everything that is latency, not bytes -- exact
//                capture pre-check + popc counts (phase A), single-jump boards re-dealt and counted bit-parallel (B), the
//                tile's tree-search boards (7 % of random-play Standard positions) walked on its lanes (C), a scan of
//                the 128 counts, a decoupled look-back over the tiles before (one 64-bit word per tile: flag | value)
//                and -- the offsets are known now -- the tree-search boards' records written at their final place by a
//                second walk.  Then it publishes the tile's word: "ready | records before this tile".
//   tile warps   (four) wait for the word of tile b - ahead -- published `ahead` blocks earlier, so it is there -- read
//                counts and board classes, scan 128 counts, then enumerate every move ONCE, at its final place: records
//                as 32-byte full-sector stores, mask bits into the tile's bit strings in shared memory; tree-search
//                boards get their bits from the records the scout wrote.  Each warp then widens its bit string to
//                bytes and streams its 32 rows with coalesced 16-byte stores; tensor planes as float4s.
//
// `ahead` is the number of blocks the GPU holds at once (occupancy x SMs), so the scout of a tile belongs to a block
// of the wave before; the grid has tiles + ahead blocks.  The first `ahead` blocks have no tile to write: all their five
// warps scout together (the same code, five warps wide) and the block ends.  Every wait in the kernel is for work of a
// block with a smaller ticket, which is running or done, and scouts never wait for tile warps: whatever number of
// blocks is really resident, nothing can deadlock; no second stream, no co-residency assumption.
//
// Why the walks are not simply done by the tile's own warps (first version, profiles/README.md round 2): a walk is about
// 1 us of latency per descent on one lane, during which the tile's 40 KB of bit strings write nothing -- 0.79 ms per Mi
// boards against 0.54 without the walks.  A second kernel of small scout blocks beside the tile kernel (second version)
// deadlocks until its spin limit: four tile blocks fix the SM's shared-memory carve-out and no scout block becomes
// resident.
//
// Monsters.  A tree that exceeds the round-0 budget (walk_budget(0): no ordinary position does) is still walked to the
// end by its scout lane -- results never depend on the budget -- but the kernel raises a word in mapped host memory; the
// host side (drg_abi.cu, use_tile_pipeline) then routes the following calls of that rule family through the split-walk
// pipeline, which deals such trees over thousands of lanes (k_walk_root / k_walk_rounds), until they stop appearing.
#pragma once
#include "drg_kernels.cuh"

namespace drg {

constexpr int kScoutWarps = 2;                        // scout warps of a block (warps kWarps ..)
constexpr int kFusedWarps = kWarps + kScoutWarps;     // four tile warps + the scout warps
constexpr int kFusedThreads = kFusedWarps * 32;
constexpr int kScoutWalkers = 32;                     // columns of the scout's DFS stack
constexpr int kScoutKeep = 4;                         // best leaves a scout lane keeps of a tree (statistics pass), so that
                                                      // the records are a copy, not a second walk
constexpr unsigned long long kTileAgg = 1ull << 62;   // look-back word of a tile: its own total is known
constexpr unsigned long long kTilePre = 2ull << 62;   // ... and the total of everything before it: value = inclusive prefix
constexpr unsigned long long kTileVal = kTileAgg - 1;
constexpr unsigned long long kTileReady = 1ull << 63; // ready word: counts, classes and tree-search records are in memory,
                                                      //             value = records before the tile

enum : uint8_t { CLS_QUIET = 0, CLS_SINGLE = 1, CLS_DEEP = 2 };

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
  uint8_t* cls;                    // [n] scratch: CLS_* of every board (scout -> tile warps)
  unsigned long long* tile_look;   // [tiles] scratch, zero before the launch: look-back words of the scouts
  unsigned long long* tile_ready;  // [tiles] scratch, zero before the launch: what the tile warps wait for
  unsigned* ticket;                // zero before the launch
  unsigned* monster;               // mapped host word: a tree exceeded the round-0 budget
  unsigned tiles, ahead;
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
__device__ __forceinline__ void word_store(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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
  unsigned long long base;
  uint8_t cap[kBlock];           // boards with captures (phase A -> B)
  uint8_t deep[kBlock];          // ... that need the tree search
  uint8_t cls[kBlock];
  uint16_t bw[kBlock];           // tree-search board k: best metric << 1 | king preference (< 2^16)
  uint8_t kept[kBlock];          // ... leaves in its keep slot, 0xFF: not (all) kept, walk again
  uint32_t cnt[kBlock];          // len(legal_moves)
  uint32_t off[kBlock];          // first record of the board within the tile
};

struct TileCtl {                 // the block's own tile
  int n_single, n_deep;
  unsigned tile;
  unsigned long long base;       // records before this tile
  uint32_t warp_sum[kWarps];
  uint8_t single[kBlock];        // boards whose captures are single man jumps
  uint8_t deep[kBlock];          // boards whose captures the scout wrote
};

constexpr int kTileRewalkers = 8;  // lanes that can re-walk a tree (boards whose records were clipped: mask bits only)

template <int S, bool WITH_MASK, bool WITH_TENSOR>
constexpr size_t fused_smem_bytes() {
  size_t b = sizeof(Tables<S>) + ((sizeof(TileCtl) + 15) & ~(size_t)15) + ((sizeof(ScoutCtl) + 15) & ~(size_t)15) +
             3 * sizeof(uint32_t) * kBlock + sizeof(Frame) * kMaxLevels * (kTileRewalkers + kScoutWalkers) +
             sizeof(DrgMove) * kScoutKeep * kScoutWalkers;
  b = (b + 15) & ~(size_t)15;
  const size_t mask_b = WITH_MASK ? (size_t)emit_bits_words<S>() * 4 : 0;
  const size_t tens_b = WITH_TENSOR ? (size_t)kBlock * 8 * 4 : 0;
  return b + (mask_b > tens_b ? mask_b : tens_b) + 16;
}

// Everything the tile warps need to know about tile `tile`, by `nw` warps (kScoutWarps: the scout warps; kFusedWarps: the
// whole block, first wave).  `slot` = index of the calling thread among the nw * 32 participants.
template <class R>
__device__ __forceinline__ void scout_tile(const FusedArgs& a, int64_t tile, int nw, int slot, const Tables<R::S>& T,
                                           ScoutCtl& sc, Frame* s_stack, DrgMove* s_keep, uint32_t& ovf) {
  constexpr int S = R::S;
  const int lane = threadIdx.x & 31;
  const int nthr = nw * 32;
  const bool scan_warp = (int)(threadIdx.x >> 5) == kWarps;  // the first scout warp
  auto barrier = [&]() {
    if (nw == kScoutWarps) named_barrier(3, kScoutWarps * 32);
    else named_barrier(1, kFusedThreads);
  };
  NullSink ns;
  const int64_t b0 = tile * kBlock;
  const int nb = (int)((a.n - b0) < kBlock ? (a.n - b0) : kBlock);
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
  for (int j = slot; j < kBlock; j += nthr) {
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
      sc.deep[atomicAdd(&sc.n_deep, 1)] = (uint8_t)j;
      sc.cls[j] = CLS_DEEP;
    }
  }
  barrier();
  lap(5);
  // ---- phase C: the tile's tree-search boards (statistics pass), walker w on lane w / nw of participant warp w % nw ----
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
      // statistics pass that also keeps the leaves of the best metric so far (PASS 2) in the walker's keep slot
      KeepSink keep;
      keep.out = s_keep + k * kScoutKeep;
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
      sc.bw[k] = (uint16_t)bw;
      sc.kept[k] = keep.n <= keep.room ? (uint8_t)keep.n : (uint8_t)0xFF;
      sc.cnt[j] += walk_kept<R>(wl.result(), bw);
    }
  }
  barrier();
  lap(6);
  // ---- scan of the 128 counts and look-back over the tiles before: the scout warp, four boards per lane ----
  if (scan_warp) {
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      v[r] = sc.cnt[lane * 4 + r];
      sum += v[r];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += u;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    uint32_t run = inc - sum;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      sc.off[lane * 4 + r] = run;
      run += v[r];
    }
    unsigned long long excl = 0;
    if (tile > 0) {
      if (lane == 0) word_store(a.tile_look + tile, kTileAgg | (unsigned long long)total);
      // every tile before this one is being scouted by a block with a smaller ticket, which is running or done: its
      // word arrives.  A wave of blocks reaches this point almost together, so the nearest finished prefix is often
      // hundreds of tiles back: four windows of 32 words are loaded per round trip.  The spin limit only keeps a logic
      // error from hanging the device (it then shows as wrong offsets in the tests)
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
    }
    if (lane == 0) {
      word_store(a.tile_look + tile, kTilePre | (excl + (unsigned long long)total));
      sc.base = excl;
    }
  }
  barrier();
  lap(7);
  // counts and classes of the tile (coalesced); the tile warps read them after they have seen the tile's word
  for (int j = slot; j < nb; j += nthr) {
    a.counts[b0 + j] = sc.cnt[j];
    a.cls[b0 + j] = sc.cls[j];
  }
  const unsigned long long gbase = sc.base;
  // ---- the tree-search boards' records, at their final place (second walk; the best metric is known) ----
  if (lane < lanes_w) {
    NoAlloc noalloc;
    for (int k = walker; k < n_deep; k += nwalk) {
      const int j = sc.deep[k];
      const Pos<R> p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[b0 + j] & 1, b0 + j);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      const unsigned long long off = gbase + sc.off[j];
      DeepSink<S> sink;
      sink.out = a.moves + off;
      sink.mrow = nullptr;  // mask bits are the tile warps': they read these records back
      sink.row0 = 0;
      sink.fix = MaskFix{nullptr, nullptr, 0};
      sink.room = (uint32_t)emit_room(off, off + sc.cnt[j], a.moves_cap);
      sink.n = R::kOptionalCapture ? (uint32_t)gen_quiet<R, true>(p, ns, T) : 0u;  // quiet moves come first (american.py:96)
      const uint32_t bw = sc.bw[k];
      const int n_kept = sc.kept[k];
      if (n_kept != 0xFF) {  // the statistics pass kept them all
        const bool kings_only = R::kFilter == FILTER_VALUE && (bw & 1u);  // frisian.py:264-268
        const uint4* src = reinterpret_cast<const uint4*>(s_keep + k * kScoutKeep);
        for (int r = 0; r < n_kept; r++) {
          const uint4 lo = src[2 * r], hi = src[2 * r + 1];
          Rec rec;
          rec.w[0] = lo.x; rec.w[1] = lo.y; rec.w[2] = lo.z; rec.w[3] = lo.w;
          rec.w[4] = hi.x; rec.w[5] = hi.y; rec.w[6] = hi.z; rec.w[7] = hi.w;
          if (kings_only && !((rec.w[2] >> 8) & DRG_MF_KINGMOVE)) continue;
          if (sink.n < sink.room) rec_store32(sink.out + sink.n, rec);
          sink.n++;
        }
      } else {
        WalkLane<R, 1> wl;
        wl.start_root(p, c, 0u, kNoBudget, T);
        wl.st.best = (int)(bw >> 1);
        wl.st.n_best_king = (int)(bw & 1u);
        wl.run(sink, stack, kScoutWalkers, T, ovf, noalloc);
      }
    }
  }
  __threadfence();
  barrier();
  lap(8);
  if (scan_warp && lane == 0) word_store_release(a.tile_ready + tile, kTileReady | gbase);
}

template <class R, bool WITH_MASK, bool WITH_TENSOR>
__global__ void __launch_bounds__(kFusedThreads) k_movegen_fused(FusedArgs a) {
  constexpr int S = R::S;
  constexpr int kRow = S * S;
  extern __shared__ __align__(16) uint8_t smem[];
  Tables<S>& T = *reinterpret_cast<Tables<S>*>(smem);
  TileCtl& ctl = *reinterpret_cast<TileCtl*>(smem + sizeof(Tables<S>));
  ScoutCtl& sc = *reinterpret_cast<ScoutCtl*>(reinterpret_cast<uint8_t*>(&ctl) + ((sizeof(TileCtl) + 15) & ~15));
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(&sc) + ((sizeof(ScoutCtl) + 15) & ~15));
  uint32_t* s_off = s_cnt + kBlock;   // first record of the board within the tile
  uint32_t* s_n = s_off + kBlock;     // capture boards: records written before the captures (quiet moves, american.py:96)
  Frame* s_restack = reinterpret_cast<Frame*>(s_n + kBlock);
  Frame* s_scstack = s_restack + kMaxLevels * kTileRewalkers;
  DrgMove* s_keep = reinterpret_cast<DrgMove*>(
      (reinterpret_cast<uintptr_t>(s_scstack + kMaxLevels * kScoutWalkers) + 15) & ~(uintptr_t)15);
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_keep + kScoutKeep * kScoutWalkers);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    ctl.tile = atomicAdd(a.ticket, 1u);
    ctl.n_single = ctl.n_deep = 0;
  }
  {  // tables (all warps)
    const uint4* src = reinterpret_cast<const uint4*>(S == 50 ? (const void*)&g_tab50 : (const void*)&g_tab32);
    uint4* d = reinterpret_cast<uint4*>(&T);
    for (int q = tid; q < (int)(sizeof(Tables<S>) / 16); q += kFusedThreads) d[q] = src[q];
  }
  if (WITH_MASK) {
    uint4* z = reinterpret_cast<uint4*>(s_bits);
    for (int q = tid; q < emit_bits_words<S>() / 4; q += kFusedThreads) z[q] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  int64_t tile = ctl.tile;  // ticket: the tile this block scouts
  uint32_t ovf = 0;

  // ---- scouting: block b prepares tile b -- the first `ahead` blocks with all their warps (they have no tile of their
  // own to write), the others with their scout warp, while the tile warps write tile b - ahead ----
  const long long tq0 = a.timing ? clock64() : 0;
  if (tile < (int64_t)a.ahead) {
    if (tile < (int64_t)a.tiles) scout_tile<R>(a, tile, kFusedWarps, tid, T, sc, s_scstack, s_keep, ovf);
    if (ovf && a.overflow) atomicAdd(a.overflow, 1ull);
    return;
  }
  if (warp >= kWarps) {
    if (tile < (int64_t)a.tiles) scout_tile<R>(a, tile, kScoutWarps, tid - kBlock, T, sc, s_scstack, s_keep, ovf);
    if (ovf && a.overflow) atomicAdd(a.overflow, 1ull);
    if (a.timing && tid == kBlock) atomicAdd(a.timing + 0, (unsigned long long)(clock64() - tq0));
    return;
  }
  tile -= a.ahead;

  // ---- tile warps (128 threads, barrier 2) ----
  const int64_t base = tile * kBlock;
  const int64_t i = base + tid;
  Pos<R> p;
  p.wm = p.wk = p.bm = p.bk = 0;
  p.turn = 0;
  if (i < a.n) p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[i] & 1, i);  // does not depend on the scout
  if (tid == 0) {
    // the scout's word for this tile: written by this block (first wave) or by the scout warp of block tile - ahead,
    // which holds a smaller ticket; the spin limit keeps a logic error from hanging the device
    unsigned long long w = word_load_acquire(a.tile_ready + tile);
    for (unsigned spins = 0; !(w & kTileReady) && spins < (1u << 24); spins++) {
      __nanosleep(100);
      w = word_load_acquire(a.tile_ready + tile);
    }
    ctl.base = w & kTileVal;
  }
  named_barrier(2, kBlock);
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
  if (cls == CLS_DEEP) ctl.deep[atomicAdd(&ctl.n_deep, 1)] = (uint8_t)tid;
  named_barrier(2, kBlock);
  uint32_t before = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++) before += w < warp ? ctl.warp_sum[w] : 0u;
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
    s_n[tid] = sink.n;
    if (cnt > sink.room) ++ovf;  // one per clipped board
  }
  named_barrier(2, kBlock);
  for (int k = tid; k < ctl.n_single; k += kBlock) {  // single-jump boards on consecutive threads
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
  if (WITH_MASK) {
    // tree-search boards: the scout has written their capture records; each one gives a mask bit
    bool rewalk = false;
    int t = 0;
    const int k = kBlock - 1 - tid;  // the LAST threads take them (the first ones are busy with the single-jump boards)
    if (k < ctl.n_deep) {
      t = ctl.deep[k];
      const uint64_t o2 = gbase + s_off[t];
      const uint32_t room = (uint32_t)emit_room(o2, o2 + s_cnt[t], a.moves_cap);
      if (room < s_cnt[t]) {
        rewalk = true;  // clipped: the records are not all there
      } else {
        const uint4* src = reinterpret_cast<const uint4*>(a.moves + o2);
        for (uint32_t r = s_n[t]; r < s_cnt[t]; r++) {
          const uint4 lo = __ldcg(src + 2 * r), hi = __ldcg(src + 2 * r + 1);
          Rec rec;
          rec.w[0] = lo.x; rec.w[1] = lo.y; rec.w[2] = lo.z; rec.w[3] = lo.w;
          rec.w[4] = hi.x; rec.w[5] = hi.y; rec.w[6] = hi.z; rec.w[7] = hi.w;
          const uint32_t b = (uint32_t)t * kRow + (uint32_t)(rec_from(rec) * S + rec_to(rec));
          atomicOr(&s_bits[b >> 5], 1u << (b & 31));
        }
      }
    }
    // clipped boards (the caller's record buffer is too small; reported in `overflow`): the mask does not depend on
    // the room, so their trees are walked again here, a few lanes at a time
    unsigned todo = __ballot_sync(0xFFFFFFFFu, rewalk);
    while (todo) {
      const int slot = __popc(todo & ((1u << lane) - 1));
      if (rewalk && slot < kTileRewalkers / kWarps) {
        const Pos<R> q = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[base + t] & 1, base + t);
        MenJumps<R> mj;
        const auto c = find_capturers<R>(q, T, mj);
        BitSink<S> sink;
        sink.out = a.moves;
        sink.bits = mbits;
        sink.bit0 = (uint32_t)t * kRow;
        sink.n = 0;
        sink.room = 0;
        uint32_t o = 0;
        emit_captures<R>(q, c, sink, s_restack + warp * (kTileRewalkers / kWarps) + slot, kTileRewalkers, T, o);
        rewalk = false;
      }
      todo = __ballot_sync(0xFFFFFFFFu, rewalk);
    }
  }
  named_barrier(2, kBlock);
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
    named_barrier(2, kBlock);  // the mask bit strings are dead from here on: their memory holds the tensor bit strings
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
