// drg_fused.cuh -- the fused call (drg_movegen: counts, offsets, records, mask rows, tensor planes) as ONE kernel.
//
// Replaces, for ordinary batches, the two-stream pipeline of drg_kernels.cuh (k_emit<rows> beside k_count -> k_walk_root
// -> scan x3 -> k_emit<records> -> k_emit_deep -> k_mask_fixup: the move generation ran three times and a dozen launches
// shared the SMs with the 2.6 GB row writer).  Here a block owns a 128-board tile from its boards to its last byte:
//
//   count    phase A  own board: exact capture pre-check, quiet moves counted with popc (nothing is enumerated)
//            phase B  capture boards re-dealt onto consecutive threads: single-jump boards counted bit-parallel
//            phase C  the tile's tree-search boards (7 % of random-play Standard positions) walked HERE, a few lanes of
//                     every warp (statistics pass with a descent budget; see "monsters" below)
//   offsets  block scan of the 128 counts, then a decoupled look-back over the tiles before this one (one 64-bit status
//            word per tile: flag | value; tiles are numbered by an atomic ticket, so every predecessor is running or done)
//   emit     the moves are enumerated ONCE, at their final place: records as 32-byte full-sector stores, mask bits into
//            the tile's bit strings in shared memory (tree-search boards too: no deferred mask bytes)
//   rows     each warp widens its bit string to bytes and streams its 32 rows with coalesced 16-byte stores; tensor planes
//            as float4s
//
// Monsters.  A tree that exceeds the inline budget (walk_budget(0): no ordinary position does) is still walked to the
// end by its lane -- results never depend on the budget -- but the kernel raises a word in mapped host memory; the host
// side (drg_abi.cu, pipeline_for) then routes the following calls of that rule family through the split-walk pipeline,
// which deals such trees over thousands of lanes (k_walk_root / k_walk_rounds), until they stop appearing.
#pragma once
#include "drg_kernels.cuh"

namespace drg {

struct TileArgs {
  int64_t n;
  const uint64_t *wm, *wk, *bm, *bk;
  const uint8_t* turn;
  uint32_t* counts;
  uint64_t* offsets;              // [n + 1]
  DrgMove* moves;
  uint64_t moves_cap;
  uint8_t* mask;                  // [n, S*S] (WITH_MASK)
  float* tensor;                  // [n, 4, S] (WITH_TENSOR)
  unsigned long long* overflow;   // boards whose records were clipped / whose capture path exceeds a record
  unsigned long long* tile_state; // [tiles] look-back words, zero before the launch
  unsigned* ticket;               // zero before the launch
  unsigned* monster;              // mapped host word: a tree exceeded the inline budget
};

constexpr unsigned long long kTileAgg = 1ull << 62;   // the tile's own total is known
constexpr unsigned long long kTilePre = 2ull << 62;   // ... and the total of everything before it: value = inclusive prefix
constexpr unsigned long long kTileVal = kTileAgg - 1;

__device__ __forceinline__ unsigned long long tile_load(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tile_store(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct TileCtl {
  int n_cap, n_single, n_deep;
  unsigned tile;
  unsigned long long base;       // records before this tile
  uint32_t warp_sum[kWarps];
  uint8_t cap[kBlock];           // boards with captures (phase A -> B)
  uint8_t single[kBlock];        // ... whose captures are single man jumps
  uint8_t deep[kBlock];          // ... that need the tree search
};

constexpr int kTileWalkers = 32;                      // tree-search lanes per tile: lanes 0..7 of each of the 4 warps
constexpr int kTileWalkLanes = kTileWalkers / kWarps;

template <int S, bool WITH_MASK, bool WITH_TENSOR>
constexpr size_t tile_smem_bytes() {
  size_t b = sizeof(Tables<S>) + ((sizeof(TileCtl) + 15) & ~(size_t)15) + 4 * sizeof(uint32_t) * kBlock +
             sizeof(Frame) * kMaxLevels * kTileWalkers;
  b = (b + 15) & ~(size_t)15;
  const size_t mask_b = WITH_MASK ? (size_t)emit_bits_words<S>() * 4 : 0;
  const size_t tens_b = WITH_TENSOR ? (size_t)kBlock * 8 * 4 : 0;
  return b + (mask_b > tens_b ? mask_b : tens_b) + 16;
}

template <class R, bool WITH_MASK, bool WITH_TENSOR>
__global__ void __launch_bounds__(kBlock) k_movegen_tile(TileArgs a) {
  using BB = typename Pos<R>::BB;
  constexpr int S = R::S;
  constexpr int kRow = S * S;
  extern __shared__ __align__(16) uint8_t smem[];
  Tables<S>& T = *reinterpret_cast<Tables<S>*>(smem);
  TileCtl& ctl = *reinterpret_cast<TileCtl*>(smem + sizeof(Tables<S>));
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(&ctl) + ((sizeof(TileCtl) + 15) & ~15));
  uint32_t* s_off = s_cnt + kBlock;   // first record of the board within the tile
  uint32_t* s_n = s_off + kBlock;     // capture boards: records written before the captures (quiet moves, american.py:96)
  uint32_t* s_bw = s_n + kBlock;      // tree-search board k of the tile: best metric << 1 | king preference
  Frame* s_stack = reinterpret_cast<Frame*>(s_bw + kBlock);
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(
      (reinterpret_cast<uintptr_t>(s_stack + kMaxLevels * kTileWalkers) + 15) & ~(uintptr_t)15);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    ctl.tile = atomicAdd(a.ticket, 1u);
    ctl.n_cap = ctl.n_single = ctl.n_deep = 0;
  }
  load_tables<S>(&T);
  if (WITH_MASK) {
    uint4* z = reinterpret_cast<uint4*>(s_bits);
    for (int q = tid; q < emit_bits_words<S>() / 4; q += kBlock) z[q] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const int64_t tile = ctl.tile;
  const int64_t base = tile * kBlock;
  const int64_t i = base + tid;
  NullSink ns;
  uint32_t ovf = 0;

  // ---- count, phase A ----
  Pos<R> p;
  p.wm = p.wk = p.bm = p.bk = 0;
  p.turn = 0;
  bool has_cap = false;
  {
    uint32_t q = 0;
    if (i < a.n) {
      p = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[i] & 1, i);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(p, T, mj);
      has_cap = c.any();
      if (R::kOptionalCapture || !has_cap) q = (uint32_t)gen_quiet<R, true>(p, ns, T);
      if (has_cap) ctl.cap[atomicAdd(&ctl.n_cap, 1)] = (uint8_t)tid;
    }
    s_cnt[tid] = q;
  }
  __syncthreads();
  // ---- count, phase B: capture boards on consecutive threads ----
  for (int k = tid; k < ctl.n_cap; k += kBlock) {
    const int t = ctl.cap[k];
    const Pos<R> q = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[base + t] & 1, base + t);
    MenJumps<R> mj;
    const auto c = find_capturers<R>(q, T, mj);
    if (!c.kings && only_single_jumps<R>(q, mj)) {
      s_cnt[t] += (uint32_t)single_jumps<R, false>(q, mj, ns, T);
      ctl.single[atomicAdd(&ctl.n_single, 1)] = (uint8_t)t;
    } else {
      ctl.deep[atomicAdd(&ctl.n_deep, 1)] = (uint8_t)t;
    }
  }
  __syncthreads();
  // ---- count, phase C: the tile's tree-search boards, walker k on lane k / 4 of warp k % 4 ----
  const int n_deep = ctl.n_deep;
  const int walker = lane * kWarps + warp;  // < kTileWalkers for the walking lanes
  Frame* stack = s_stack + walker;
  if (lane < kTileWalkLanes) {
    for (int k = walker; k < n_deep; k += kTileWalkers) {
      const int t = ctl.deep[k];
      const Pos<R> q = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[base + t] & 1, base + t);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(q, T, mj);
      WalkLane<R, 0> wl;
      wl.start_root(q, c, 0u, walk_budget(0), T);
      for (;;) {
        const int s = wl.advance(ns, stack, kTileWalkers, T, ovf);
        if (s == kWalkDone) break;
        if (s == kWalkWantsSplit) {  // a monster: finished here all the same; the host hears about it
          wl.budget = kNoBudget;
          *reinterpret_cast<volatile unsigned*>(a.monster) = 1u;
        }
      }
      const uint32_t bw = walk_board_word<R>(wl.st);
      s_bw[k] = bw;
      s_cnt[t] += walk_kept<R>(wl.result(), bw);
    }
  }
  __syncthreads();

  // ---- offsets: block scan, then look-back over the earlier tiles ----
  const uint32_t cnt = s_cnt[tid];
  uint32_t inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) ctl.warp_sum[warp] = inc;
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++) {
    const uint32_t v = ctl.warp_sum[w];
    before += w < warp ? v : 0u;
    total += v;
  }
  const uint32_t local = before + inc - cnt;
  s_off[tid] = local;
  if (warp == 0) {
    unsigned long long excl = 0;
    if (tile > 0) {
      if (lane == 0) tile_store(a.tile_state + tile, kTileAgg | (unsigned long long)total);
      for (int64_t look = tile - 1;; look -= 32) {
        const int64_t idx = look - lane;
        unsigned long long v = idx >= 0 ? tile_load(a.tile_state + idx) : kTilePre;
        // every predecessor holds a ticket, so it is running or done: its word arrives.  The spin limit only keeps a
        // logic error from hanging the device (it then shows as wrong offsets in the tests)
        for (unsigned spins = 0; __any_sync(0xFFFFFFFFu, (v >> 62) == 0) && spins < (1u << 22); spins++) {
          if ((v >> 62) == 0) {
            __nanosleep(40);
            v = tile_load(a.tile_state + idx);
          }
        }
        const unsigned pre = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
        const int first = pre ? __ffs((int)pre) - 1 : 31;  // nearest tile whose inclusive prefix is known
        unsigned long long part = lane <= first ? (v & kTileVal) : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) part += __shfl_down_sync(0xFFFFFFFFu, part, o);
        excl += __shfl_sync(0xFFFFFFFFu, part, 0);
        if (pre) break;
      }
    }
    if (lane == 0) {
      tile_store(a.tile_state + tile, kTilePre | (excl + (unsigned long long)total));
      ctl.base = excl;
    }
  }
  __syncthreads();
  const uint64_t gbase = ctl.base;
  const uint64_t off = gbase + local;
  if (i < a.n) {
    a.counts[i] = cnt;
    a.offsets[i] = off;
    if (i == a.n - 1) a.offsets[a.n] = off + cnt;
  }

  // ---- emit: every move once, at its place ----
  uint32_t* mbits = WITH_MASK ? s_bits : nullptr;
  if (i < a.n) {
    BitSink<S> sink;
    const uint64_t room64 = emit_room(off, off + cnt, a.moves_cap);
    sink.out = a.moves + off;
    sink.bits = mbits;
    sink.bit0 = (uint32_t)tid * kRow;
    sink.n = 0;
    sink.room = (uint32_t)room64;
    if (R::kOptionalCapture || !has_cap) gen_quiet<R, false>(p, sink, T);
    s_n[tid] = sink.n;
    if (cnt > sink.room) ++ovf;  // one per clipped board
  }
  __syncthreads();
  for (int k = tid; k < ctl.n_single; k += kBlock) {
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
  if (lane < kTileWalkLanes) {
    NoAlloc noalloc;
    for (int k = walker; k < n_deep; k += kTileWalkers) {
      const int t = ctl.deep[k];
      const Pos<R> q = load_pos<R>(a.wm, a.wk, a.bm, a.bk, a.turn[base + t] & 1, base + t);
      MenJumps<R> mj;
      const auto c = find_capturers<R>(q, T, mj);
      const uint64_t o2 = gbase + s_off[t];
      BitSink<S> sink;
      sink.out = a.moves + o2;
      sink.bits = mbits;
      sink.bit0 = (uint32_t)t * kRow;
      sink.n = s_n[t];
      sink.room = (uint32_t)emit_room(o2, o2 + s_cnt[t], a.moves_cap);
      const uint32_t bw = s_bw[k];
      WalkLane<R, 1> wl;
      wl.start_root(q, c, 0u, kNoBudget, T);
      wl.st.best = (int)(bw >> 1);
      wl.st.n_best_king = (int)(bw & 1u);
      wl.run(sink, stack, kTileWalkers, T, ovf, noalloc);
    }
  }
  __syncthreads();
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
  if (WITH_TENSOR && nboards) {
    // to_tensor(perspective = side to move) (base.py:753-805): 4 planes of S floats per board, packed into a 4*S-bit
    // string (8 words) per board, then the warp streams float4s
    __syncthreads();  // the mask bit strings are dead from here on: their memory holds the tensor bit strings
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

}  // namespace drg
