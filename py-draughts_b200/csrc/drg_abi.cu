// drg_abi.cu -- the C-ABI of libdrg_b200.so (include/drg.h): context, launch wrappers, host-buffer
// entry points and the perft driver.  No torch, no C++ types across the boundary.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <cmath>
#include <mutex>
#include <vector>

#include "../../include/drg.h"
#include "drg_kernels.cuh"

using namespace drg;

namespace drg_host {  // drg_host.cpp: transport decoder of the packed mask rows
void expand_pipelined(const uint8_t* bits, uint8_t* out, size_t total, size_t chunk, int nchunks,
                      void (*wait_chunk)(void*, int), void* ctx, int threads);
int default_threads();
}  // namespace drg_host

namespace {

thread_local char t_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof t_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(DRG_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                         \
  } while (0)

// Host-buffer entry points queue asynchronous copies into the CALLER's memory; whatever way such a call returns (an
// error in the middle included), nothing may still be in flight towards buffers the caller is free to release
struct DrainOnExit {
  cudaStream_t st;
  ~DrainOnExit() { cudaStreamSynchronize(st); }
};

struct DevBuf {  // grow-only device scratch buffer
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return DRG_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      cudaGetLastError();  // a failed allocation may be tolerated by the caller: do not leave it as the "last error"
      return fail(DRG_E_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    cap = want;
    return DRG_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() { return reinterpret_cast<T*>(p); }
};

struct Context {
  bool inited = false;
  int device = -1;
  int generation = 0;   // bumped whenever the context is bound to a device: per-device function attributes are set again
  int sm_count = 0;
  int64_t launches = 0;
  std::mutex mu;  // host-buffer entry points share the scratch buffers
  DevBuf scan_tiles;  // tile sums of the offsets scan
  DevBuf ctr;         // device counters
  DevBuf in, counts, offsets, moves, mask, tensor, misc;  // host-API staging
  DevBuf maskbits;                                        // bit-packed mask rows (PCIe transport)
  DevBuf deep, deep_stat;                                 // grid-wide list of tree-search boards + their statistics
  DevBuf pst;                                             // piece-square tables of drg_evaluate: [4*50 | 4*32] doubles
  DevBuf fix;                                             // deferred mask bytes of k_emit_deep (MaskFix)
  DevBuf ckeep, cpos;                                     // drg_selfplay_compact: running flags and their scan
  DevBuf pdeep[2], level_alt;                             // perft leaf ply: tree-search lists of two chunks in flight, second leaf frontier
  cudaEvent_t pf_count[2] = {}, pf_deep[2] = {};          // ... and the events that order them between the two streams
  DevBuf xpool;                                           // children kept by the counting walk of k_expand_deep
  DevBuf pool;                                            // leaves kept by k_walk_root for k_emit_deep of the same batch
  unsigned pool_boards = 0;                               // ... kKeepRecords records for each of the first pool_boards list entries
  DevBuf shard;                                           // this rank's share of the perft frontier that is dealt
  DevBuf wrecs, wres, wsub, wbase, rres, bword, wctr;     // split walk: record pool, results, tree sums, per-board words, counters
  cudaStream_t side = nullptr;                            // k_emit_deep runs here, next to k_emit
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_total = nullptr;
  bool profile = false;                                   // drg_profile(): CUDA events around the k_emit launch
  cudaEvent_t prof_ev[2] = {};
  uint8_t* h_bits = nullptr;                              // pinned landing buffer of the packed rows
  size_t h_bits_cap = 0;
  // pinned landing words for scalars read back mid-call: a device-to-host copy into PAGEABLE memory only returns
  // once everything queued before it has finished, which would serialise the transfers with the host work
  unsigned long long* h_word = nullptr;
  uint8_t* h_small = nullptr;                             // pinned + mapped: inputs and result block of the small paths
  size_t h_small_cap = 0;
  cudaEvent_t chunk_ev[64] = {};                          // one event per packed-mask D2H chunk

  std::vector<DevBuf> levels;                             // perft frontiers
} g;
constexpr int kMaxChunks = 64;


// AlphaBetaEngine._get_pst_tables (alpha_beta.py:27-57, 166-177): man black, man white (reversed), king black,
// king white (reversed).  Each line is one IEEE double operation of the Python expression, in its order.
void build_pst(int S, double* out) {
  const int rows = S == 50 ? 10 : 8;  // alpha_beta.py:271-276
  const int per_row = S / rows;
  for (int i = 0; i < S; i++) {
    const int row = i / per_row, col = i % per_row;
    const double half = (double)per_row / 2.0;
    const double advancement = (double)(rows - 1 - row) / (double)(rows - 1) * 0.3;
    const double centre = (1.0 - std::fabs((double)col - half) / half) * 0.05;
    out[i] = advancement + centre;
    const double mid_row = (double)rows / 2.0;
    const double row_dist = std::fabs((double)row - mid_row) / mid_row;
    const double col_dist = std::fabs((double)col - half) / half;
    out[2 * S + i] = (1.0 - (row_dist + col_dist) / 2.0) * 0.25;
  }
  for (int i = 0; i < S; i++) {
    out[S + i] = out[S - 1 - i];
    out[3 * S + i] = out[2 * S + S - 1 - i];
  }
}

enum Family { F_STANDARD, F_AMERICAN, F_FRISIAN, F_RUSSIAN, F_BRAZILIAN, F_BAD };

Family family_of(int variant) {
  switch (variant) {
    case DRG_STANDARD: case DRG_ANTIDRAUGHTS: case DRG_BREAKTHROUGH: return F_STANDARD;
    case DRG_AMERICAN: return F_AMERICAN;
    case DRG_FRISIAN: case DRG_FRYSK: return F_FRISIAN;
    case DRG_RUSSIAN: return F_RUSSIAN;
    case DRG_BRAZILIAN: return F_BRAZILIAN;
  }
  return F_BAD;
}

int draw_family(int variant) {
  switch (family_of(variant)) {
    case F_STANDARD: return DRAW_STANDARD;
    case F_AMERICAN: return DRAW_AMERICAN;
    case F_FRISIAN: return DRAW_FRISIAN;
    default: return DRAW_RUSSIAN;
  }
}

// call `fn.template operator()<Rules>()` for the rule set of `variant`
template <class Fn>
int dispatch(int variant, Fn&& fn) {
  switch (family_of(variant)) {
    case F_STANDARD: return fn.template operator()<RulesStandard>();
    case F_AMERICAN: return fn.template operator()<RulesAmerican>();
    case F_FRISIAN: return fn.template operator()<RulesFrisian>();
    case F_RUSSIAN: return fn.template operator()<RulesRussian>();
    case F_BRAZILIAN: return fn.template operator()<RulesBrazilian>();
    default: return fail(DRG_E_ARG, "unknown variant id %d", variant);
  }
}

int need_init() {
  if (!g.inited) return fail(DRG_E_CUDA, "drg_init() has not been called (or failed): no CUDA context, and there is no CPU fallback");
  return DRG_OK;
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

int launch_check(const char* what) {
  g.launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DRG_E_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
  return DRG_OK;
}

// device counters: [0] perft total, [1] overflow, [2] illegal, [3] expand cursor, [4] movegen positions, [5] deep list size
enum { K_TOTAL = 0, K_OVF = 1, K_ILLEGAL = 2, K_CURSOR = 3, K_MOVEGEN = 4, K_DEEP = 5, K_CAPS = 6, K_FIX = 7, K_PDEEP0 = 8, K_PDEEP1 = 9, K_NCTR = 10 };

// scratch of the grid-wide tree-search list (DeepList) shared by the count kernels; stream-ordered use
int deep_list(int64_t n, DeepList& dl, cudaStream_t st) {
  if (n >= (int64_t)1 << 32) return fail(DRG_E_ARG, "at most 2^32-1 boards per call");
  int rc = g.deep.ensure((size_t)n * sizeof(uint32_t));
  if (rc) return rc;
  dl.items = g.deep.as<uint32_t>();
  dl.n = reinterpret_cast<unsigned*>(g.ctr.as<unsigned long long>() + K_DEEP);
  CUDA_TRY(cudaMemsetAsync(dl.n, 0, sizeof(unsigned), st));
  return DRG_OK;
}

inline unsigned deep_grid() { return (unsigned)(g.sm_count * 8); }
// Grid of the tree-search kernels that run BESIDE the mask-row kernel (fused drg_movegen).  That kernel keeps four 46 KB
// blocks per SM resident; eight 16 KB tree-search blocks per SM take shared memory from them, but a smaller grid makes
// the chain longer than the row kernel, which costs more (sweep in profiles/r2_fused_knob_sweep.txt: 8 per SM 0.560 ms,
// 4: 0.559, 3: 0.569, 2: 0.573 per step) -- so the stand-alone grid stays the default.
inline unsigned deep_grid_beside() {
  static const unsigned per_sm = [] { const char* e = getenv("DRG_DEEP_BLOCKS_BESIDE"); return e ? (unsigned)atoi(e) : 8u; }();
  return (unsigned)g.sm_count * (per_sm ? per_sm : 8u);
}

// DRG_TIMELINE=1 (measurement aid, synchronises): when each stage of the fused call's two streams finished, relative to
// the fork, printed to stderr
struct Timeline {
  static constexpr int kMax = 16;
  cudaEvent_t ev[kMax] = {};
  const char* name[kMax] = {};
  int n = 0;
  bool on = false;
  void begin() { on = getenv("DRG_TIMELINE") != nullptr; n = 0; }
  void mark(const char* what, cudaStream_t st) {
    if (!on || n >= kMax) return;
    if (!ev[n]) cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], st);
    name[n++] = what;
  }
  void print() {
    if (!on || n < 2) return;
    cudaEventSynchronize(ev[n - 1]);
    cudaDeviceSynchronize();
    for (int i = 1; i < n; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[0], ev[i]);
      fprintf(stderr, "%s %.1f us%s", name[i], ms * 1000.f, i + 1 < n ? " | " : "\n");
    }
    on = false;
  }
} g_tl;

// ---- split walk of the tree-search boards (drg_kernels.cuh: k_walk_root / k_walk_rounds / k_emit_deep / k_emit_recs) ----
// counters: count[kWalkRounds + 1] | cursor[kWalkRounds + 4] | barrier | limit[kWalkRounds + 1] | budget[kWalkRounds + 1];
// all start at zero
constexpr size_t kWalkCtrZero = (kWalkRounds + 1) + (kWalkRounds + 4) + 1;
constexpr size_t kWalkCtrWords = kWalkCtrZero + 2 * (kWalkRounds + 1) + 1;  // + the scan tail's barrier word

// scratch of one walk pipeline over the tree-search boards of an n-board call; zeroes the counters on `st`
int walk_pool(int64_t n, WalkPool& wp, cudaStream_t st) {
  // records: a board of the worst batches seen (synthetic Frisian king endgames) hands over ~30; never more than 8 Mi
  int64_t cap = 32 * n;
  if (cap < (1 << 16)) cap = 1 << 16;
  if (cap > (8 << 20)) cap = 8 << 20;
  if (const char* e = getenv("DRG_WALK_POOL_CAP")) cap = atoll(e);  // tests: a pool too small for the batch
  int rc;
  if ((rc = g.wctr.ensure(kWalkCtrWords * sizeof(unsigned)))) return rc;
  if ((rc = g.rres.ensure((size_t)n * sizeof(WalkRes)))) return rc;
  if ((rc = g.bword.ensure((size_t)n * sizeof(uint32_t)))) return rc;
  if ((rc = g.deep_stat.ensure((size_t)n * sizeof(uint32_t)))) return rc;
  // the pool is an optimisation: without it (no memory) every walk simply runs without a budget
  if (g.wrecs.ensure((size_t)cap * sizeof(WalkRec)) != DRG_OK || g.wres.ensure((size_t)cap * sizeof(WalkRes)) != DRG_OK ||
      g.wsub.ensure((size_t)cap * sizeof(uint32_t)) != DRG_OK || g.wbase.ensure((size_t)cap * sizeof(uint32_t)) != DRG_OK)
    cap = 0;
  // kept leaves of the root walks (16 records per tree-search board, for at most 4 Mi of them; the rest walk twice)
  const unsigned boards = (unsigned)(n < ((int64_t)4 << 20) ? n : ((int64_t)4 << 20));
  if (g.pool.ensure((size_t)boards * kKeepRecords * sizeof(DrgMove)) == DRG_OK) g.pool_boards = boards;
  else g.pool_boards = 0;
  unsigned* c = g.wctr.as<unsigned>();
  CUDA_TRY(cudaMemsetAsync(c, 0, kWalkCtrWords * sizeof(unsigned), st));
  wp.recs = g.wrecs.as<WalkRec>();
  wp.res = g.wres.as<WalkRes>();
  wp.sub = g.wsub.as<uint32_t>();
  wp.base = g.wbase.as<uint32_t>();
  wp.count = c;
  wp.cursor = c + (kWalkRounds + 1);
  wp.bar = c + (kWalkRounds + 1) + (kWalkRounds + 4);
  wp.limit = c + kWalkCtrZero;
  wp.budget = c + kWalkCtrZero + (kWalkRounds + 1);
  wp.cap = (unsigned)cap;
  return DRG_OK;
}

// the pool as the emit step sees it after a count step of the same batch (nothing is zeroed)
WalkPool walk_pool_view() {
  WalkPool wp;
  unsigned* c = g.wctr.as<unsigned>();
  wp.recs = g.wrecs.as<WalkRec>();
  wp.res = g.wres.as<WalkRes>();
  wp.sub = g.wsub.as<uint32_t>();
  wp.base = g.wbase.as<uint32_t>();
  wp.count = c;
  wp.cursor = c + (kWalkRounds + 1);
  wp.bar = c + (kWalkRounds + 1) + (kWalkRounds + 4);
  wp.limit = c + kWalkCtrZero;
  wp.budget = c + kWalkCtrZero + (kWalkRounds + 1);
  wp.cap = 0;
  return wp;
}

// k_walk_root + k_walk_rounds over the list `dl` (built on `st` before): afterwards g.rres / g.bword / g.deep_stat / the
// pool describe every board's capture moves, and counts[j] (if given) has them added (set_counts: = quiet part + them)
// beside: the call runs beside the mask-row kernel (fused drg_movegen), whose blocks leave 43 KB of shared memory per
// SM: two blocks per SM fit without waiting for it; otherwise as many as are resident (latency-bound walks want warps)
template <class R>
int walk_boards(int64_t n, DeepList dl, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm, const uint64_t* bk,
                const uint8_t* turn, uint32_t* counts, bool set_counts, unsigned long long* ovf, cudaStream_t st,
                bool beside = false, uint64_t* offsets = nullptr) {
  WalkPool wp;
  int rc = walk_pool(n, wp, st);
  if (rc) return rc;
  k_walk_root<R><<<beside ? deep_grid_beside() : deep_grid(), kBlock, 0, st>>>(dl, wm, wk, bm, bk, turn, counts, set_counts, ovf, g.deep_stat.as<uint32_t>(),
                                                                g.pool.as<DrgMove>(), g.pool_boards, g.rres.as<WalkRes>(),
                                                                g.bword.as<uint32_t>(), wp);
  if ((rc = launch_check("k_walk_root"))) return rc;
  g_tl.mark("walk_root", st);
  // cooperative launch: the rounds are separated by a grid barrier, so every block must be resident
  static int per_sm = 0, per_sm_gen = 0;  // one per instantiation
  if (per_sm_gen != g.generation) {
    per_sm_gen = g.generation;
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_walk_rounds<R>, kBlock, 0));
    if (occ < 1) return fail(DRG_E_CUDA, "k_walk_rounds does not fit an SM");
    per_sm = occ;
  }
  const int blocks_per_sm = beside && per_sm > 2 ? 2 : per_sm;
  const WalkRes* rres = g.rres.as<WalkRes>();
  uint32_t* bword = g.bword.as<uint32_t>();
  // offsets given: the kernel ends with the counts -> offsets scan (one grid barrier instead of three more launches)
  const unsigned grid = (unsigned)(blocks_per_sm * g.sm_count);
  ScanTail sc{n, offsets, nullptr, g.wctr.as<unsigned>() + kWalkCtrWords - 1};
  if (offsets) {
    if ((rc = g.scan_tiles.ensure((size_t)(grid + 1) * sizeof(unsigned long long)))) return rc;
    sc.share_sums = g.scan_tiles.as<unsigned long long>();
  }
  void* args[] = {&dl, &wm, &wk, &bm, &bk, &turn, &counts, &set_counts, &ovf, &rres, &bword, &wp, &sc};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_walk_rounds<R>, dim3(grid), dim3(kBlock), args, 0, st));
  return launch_check("k_walk_rounds");
}

struct ChildrenFn {
  int64_t n; const uint64_t *wm, *wk, *bm, *bk; const uint8_t *turn, *clock; const uint64_t* offsets; int64_t total;
  const DrgMove* moves; uint64_t *owm, *owk, *obm, *obk; uint8_t *oturn, *oclock; uint32_t* parent; uint8_t* status;
  cudaStream_t st;
  template <class R> int operator()() {
    k_children<R><<<grid_for(total), kBlock, 0, st>>>(n, wm, wk, bm, bk, turn, clock, offsets, total, moves, owm, owk, obm,
                                                      obk, oturn, oclock, parent, status);
    return launch_check("k_children");
  }
};

struct SmallFn {
  SmallArgs a;
  cudaStream_t st;
  template <class R> int operator()() {
    k_small<R><<<1, kBlock, 0, st>>>(a);
    return launch_check("k_small");
  }
};

struct CountFn {
  int64_t n; const uint64_t *wm, *wk, *bm, *bk; const uint8_t* turn; uint32_t* counts; unsigned long long* ovf; cudaStream_t st;
  unsigned max_blocks = 0;  // 0: one block per 128-board tile; else a grid-stride launch of at most this many blocks
  uint64_t* offsets = nullptr;  // given: the exclusive scan of the counts as well (tail of k_walk_rounds)
  template <class R> int operator()() {
    DeepList dl;
    int rc = deep_list(n, dl, st);
    if (rc) return rc;
    unsigned grid = grid_for(n);
    if (max_blocks && grid > max_blocks) grid = max_blocks;
    k_count<R><<<grid, kBlock, 0, st>>>(n, wm, wk, bm, bk, turn, counts, dl);
    if ((rc = launch_check("k_count"))) return rc;
    g_tl.mark("count", st);
    return walk_boards<R>(n, dl, wm, wk, bm, bk, turn, counts, false, ovf, st, max_blocks != 0, offsets);
  }
};

// emit step of the tree-search boards on `st` (walk_boards ran on the same list): root walks, then the records
template <class R, bool M>
int emit_tree_boards(const EmitArgs& a, DeepList dl, MaskFix fix, cudaStream_t st, bool beside = false) {
  const WalkPool wp = walk_pool_view();
  k_emit_deep<R, M><<<beside ? deep_grid_beside() : deep_grid(), kBlock, 0, st>>>(a, dl, g.deep_stat.as<uint32_t>(), fix, g.pool.as<DrgMove>(),
                                                    g.pool_boards, g.rres.as<WalkRes>(), g.bword.as<uint32_t>(), wp);
  int rc = launch_check("k_emit_deep");
  if (rc) return rc;
  k_emit_recs<R, M><<<deep_grid(), kBlock, 0, st>>>(a, dl, fix, g.bword.as<uint32_t>(), wp, wp.cursor + kWalkRounds);
  return launch_check("k_emit_recs");
}

// capacity of the deferred-mask-byte list: one entry per capture record of a tree-search board, so as many as the
// record buffer holds records suffice (a board whose records were clipped is an overflow anyway)
inline unsigned fix_capacity(int64_t n, uint64_t moves_cap) {
  uint64_t cap = (uint64_t)(4 * n);
  if (moves_cap != ~0ull && moves_cap > cap) cap = moves_cap;
  if (moves_cap == ~0ull) cap = (uint64_t)(64 * n);
  return (unsigned)(cap < 0xFFFFFFF0ull ? cap : 0xFFFFFFF0ull);
}

struct EmitFn {
  int64_t n; const uint64_t *wm, *wk, *bm, *bk; const uint8_t* turn; const uint64_t* offsets; DrgMove* moves;
  uint8_t* mask; float* tensor; uint32_t* counts; unsigned long long* ovf; cudaStream_t st;
  uint64_t moves_cap = ~0ull;
  bool after_count = false;  // the count step of the SAME boards ran just before on this stream: reuse its tree-search
                             // list and statistics (g.deep / g.deep_stat) instead of rebuilding them
  template <class R, bool M, bool T> int go() {
    return after_count ? go2<R, M, T, true>() : go2<R, M, T, false>();
  }
  template <class R, bool M, bool T, bool HL> int go2() {
    if (reinterpret_cast<uintptr_t>(moves) & 31) return fail(DRG_E_ARG, "moves must be 32-byte aligned (one full-sector store per record)");
    if (M && (reinterpret_cast<uintptr_t>(mask) & 15)) return fail(DRG_E_ARG, "mask must be 16-byte aligned");
    if (T && (reinterpret_cast<uintptr_t>(tensor) & 15)) return fail(DRG_E_ARG, "tensor must be 16-byte aligned");
    constexpr size_t smem = emit_smem_bytes<R::S, M, T>();
    static int attr_gen = 0;  // one per instantiation; function attributes belong to the device the context is bound to
    if (attr_gen != g.generation) {
      CUDA_TRY(cudaFuncSetAttribute(k_emit<R, M, T, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_gen = g.generation;
    }
    EmitArgs a{n, wm, wk, bm, bk, turn, offsets, moves, moves_cap, mask, tensor, counts, ovf};
    DeepList dl;
    int rc;
    if (HL) {
      dl.items = g.deep.as<uint32_t>();
      dl.n = reinterpret_cast<unsigned*>(g.ctr.as<unsigned long long>() + K_DEEP);
    } else if ((rc = deep_list(n, dl, st))) {
      return rc;
    }
    MaskFix none{nullptr, nullptr, 0};
    const bool side_by_side = HL && !getenv("DRG_NO_SIDE_STREAM");
    MaskFix fix = none;
    if (side_by_side) {
      // The list exists already (count step), so the tree-search boards need not wait for k_emit: they run on a second
      // stream NEXT TO it (a latency-bound kernel beside a bandwidth-bound one).  Their records and counts are disjoint
      // from what k_emit writes; their mask bytes would land in rows k_emit is still writing, so they are deferred to a
      // list stored by k_mask_fixup after both kernels.
      fix.cap = fix_capacity(n, moves_cap);
      if (M && (rc = g.fix.ensure((size_t)fix.cap * sizeof(uint64_t)))) return rc;
      fix.at = M ? g.fix.as<uint64_t>() : nullptr;
      fix.n = reinterpret_cast<unsigned*>(g.ctr.as<unsigned long long>() + K_FIX);
      CUDA_TRY(cudaMemsetAsync(fix.n, 0, sizeof(unsigned), st));
      CUDA_TRY(cudaEventRecord(g.ev_fork, st));
      CUDA_TRY(cudaStreamWaitEvent(g.side, g.ev_fork, 0));
    }
    if (g.profile) CUDA_TRY(cudaEventRecord(g.prof_ev[0], st));
    k_emit<R, M, T, HL><<<grid_for(n), kBlock, smem, st>>>(a, dl);
    if ((rc = launch_check("k_emit"))) return rc;
    if (g.profile) CUDA_TRY(cudaEventRecord(g.prof_ev[1], st));
    if (side_by_side) {
      // launched AFTER k_emit so that k_emit's blocks (46 KB of shared memory each) take the SMs first and the
      // tree-search blocks fill what is left
      if ((rc = emit_tree_boards<R, M>(a, dl, fix, g.side))) return rc;
      CUDA_TRY(cudaEventRecord(g.ev_join, g.side));
      CUDA_TRY(cudaStreamWaitEvent(st, g.ev_join, 0));
      if (M) {
        k_mask_fixup<<<(unsigned)g.sm_count, 256, 0, st>>>(mask, fix, ovf);
        if ((rc = launch_check("k_mask_fixup"))) return rc;
      }
      return DRG_OK;
    }
    if (!HL) {
      // stand-alone emit: k_emit has just built the list; walk it (root walks, records, tree sums) before emitting
      if ((rc = walk_boards<R>(n, dl, wm, wk, bm, bk, turn, counts, true, ovf, st))) return rc;
    }
    return emit_tree_boards<R, M>(a, dl, none, st);
  }
  template <class R> int operator()() {
    if (mask && tensor) return go<R, true, true>();
    if (mask) return go<R, true, false>();
    if (tensor) return go<R, false, true>();
    return go<R, false, false>();
  }
};

int scan_counts(int64_t n, const uint32_t* counts, uint64_t* offsets, cudaStream_t st);


// The fused call with a mask: the mask rows need neither counts nor offsets, so their kernel (the 2.6 GB writer)
// starts at once on the caller's stream, and the whole count -> scan -> records (+ tensor) -> tree-search chain runs
// BESIDE it on the high-priority side stream; mask bytes of tree-search boards are deferred (MaskFix) and stored
// after both have finished.
struct SplitMovegenFn {
  int64_t n; const uint64_t *wm, *wk, *bm, *bk; const uint8_t* turn; uint32_t* counts; uint64_t* offsets; DrgMove* moves;
  uint64_t moves_cap; uint8_t* mask; float* tensor; unsigned long long* ovf; cudaStream_t st;
  template <class R, bool T> int go() {
    if (reinterpret_cast<uintptr_t>(moves) & 31) return fail(DRG_E_ARG, "moves must be 32-byte aligned (one full-sector store per record)");
    if (reinterpret_cast<uintptr_t>(mask) & 15) return fail(DRG_E_ARG, "mask must be 16-byte aligned");
    if (T && (reinterpret_cast<uintptr_t>(tensor) & 15)) return fail(DRG_E_ARG, "tensor must be 16-byte aligned");
    constexpr size_t smem_mask = emit_smem_bytes<R::S, true, false>();
    constexpr size_t smem_rec = emit_smem_bytes<R::S, false, T>();
    static int attr_gen = 0;
    if (attr_gen != g.generation) {
      CUDA_TRY(cudaFuncSetAttribute(k_emit<R, true, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mask));
      CUDA_TRY(cudaFuncSetAttribute(k_emit<R, false, T, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rec));
      attr_gen = g.generation;
    }
    int rc;
    MaskFix fix;
    fix.cap = fix_capacity(n, moves_cap);
    if ((rc = g.fix.ensure((size_t)fix.cap * sizeof(uint64_t)))) return rc;
    if ((rc = g.deep.ensure((size_t)n * sizeof(uint32_t)))) return rc;       // before the fork: no allocation between
    if ((rc = g.deep_stat.ensure((size_t)n * sizeof(uint32_t)))) return rc;  // launches of the two streams
    fix.at = g.fix.as<uint64_t>();
    fix.n = reinterpret_cast<unsigned*>(g.ctr.as<unsigned long long>() + K_FIX);
    CUDA_TRY(cudaMemsetAsync(fix.n, 0, sizeof(unsigned), st));
    // DRG_SPLIT_SERIAL=1 (measurement aid): same kernels, one after the other on the caller's stream
    const bool serial = getenv("DRG_SPLIT_SERIAL") != nullptr;
    cudaStream_t chain = serial ? st : g.side;
    g_tl.begin();
    g_tl.mark("fork", st);
    if (!serial) {
      CUDA_TRY(cudaEventRecord(g.ev_fork, st));
      CUDA_TRY(cudaStreamWaitEvent(g.side, g.ev_fork, 0));
    }
    // caller's stream: every mask row (quiet moves + single jumps; tree-search boards get their capture bytes later)
    DeepList none{nullptr, nullptr};
    EmitArgs am{n, wm, wk, bm, bk, turn, nullptr, nullptr, 0, mask, nullptr, nullptr, nullptr};
    if (g.profile) CUDA_TRY(cudaEventRecord(g.prof_ev[0], st));
    // DRG_ROW_PAD_KB (measurement aid): extra dynamic shared memory per row-kernel block = fewer resident blocks per SM
    // Four row blocks per SM, not the five that fit since the geometry tables shrank (1 KB of padding): five are 0.7 %
    // faster on most GPUs (0.547 against 0.551 ms) and 5-7 % slower on others -- on one 8-GPU box three GPUs ran the step
    // in 0.575-0.589 ms with five blocks and all eight in 0.550-0.560 with four (profiles/README.md); the chain-bound
    // families (Frisian: 1.005 against 0.973 ms) prefer four anyway.  DRG_ROW_PAD_KB=0 restores five.
    static const size_t row_pad = [] { const char* e = getenv("DRG_ROW_PAD_KB"); return e ? (size_t)atoi(e) * 1024 : (size_t)1024; }();
    if (row_pad) CUDA_TRY(cudaFuncSetAttribute(k_emit<R, true, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_mask + row_pad)));
    k_emit<R, true, false, true, false><<<grid_for(n), kBlock, smem_mask + row_pad, st>>>(am, none);
    if ((rc = launch_check("k_emit(mask)"))) return rc;
    if (g.profile) CUDA_TRY(cudaEventRecord(g.prof_ev[1], st));
    g_tl.mark("rows", st);
    // side stream (high priority): counts, offsets, records, tensor, tree-search boards
    // k_count's resident blocks per SM: 6 leave the row kernel its four 46 KB blocks at full size (sweep in
    // profiles/r2_fused_knob_sweep.txt); below ~3/4 Mi boards the chain is the longer of the two streams (its kernels are
    // latency-bound: k_count takes ~18 us per tile and block), so it gets twice that (DRG_CHAIN_BLOCKS_SMALL)
    static const unsigned per_sm_big = [] { const char* e = getenv("DRG_CHAIN_BLOCKS"); return e ? (unsigned)atoi(e) : 6u; }();
    static const unsigned per_sm_small = [] { const char* e = getenv("DRG_CHAIN_BLOCKS_SMALL"); return e ? (unsigned)atoi(e) : 12u; }();
    const unsigned per_sm = n >= (int64_t)3 << 18 ? per_sm_big : per_sm_small;
    DeepList dl;
    EmitArgs ar{n, wm, wk, bm, bk, turn, offsets, moves, moves_cap, nullptr, tensor, nullptr, ovf};
    static const unsigned rec_per_sm = [] { const char* e = getenv("DRG_CHAIN_REC_BLOCKS"); return e ? (unsigned)atoi(e) : 0u; }();
    const unsigned chain_cap = rec_per_sm ? rec_per_sm * (unsigned)g.sm_count : 0xFFFFFFFFu;
    const unsigned chain_grid = grid_for(n) < chain_cap ? grid_for(n) : chain_cap;
    {
      static const bool tail_scan = getenv("DRG_NO_TAIL_SCAN") == nullptr;
      CountFn cf{n, wm, wk, bm, bk, turn, counts, ovf, chain, serial ? 0u : per_sm * (unsigned)g.sm_count, tail_scan ? offsets : nullptr};
      if ((rc = cf.template operator()<R>())) return rc;
      g_tl.mark("rounds", chain);
      if (!tail_scan && (rc = scan_counts(n, counts, offsets, chain))) return rc;
      g_tl.mark("scan", chain);
      dl.items = g.deep.as<uint32_t>();
      dl.n = reinterpret_cast<unsigned*>(g.ctr.as<unsigned long long>() + K_DEEP);
      k_emit<R, false, T, true, true><<<chain_grid, kBlock, smem_rec, chain>>>(ar, dl);
      if ((rc = launch_check("k_emit(records)"))) return rc;
    }
    g_tl.mark("records", chain);
    EmitArgs ad{n, wm, wk, bm, bk, turn, offsets, moves, moves_cap, mask, nullptr, nullptr, ovf};
    if ((rc = emit_tree_boards<R, true>(ad, dl, fix, chain, !serial))) return rc;
    g_tl.mark("tree", chain);
    if (!serial) {
      CUDA_TRY(cudaEventRecord(g.ev_join, g.side));
      CUDA_TRY(cudaStreamWaitEvent(st, g.ev_join, 0));
    }
    k_mask_fixup<<<(unsigned)g.sm_count, 256, 0, st>>>(mask, fix, ovf);
    rc = launch_check("k_mask_fixup");
    g_tl.mark("fixup", st);
    g_tl.print();
    return rc;
  }
  template <class R> int operator()() { return tensor ? go<R, true>() : go<R, false>(); }
};

struct ApplyFn {
  int64_t n; uint64_t *wm, *wk, *bm, *bk; uint8_t *turn, *clock; const DrgMove* chosen; uint8_t* status; cudaStream_t st;
  template <class R> int operator()() {
    k_apply<R><<<grid_for(n), kBlock, 0, st>>>(n, wm, wk, bm, bk, turn, clock, chosen, status);
    return launch_check("k_apply");
  }
};

// Leaf ply of perft.  The tree-search boards of a chunk (k_perft_count_deep: few warps, long dependent walks) run on
// the high-priority side stream BESIDE the next chunk's expansion and count on the caller's stream -- both are
// issue-bound at 50-59 % of the slots when they run alone.  Two chunks are in flight: list / frontier buffer `slot`
// is reused only after the event of its tree-search kernel (perft_rec).
struct PerftCountFn {
  int64_t n; const uint64_t *wm, *wk, *bm, *bk; int turn; unsigned long long *total, *ovf; cudaStream_t st;
  int slot = 0;
  bool overlap = false;
  template <class R> int operator()() {
    int64_t blocks = (n + kBlock - 1) / kBlock;
    const int64_t maxb = (int64_t)g.sm_count * 64;
    if (blocks > maxb) blocks = maxb;
    if (n >= (int64_t)1 << 32) return fail(DRG_E_ARG, "at most 2^32-1 boards per call");
    int rc = g.pdeep[slot].ensure((size_t)n * sizeof(uint32_t));
    if (rc) return rc;
    DeepList dl;
    dl.items = g.pdeep[slot].as<uint32_t>();
    dl.n = reinterpret_cast<unsigned*>(g.ctr.as<unsigned long long>() + (slot ? K_PDEEP1 : K_PDEEP0));
    // overlap: the tree-search kernel of this chunk's leaves goes to the side stream; for the flying-king families without
    // orthogonal captures the count kernel as well (the caller's stream then only expands): Standard d12 0.1326 against
    // 0.1373 s, Russian d12 9.0 against 9.3 ms -- but American d12 0.451 against 0.433 s and Frisian d11 34.3 against
    // 29.2 ms, where the leaf ply is the heavier half.  DRG_PERFT_OVERLAP=deep / leaf forces one form.
    static const int forced = [] { const char* e = getenv("DRG_PERFT_OVERLAP"); return !e ? 0 : (!strcmp(e, "leaf") ? 1 : (!strcmp(e, "deep") ? 2 : 0)); }();
    const bool deep_only = forced ? forced == 2 : !(R::kFlying && !R::kOrtho);
    cudaStream_t cs = st, ds = st;
    if (overlap) {
      for (auto* ev : {&g.pf_count[slot], &g.pf_deep[slot]})
        if (!*ev) CUDA_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
      ds = g.side;
      if (!deep_only) {
        CUDA_TRY(cudaEventRecord(g.pf_count[slot], st));
        CUDA_TRY(cudaStreamWaitEvent(g.side, g.pf_count[slot], 0));
        cs = g.side;
      }
    }
    CUDA_TRY(cudaMemsetAsync(dl.n, 0, sizeof(unsigned), cs));
    k_perft_count<R><<<(unsigned)blocks, kBlock, 0, cs>>>(n, wm, wk, bm, bk, turn, total, dl);
    if ((rc = launch_check("k_perft_count"))) return rc;
    if (overlap && deep_only) {
      CUDA_TRY(cudaEventRecord(g.pf_count[slot], st));
      CUDA_TRY(cudaStreamWaitEvent(g.side, g.pf_count[slot], 0));
    }
    k_perft_count_deep<R><<<deep_grid(), kBlock, 0, ds>>>(dl, wm, wk, bm, bk, turn, total, ovf);
    if ((rc = launch_check("k_perft_count_deep"))) return rc;
    if (overlap) CUDA_TRY(cudaEventRecord(g.pf_deep[slot], g.side));
    return DRG_OK;
  }
};

struct ExpandFn {
  int64_t n; const uint64_t *wm, *wk, *bm, *bk; int turn; uint64_t *owm, *owk, *obm, *obk; uint64_t cap;
  unsigned long long *cursor, *ovf, *illegal; cudaStream_t st;
  template <class R> int operator()() {
    DeepList dl;
    int rc = deep_list(n, dl, st);
    if (rc) return rc;
    k_expand<R><<<grid_for(n), kBlock, 0, st>>>(n, wm, wk, bm, bk, turn, owm, owk, obm, obk, cap, cursor, illegal, dl);
    if ((rc = launch_check("k_expand"))) return rc;
    // thread-private slots for the children kept by the counting walk (16 x 16 bytes per thread of the grid)
    if ((rc = g.xpool.ensure((size_t)deep_grid() * kBlock * kKeepRecords * sizeof(uint4)))) return rc;
    k_expand_deep<R><<<deep_grid(), kBlock, 0, st>>>(dl, wm, wk, bm, bk, turn, owm, owk, obm, obk, cap, cursor, ovf, illegal,
                                                     g.xpool.as<uint4>());
    return launch_check("k_expand_deep");
  }
};

struct RolloutFn {
  int variant; int64_t n; uint64_t seed, gid0; int max_plies; const uint16_t* stop; uint32_t flags;
  uint64_t *wm, *wk, *bm, *bk; uint8_t *turn, *clock, *result; uint16_t* plies; unsigned long long* counters; cudaStream_t st;
  template <class R> int operator()() {
    k_rollout<R><<<grid_for(n), kBlock, 0, st>>>(variant, draw_family(variant), n, seed, gid0, max_plies, stop, flags,
                                                 wm, wk, bm, bk, turn, clock, result, plies, counters);
    return launch_check("k_rollout");
  }
};

struct SelfplayFn {
  int variant; int64_t n; uint64_t seed, gid0; int max_plies; const uint16_t* stop; uint32_t flags;
  uint64_t *wm, *wk, *bm, *bk; uint8_t *turn, *clock; uint16_t* plies; uint8_t* state; uint32_t* hist;
  const uint32_t* counts; const uint64_t* offsets; const DrgMove* moves; uint64_t moves_cap; const int32_t* choice;
  const uint64_t* gids; unsigned long long *running, *illegal; cudaStream_t st;
  template <class R> int operator()() {
    k_selfplay_step<R><<<grid_for(n), kBlock, 0, st>>>(variant, draw_family(variant), n, seed, gid0, max_plies, stop,
                                                     flags, wm, wk, bm, bk, turn, clock, plies, state, hist, counts,
                                                     offsets, moves, moves_cap, choice, gids, running, illegal);
    return launch_check("k_selfplay_step");
  }
};

int scan_counts(int64_t n, const uint32_t* counts, uint64_t* offsets, cudaStream_t st) {
  if (n == 0) {
    CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(uint64_t), st));
    return DRG_OK;
  }
  const int64_t ntiles = (n + kScanTile - 1) / kScanTile;
  int rc = g.scan_tiles.ensure((size_t)(ntiles + 1) * sizeof(unsigned long long));
  if (rc) return rc;
  unsigned long long* ts = g.scan_tiles.as<unsigned long long>();
  k_scan_tiles<<<(unsigned)ntiles, kScanBlock, 0, st>>>(n, counts, ts);
  if ((rc = launch_check("k_scan_tiles"))) return rc;
  k_scan_sums<<<1, kScanBlock, 0, st>>>(ntiles, ts);
  if ((rc = launch_check("k_scan_sums"))) return rc;
  k_scan_apply<<<(unsigned)ntiles, kScanBlock, 0, st>>>(n, counts, ts, offsets, ntiles);
  return launch_check("k_scan_apply");
}


}  // namespace

// =============================================================================================
extern "C" {

int drg_abi_version(void) { return DRG_ABI_VERSION; }

const char* drg_last_error(void) { return t_err; }

int drg_squares(int variant) {
  switch (family_of(variant)) {
    case F_STANDARD: case F_FRISIAN: return 50;
    case F_AMERICAN: case F_RUSSIAN: case F_BRAZILIAN: return 32;
    default: return fail(DRG_E_ARG, "unknown variant id %d", variant);
  }
}

// host copy of the neighbour table (tests compare it with the reference's tables; no GPU needed)
int drg_geometry_table(int squares, uint8_t* nb /*[512]*/) {
  if (!nb) return fail(DRG_E_ARG, "nb is NULL");
  memset(nb, kNone, 512);  // the caller's table has 64 rows; the rows past `squares` stay "off board"
  if (squares == 50) build_nb<50>(nb);
  else if (squares == 32) build_nb<32>(nb);
  else return fail(DRG_E_ARG, "squares must be 50 or 32");
  return DRG_OK;
}

int64_t drg_launch_count(void) { return g.launches; }

int drg_profile(int enable, float* k_emit_ms) {
  int rc = need_init();
  if (rc) return rc;
  for (auto& e : g.prof_ev)
    if (!e) CUDA_TRY(cudaEventCreate(&e));
  if (k_emit_ms) {  // duration of the most recent k_emit launch made while profiling was on (synchronises on it)
    CUDA_TRY(cudaEventSynchronize(g.prof_ev[1]));
    CUDA_TRY(cudaEventElapsedTime(k_emit_ms, g.prof_ev[0], g.prof_ev[1]));
  }
  g.profile = enable != 0;
  return DRG_OK;
}

int drg_memset_async(void* dev_ptr, int value, size_t bytes, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(dev_ptr, value, bytes, (cudaStream_t)stream));
  return DRG_OK;
}

namespace { int release_locked(); }

int drg_init(int device) {
  std::lock_guard<std::mutex> lk(g.mu);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(DRG_E_CUDA, "no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(DRG_E_ARG, "device %d out of range [0,%d)", device, ndev);
  if (g.inited && g.device == device) {
    CUDA_TRY(cudaSetDevice(device));
    return DRG_OK;
  }
  if (g.inited) {
    // One context per process: binding to another device first releases everything held on the old one (scratch
    // buffers, the side stream, events, the pinned staging memory); data the caller keeps there stays the caller's.
    CUDA_TRY(cudaSetDevice(g.device));
    release_locked();
  }
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(DRG_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  g.sm_count = prop.multiProcessorCount;
  {
    static Tables<50> t50;
    static Tables<32> t32;
    build_nb<50>(t50.nb);
    build_nb<32>(t32.nb);
    build_ray_jmp<50>(t50.ray, t50.jmp, t50.nb);
    build_ray_jmp<32>(t32.ray, t32.jmp, t32.nb);
    CUDA_TRY(cudaMemcpyToSymbol(g_tab50, &t50, sizeof t50));
    CUDA_TRY(cudaMemcpyToSymbol(g_tab32, &t32, sizeof t32));
  }
  int rc = g.ctr.ensure(K_NCTR * sizeof(unsigned long long));
  if (rc) return rc;
  {
    double pst[4 * 50 + 4 * 32];
    build_pst(50, pst);
    build_pst(32, pst + 4 * 50);
    if ((rc = g.pst.ensure(sizeof pst))) return rc;
    CUDA_TRY(cudaMemcpy(g.pst.p, pst, sizeof pst, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMemset(g.ctr.p, 0, K_NCTR * sizeof(unsigned long long)));
  if (!g.h_word) CUDA_TRY(cudaMallocHost(&g.h_word, 8 * sizeof(unsigned long long)));
  if (!g.side) {
    int lo = 0, hi = 0;  // numerically lowest = highest priority
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&g.side, cudaStreamNonBlocking, hi));
  }
  if (!g.ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
  if (!g.ev_join) CUDA_TRY(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
  if (!g.ev_total) CUDA_TRY(cudaEventCreateWithFlags(&g.ev_total, cudaEventDisableTiming));
  g.device = device;
  g.inited = true;
  g.generation++;
  g.launches = 0;
  return DRG_OK;
}

namespace {
// everything the context holds on its device (caller holds the lock; the device is current)
int release_locked() {
  cudaDeviceSynchronize();
  for (DevBuf* b : {&g.scan_tiles, &g.ctr, &g.in, &g.counts, &g.offsets, &g.moves, &g.mask, &g.tensor, &g.misc}) b->release();
  for (auto& b : g.levels) b.release();
  g.levels.clear();
  g.maskbits.release();
  g.deep.release();
  g.deep_stat.release();
  g.pst.release();
  if (g.h_bits) cudaFreeHost(g.h_bits);
  g.h_bits = nullptr;
  g.h_bits_cap = 0;
  if (g.h_word) cudaFreeHost(g.h_word);
  g.h_word = nullptr;
  g.fix.release();
  g.ckeep.release();
  g.cpos.release();
  g.pdeep[0].release();
  g.pdeep[1].release();
  g.level_alt.release();
  for (auto* ev : {&g.pf_count[0], &g.pf_count[1], &g.pf_deep[0], &g.pf_deep[1]})
    if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
  g.pool.release();
  g.xpool.release();
  for (DevBuf* b : {&g.wrecs, &g.wres, &g.wsub, &g.wbase, &g.rres, &g.bword, &g.wctr, &g.shard}) b->release();
  g.pool_boards = 0;
  if (g.side) cudaStreamDestroy(g.side);
  g.side = nullptr;
  if (g.ev_fork) cudaEventDestroy(g.ev_fork);
  if (g.ev_join) cudaEventDestroy(g.ev_join);
  if (g.ev_total) cudaEventDestroy(g.ev_total);
  g.ev_fork = g.ev_join = g.ev_total = nullptr;
  if (g.h_small) cudaFreeHost(g.h_small);
  g.h_small = nullptr;
  g.h_small_cap = 0;
  for (auto& e : g.chunk_ev)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  g.inited = false;
  return DRG_OK;
}
}  // namespace

int drg_shutdown(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.inited) return DRG_OK;
  CUDA_TRY(cudaSetDevice(g.device));
  return release_locked();
}

// pinned host memory for the host-buffer entry points (numpy wraps it; avoids pageable staging copies)
void* drg_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail(DRG_E_NOMEM, "cudaMallocHost(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
int drg_host_free(void* p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return DRG_OK;
}

// ---- device entry points -----------------------------------------------------------------------

int drg_movegen_count(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                      const uint64_t* bk, const uint8_t* turn, uint32_t* counts, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !counts))) return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return family_of(variant) == F_BAD ? fail(DRG_E_ARG, "unknown variant id %d", variant) : DRG_OK;
  CountFn f{n, wm, wk, bm, bk, turn, counts, g.ctr.as<unsigned long long>() + K_OVF, (cudaStream_t)stream};
  return dispatch(variant, f);
}

int drg_scan_counts(int64_t n, const uint32_t* counts, uint64_t* offsets, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (n < 0 || !offsets || (n > 0 && !counts)) return fail(DRG_E_ARG, "bad arguments");
  return scan_counts(n, counts, offsets, (cudaStream_t)stream);
}

int drg_movegen_emit(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                     const uint64_t* bk, const uint8_t* turn, const uint64_t* offsets, DrgMove* moves, uint8_t* mask,
                     float* tensor, uint32_t* counts, uint64_t* overflow, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !offsets || !moves))) return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return family_of(variant) == F_BAD ? fail(DRG_E_ARG, "unknown variant id %d", variant) : DRG_OK;
  EmitFn f{n, wm, wk, bm, bk, turn, offsets, moves, mask, tensor, counts,
           overflow ? reinterpret_cast<unsigned long long*>(overflow) : g.ctr.as<unsigned long long>() + K_OVF,
           (cudaStream_t)stream};
  return dispatch(variant, f);
}

int drg_movegen(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm, const uint64_t* bk,
                const uint8_t* turn, uint32_t* counts, uint64_t* offsets, DrgMove* moves, int64_t moves_cap, uint8_t* mask,
                float* tensor, uint64_t* overflow, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (n < 0 || moves_cap < 0 || !offsets || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !counts || !moves)))
    return fail(DRG_E_ARG, "bad arguments");
  if (family_of(variant) == F_BAD) return fail(DRG_E_ARG, "unknown variant id %d", variant);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* ovf = overflow ? reinterpret_cast<unsigned long long*>(overflow) : g.ctr.as<unsigned long long>() + K_OVF;
  if (n > 0 && mask && !getenv("DRG_NO_SPLIT_EMIT")) {
    SplitMovegenFn f{n, wm, wk, bm, bk, turn, counts, offsets, moves, (uint64_t)moves_cap, mask, tensor, ovf, st};
    return dispatch(variant, f);
  }
  if (n > 0) {
    CountFn cf{n, wm, wk, bm, bk, turn, counts, ovf, st, 0u, offsets};  // counts and, as the walk kernel's tail, offsets
    if ((rc = dispatch(variant, cf))) return rc;
  } else if ((rc = scan_counts(n, counts, offsets, st))) {
    return rc;
  }
  if (n == 0) return DRG_OK;
  EmitFn ef{n, wm, wk, bm, bk, turn, offsets, moves, mask, tensor, nullptr, ovf, st, (uint64_t)moves_cap, true};
  return dispatch(variant, ef);
}

int drg_children(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm, const uint64_t* bk,
                 const uint8_t* turn, const uint8_t* clock, const uint64_t* offsets, int64_t total, const DrgMove* moves,
                 uint64_t* owm, uint64_t* owk, uint64_t* obm, uint64_t* obk, uint8_t* oturn, uint8_t* oclock,
                 uint32_t* parent, uint8_t* status, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (family_of(variant) == F_BAD) return fail(DRG_E_ARG, "unknown variant id %d", variant);
  if (n < 0 || total < 0 || n >= (int64_t)1 << 32) return fail(DRG_E_ARG, "bad arguments");
  if (total > 0 && (n == 0 || !wm || !wk || !bm || !bk || !turn || !offsets || !moves || !owm || !owk || !obm || !obk || !oturn))
    return fail(DRG_E_ARG, "bad arguments");
  if (total == 0) return DRG_OK;
  ChildrenFn f{n, wm, wk, bm, bk, turn, clock, offsets, total, moves, owm, owk, obm, obk, oturn, oclock, parent, status,
               (cudaStream_t)stream};
  return dispatch(variant, f);
}

int drg_evaluate(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm, const uint64_t* bk,
                 const uint8_t* turn, double* score, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  const int S = drg_squares(variant);
  if (S < 0) return fail(DRG_E_ARG, "unknown variant id %d", variant);
  if (n < 0 || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !score))) return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return DRG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 50) k_evaluate<50><<<grid_for(n), kBlock, 0, st>>>(n, wm, wk, bm, bk, turn, g.pst.as<double>(), score);
  else k_evaluate<32><<<grid_for(n), kBlock, 0, st>>>(n, wm, wk, bm, bk, turn, g.pst.as<double>() + 4 * 50, score);
  return launch_check("k_evaluate");
}

int drg_pst_table(int variant, double* out) {
  const int S = drg_squares(variant);
  if (S < 0 || !out) return fail(DRG_E_ARG, "bad arguments");
  build_pst(S, out);
  return 4 * S;
}

int drg_apply(int variant, int64_t n, uint64_t* wm, uint64_t* wk, uint64_t* bm, uint64_t* bk, uint8_t* turn,
              uint8_t* clock, const DrgMove* chosen, uint8_t* status, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !chosen))) return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return family_of(variant) == F_BAD ? fail(DRG_E_ARG, "unknown variant id %d", variant) : DRG_OK;
  ApplyFn f{n, wm, wk, bm, bk, turn, clock, chosen, status, (cudaStream_t)stream};
  return dispatch(variant, f);
}

int drg_rollout(int variant, int64_t n_games, uint64_t seed, uint64_t game_id0, int max_plies, const uint16_t* stop_ply,
                uint32_t flags, uint64_t* wm, uint64_t* wk, uint64_t* bm, uint64_t* bk, uint8_t* turn, uint8_t* clock,
                uint8_t* result, uint16_t* plies, uint64_t* counters, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (n_games < 0 || (n_games > 0 && (!wm || !wk || !bm || !bk || !turn || !clock))) return fail(DRG_E_ARG, "bad arguments");
  if (n_games == 0) return family_of(variant) == F_BAD ? fail(DRG_E_ARG, "unknown variant id %d", variant) : DRG_OK;
  RolloutFn f{variant, n_games, seed, game_id0, max_plies, stop_ply, flags, wm, wk, bm, bk, turn, clock, result, plies,
              reinterpret_cast<unsigned long long*>(counters), (cudaStream_t)stream};
  return dispatch(variant, f);
}

int drg_selfplay_step(int variant, int64_t n, uint64_t seed, uint64_t game_id0, int max_plies, const uint16_t* stop_ply,
                      uint32_t flags, uint64_t* wm, uint64_t* wk, uint64_t* bm, uint64_t* bk, uint8_t* turn,
                      uint8_t* clock, uint16_t* plies, uint8_t* state, uint32_t* hist, const uint32_t* counts,
                      const uint64_t* offsets, const DrgMove* moves, int64_t moves_cap, const int32_t* choice,
                      const uint64_t* game_ids, uint64_t* running, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !clock || !plies || !state || !hist || !counts ||
                          !offsets || !moves || !running)))
    return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return family_of(variant) == F_BAD ? fail(DRG_E_ARG, "unknown variant id %d", variant) : DRG_OK;
  if (moves_cap < 0) return fail(DRG_E_ARG, "bad arguments");
  SelfplayFn f{variant, n, seed, game_id0, max_plies, stop_ply, flags, wm, wk, bm, bk, turn, clock, plies, state, hist, counts,
      offsets, moves, (uint64_t)moves_cap, choice, game_ids, reinterpret_cast<unsigned long long*>(running),
      g.ctr.as<unsigned long long>() + K_ILLEGAL, (cudaStream_t)stream};
  return dispatch(variant, f);
}

// ---- NCCL, bound at run time (no link-time dependency: the process may already hold torch's copy) ----------------
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool tried = false, ok = false;
} nccl;

int nccl_bind() {
  if (!nccl.tried) {
    nccl.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy this process has loaded already (torch's)
    if (!h) h = dlopen(nullptr, RTLD_NOW);
    if (!h || !dlsym(h, "ncclAllReduce")) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h || !dlsym(h, "ncclAllReduce")) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      nccl.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(h, "ncclGetUniqueId"));
      nccl.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(h, "ncclCommInitRank"));
      nccl.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(h, "ncclAllReduce"));
      nccl.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
      nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
      nccl.ok = nccl.GetUniqueId && nccl.CommInitRank && nccl.AllReduce && nccl.CommDestroy;
    }
  }
  return nccl.ok ? DRG_OK : fail(DRG_E_CUDA, "no NCCL library could be bound (libnccl.so.2 is neither loaded nor on the loader path)");
}

int nccl_fail(const char* what, int rc) {
  return fail(DRG_E_CUDA, "%s: %s", what, nccl.GetErrorString ? nccl.GetErrorString(rc) : "NCCL error");
}
}  // namespace

int drg_nccl_unique_id(uint8_t* id128) {
  if (!id128) return fail(DRG_E_ARG, "bad arguments");
  int rc = nccl_bind();
  if (rc) return rc;
  NcclId id;
  if ((rc = nccl.GetUniqueId(&id))) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id128, id.internal, sizeof id.internal);
  return DRG_OK;
}

int drg_comm_init(const uint8_t* id128, int rank, int world, void** comm) {
  int rc = need_init();
  if (rc) return rc;
  if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) return fail(DRG_E_ARG, "bad arguments");
  if ((rc = nccl_bind())) return rc;
  CUDA_TRY(cudaSetDevice(g.device));
  NcclId id;
  memcpy(id.internal, id128, sizeof id.internal);
  *comm = nullptr;
  if ((rc = nccl.CommInitRank(comm, world, id, rank))) return nccl_fail("ncclCommInitRank", rc);
  return DRG_OK;
}

int drg_allreduce_counters(void* comm, uint64_t* counters, int n, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (!comm || n < 0 || (n > 0 && !counters)) return fail(DRG_E_ARG, "bad arguments");
  if ((rc = nccl_bind())) return rc;
  if (n == 0) return DRG_OK;
  constexpr int kNcclUint64 = 5, kNcclSum = 0;  // nccl.h: ncclDataType_t / ncclRedOp_t
  if ((rc = nccl.AllReduce(counters, counters, (size_t)n, kNcclUint64, kNcclSum, comm, (cudaStream_t)stream)))
    return nccl_fail("ncclAllReduce", rc);
  return DRG_OK;
}

int drg_comm_destroy(void* comm) {
  if (!comm) return DRG_OK;
  int rc = nccl_bind();
  if (rc) return rc;
  if ((rc = nccl.CommDestroy(comm))) return nccl_fail("ncclCommDestroy", rc);
  return DRG_OK;
}

int drg_selfplay_compact(int64_t n, const DrgGames* src, const DrgGames* dst, const DrgGames* fin, int64_t* n_running,
                         void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (n < 0 || !src || !fin || (dst && !n_running)) return fail(DRG_E_ARG, "bad arguments");
  if (n_running) *n_running = 0;
  if (n == 0) return DRG_OK;
  auto whole = [](const DrgGames* s, bool working) {
    return s->wm && s->wk && s->bm && s->bk && s->turn && s->clock && s->plies && s->state &&
           (!working || (s->hist && s->ids && s->game_ids));
  };
  if (!whole(src, true) || !whole(fin, false) || (dst && !whole(dst, true))) return fail(DRG_E_ARG, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  auto gs = [](const DrgGames* s) { return GameSet{s->wm, s->wk, s->bm, s->bk, s->turn, s->clock, s->plies, s->state, s->hist, s->ids, s->game_ids}; };
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (!dst) {
    k_compact_games<<<grid, 256, 0, st>>>(n, gs(src), gs(fin), gs(fin), false, nullptr);
    return launch_check("k_compact_games");
  }
  std::lock_guard<std::mutex> lk(g.mu);  // g.scan_tiles, g.h_word
  if ((rc = g.ckeep.ensure((size_t)n * sizeof(uint32_t)))) return rc;
  if ((rc = g.cpos.ensure((size_t)(n + 1) * sizeof(uint64_t)))) return rc;
  k_compact_flags<<<grid, 256, 0, st>>>(n, src->state, g.ckeep.as<uint32_t>());
  if ((rc = launch_check("k_compact_flags"))) return rc;
  if ((rc = scan_counts(n, g.ckeep.as<uint32_t>(), g.cpos.as<uint64_t>(), st))) return rc;
  k_compact_games<<<grid, 256, 0, st>>>(n, gs(src), gs(dst), gs(fin), true, g.cpos.as<uint64_t>());
  if ((rc = launch_check("k_compact_games"))) return rc;
  CUDA_TRY(cudaMemcpyAsync(g.h_word + 4, g.cpos.as<uint64_t>() + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *n_running = (int64_t)g.h_word[4];
  return DRG_OK;
}

int drg_sample_action(int variant, int64_t n, const float* logits, const uint32_t* counts, const uint64_t* offsets,
                      const DrgMove* moves, int64_t moves_cap, float temperature, int mode, uint64_t seed, uint64_t game_id0,
                      const uint64_t* game_ids, const uint16_t* plies, int32_t* action, int32_t* choice, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  const int S = drg_squares(variant);
  if (S < 0) return S;
  if (n < 0 || moves_cap < 0 || (mode != 0 && mode != 1) || !(temperature > 0.0f) ||
      (n > 0 && (!logits || !counts || !offsets || !moves || !action || !choice)))
    return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return DRG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + kWarps - 1) / kWarps);
  if (S == 50)
    k_sample_action<50><<<grid, kBlock, 0, st>>>(n, logits, counts, offsets, moves, (uint64_t)moves_cap, 1.0f / temperature, mode,
                                                 seed, game_id0, game_ids, plies, action, choice);
  else
    k_sample_action<32><<<grid, kBlock, 0, st>>>(n, logits, counts, offsets, moves, (uint64_t)moves_cap, 1.0f / temperature, mode,
                                                 seed, game_id0, game_ids, plies, action, choice);
  return launch_check("k_sample_action");
}

int drg_action_to_choice(int variant, int64_t n, const int32_t* action, const uint32_t* counts, const uint64_t* offsets,
                         const DrgMove* moves, int64_t moves_cap, int32_t* choice, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  const int S = drg_squares(variant);
  if (S < 0) return S;
  if (n < 0 || moves_cap < 0 || (n > 0 && (!action || !counts || !offsets || !moves || !choice))) return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return DRG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 50) k_action_to_choice<50><<<grid_for(n), kBlock, 0, st>>>(n, action, counts, offsets, moves, (uint64_t)moves_cap, choice);
  else k_action_to_choice<32><<<grid_for(n), kBlock, 0, st>>>(n, action, counts, offsets, moves, (uint64_t)moves_cap, choice);
  return launch_check("k_action_to_choice");
}

// ---- perft driver --------------------------------------------------------------------------------
// Level-synchronous frontier expansion in HBM (4 x uint64 per board, the side to move is uniform per
// level), chunked depth-first so that no level ever holds more than kLevelCap boards; the last ply is
// bulk-counted (k_perft_count), so a counted node costs ~64 B / branching-factor of HBM traffic.
namespace {

constexpr int64_t kLevelCap = 1 << 24;   // boards per frontier buffer (512 MiB per level)
constexpr int64_t kExpandGuess = 20;     // assumed upper bound of the mean branching factor of a chunk

struct PerftRun {
  int variant;
  cudaStream_t st;
  unsigned long long* ctr;  // device counters
  uint64_t movegen = 0;
  int rc = DRG_OK;
  // several GPUs: this rank keeps its share (by content hash) of the frontier at ONE ply, the first one that is wide
  // enough for an even deal (or the last but one); every rank expands the plies above it identically
  uint32_t rank = 0, world = 1;
  int shard_level = -1;  // level of the frontier that is dealt; -1: not decided yet
  int toggle = 0;                          // leaf-ply slot of the next chunk (PerftCountFn)
  bool deep_pending[2] = {false, false};   // a tree-search kernel of that slot is queued on the side stream
  bool no_overlap = false;                 // DRG_PERFT_NO_OVERLAP (measurement aid)
  std::vector<double> branching;           // per level: children per parent to size the chunks with
  std::vector<char> branching_seen;
};
constexpr int64_t kShardBoardsPerRank = 16384;

int perft_rec(PerftRun& R, int level, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm, const uint64_t* bk,
              int64_t n, int turn, int remaining, int leaf_slot = -1) {
  if (n == 0) return DRG_OK;
  if (remaining == 1) {
    // leaf_slot >= 0: the boards live in the leaf frontier buffer of that slot, which the caller will not touch before
    // the slot's event: their tree-search kernel may run beside what follows
    const bool overlap = leaf_slot >= 0 && !R.no_overlap;
    const int slot = overlap ? leaf_slot : 0;
    if (R.deep_pending[slot]) {  // the slot's list is still being read
      CUDA_TRY(cudaStreamWaitEvent(R.st, g.pf_deep[slot], 0));
      R.deep_pending[slot] = false;
    }
    PerftCountFn f{n, wm, wk, bm, bk, turn, R.ctr + K_TOTAL, R.ctr + K_OVF, R.st, slot, overlap};
    R.movegen += (uint64_t)n;
    const int rc = dispatch(R.variant, f);
    if (rc == DRG_OK && overlap) R.deep_pending[slot] = true;
    return rc;
  }
  if ((int)g.levels.size() <= level) g.levels.resize(level + 1);
  int rc = g.levels[level].ensure((size_t)kLevelCap * 4 * sizeof(uint64_t));
  if (rc) return rc;
  // remaining == 2: the children are the leaf ply; consecutive chunks alternate between two frontier buffers so that
  // the tree-search kernel of one chunk's leaves can still run while the next chunk is expanded (PerftCountFn)
  const bool leaves_next = remaining == 2 && !R.no_overlap;
  if (leaves_next && (rc = g.level_alt.ensure((size_t)kLevelCap * 4 * sizeof(uint64_t)))) return rc;
  // chunk size: parents whose children fill ~3/4 of a frontier buffer, from the largest branching factor this level
  // has shown so far (kExpandGuess before any chunk of the level has run); a chunk that overflows is redone in halves
  if ((int)R.branching.size() <= level) {
    R.branching.resize(level + 1, (double)kExpandGuess);
    R.branching_seen.resize(level + 1, 0);
  }
  int64_t chunk = (int64_t)(0.75 * (double)kLevelCap / R.branching[level]);
  if (chunk < 1) chunk = 1;
  int64_t start = 0;
  while (start < n) {
    const int64_t m = (n - start) < chunk ? (n - start) : chunk;
    const int slot = leaves_next ? R.toggle : -1;
    uint64_t* base = slot == 1 ? g.level_alt.as<uint64_t>() : g.levels[level].as<uint64_t>();
    uint64_t *owm = base, *owk = base + kLevelCap, *obm = base + 2 * kLevelCap, *obk = base + 3 * kLevelCap;
    if (slot >= 0 && R.deep_pending[slot]) {  // the buffer's previous leaves are still being walked
      CUDA_TRY(cudaStreamWaitEvent(R.st, g.pf_deep[slot], 0));
      R.deep_pending[slot] = false;
    }
    CUDA_TRY(cudaMemsetAsync(R.ctr + K_CURSOR, 0, sizeof(unsigned long long), R.st));
    ExpandFn f{m, wm + start, wk + start, bm + start, bk + start, turn, owm, owk, obm, obk, (uint64_t)kLevelCap,
               R.ctr + K_CURSOR, R.ctr + K_OVF, R.ctr + K_ILLEGAL, R.st};
    if ((rc = dispatch(R.variant, f))) return rc;
    CUDA_TRY(cudaMemcpyAsync(g.h_word, R.ctr + K_CURSOR, sizeof *g.h_word, cudaMemcpyDeviceToHost, R.st));
    CUDA_TRY(cudaStreamSynchronize(R.st));
    const unsigned long long produced = g.h_word[0];
    if ((int64_t)produced > kLevelCap) {  // did not fit: redo this chunk in smaller pieces
      if (m == 1) return fail(DRG_E_OVERFLOW, "a single board has more than %lld children", (long long)kLevelCap);
      chunk = m / 2 > 0 ? m / 2 : 1;
      const double b = (double)produced / (double)m * 1.15 + 0.5;
      if (b > R.branching[level]) R.branching[level] = b;
      R.branching_seen[level] = 1;
      continue;
    }
    R.movegen += (uint64_t)m;
    if (m >= 4096) {  // a chunk large enough for its mean to mean something
      const double b = (double)produced / (double)m * 1.15 + 0.5;
      if (!R.branching_seen[level] || b > R.branching[level]) R.branching[level] = b;
      R.branching_seen[level] = 1;
      chunk = (int64_t)(0.75 * (double)kLevelCap / R.branching[level]);
      if (chunk < 1) chunk = 1;
    }
    const uint64_t *cwm = owm, *cwk = owk, *cbm = obm, *cbk = obk;
    int64_t kept = (int64_t)produced;
    if (R.world > 1) {
      if (R.shard_level < 0 && ((int64_t)produced >= kShardBoardsPerRank * (int64_t)R.world || remaining - 1 <= 2))
        R.shard_level = level + 1;  // produced counts are the same on every rank, so is this decision
      if (R.shard_level == level + 1 && produced) {
        if ((rc = g.shard.ensure((size_t)produced * 4 * sizeof(uint64_t)))) return rc;
        uint64_t* sb = g.shard.as<uint64_t>();
        CUDA_TRY(cudaMemsetAsync(R.ctr + K_CURSOR, 0, sizeof(unsigned long long), R.st));
        k_perft_shard<<<(unsigned)((produced + 255) / 256), 256, 0, R.st>>>((int64_t)produced, owm, owk, obm, obk, R.rank, R.world,
                                                                             sb, sb + produced, sb + 2 * produced, sb + 3 * produced,
                                                                             R.ctr + K_CURSOR);
        if ((rc = launch_check("k_perft_shard"))) return rc;
        CUDA_TRY(cudaMemcpyAsync(g.h_word, R.ctr + K_CURSOR, sizeof *g.h_word, cudaMemcpyDeviceToHost, R.st));
        CUDA_TRY(cudaStreamSynchronize(R.st));
        kept = (int64_t)g.h_word[0];
        cwm = sb; cwk = sb + produced; cbm = sb + 2 * produced; cbk = sb + 3 * produced;
      }
    }
    // leaves dealt into the shard buffer (one buffer) are walked in sequence
    const int child_slot = slot >= 0 && cwm == owm ? slot : -1;
    if ((rc = perft_rec(R, level + 1, cwm, cwk, cbm, cbk, kept, turn ^ 1, remaining - 1, child_slot))) return rc;
    if (child_slot >= 0) R.toggle ^= 1;
    start += m;
  }
  return DRG_OK;
}

}  // namespace

namespace {
// drg_perft / drg_perft_sharded with the context lock held by the caller
int perft_locked(int variant, int64_t n_roots, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                 const uint64_t* bk, int turn, int depth, int rank, int world, uint64_t* nodes, uint64_t* counters,
                 uint64_t* counters_dev, cudaStream_t st) {
  int rc;
  unsigned long long* ctr = g.ctr.as<unsigned long long>();
  CUDA_TRY(cudaMemsetAsync(ctr, 0, K_NCTR * sizeof(unsigned long long), st));
  PerftRun R{variant, st, ctr};
  R.rank = (uint32_t)rank;
  R.world = (uint32_t)world;
  R.no_overlap = getenv("DRG_PERFT_NO_OVERLAP") != nullptr;
  unsigned long long h[K_NCTR] = {0};
  // the tree-search kernels still queued on the side stream are joined before the counters are read
  auto join = [&R, st]() {
    for (int s2 = 0; s2 < 2; s2++)
      if (R.deep_pending[s2]) {
        cudaStreamWaitEvent(st, g.pf_deep[s2], 0);
        R.deep_pending[s2] = false;
      }
  };
  if (depth == 0 || (world > 1 && depth == 1)) {
    // nothing to deal below the roots: rank 0 counts them
    if (rank == 0) {
      if (depth == 0) h[K_TOTAL] = (unsigned long long)n_roots;
      else {
        R.world = 1;
        rc = perft_rec(R, 0, wm, wk, bm, bk, n_roots, turn & 1, depth);
        join();
        if (rc) return rc;
      }
    }
  } else {
    rc = perft_rec(R, 0, wm, wk, bm, bk, n_roots, turn & 1, depth);
    join();
    if (rc) return rc;
  }
  if (depth > 0) {
    if (counters_dev) {
      k_perft_counters<<<1, 32, 0, st>>>(ctr, K_TOTAL, K_OVF, K_ILLEGAL, (unsigned long long)R.movegen,
                                         reinterpret_cast<unsigned long long*>(counters_dev));
      if ((rc = launch_check("k_perft_counters"))) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(h, ctr, sizeof h, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  } else if (counters_dev) {
    unsigned long long v[DRG_N_COUNTERS] = {0};
    v[DRG_C_NODES] = h[K_TOTAL];
    CUDA_TRY(cudaMemcpyAsync(counters_dev, v, sizeof v, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  *nodes = h[K_TOTAL];
  if (counters) {
    memset(counters, 0, DRG_N_COUNTERS * sizeof(uint64_t));
    counters[DRG_C_NODES] = h[K_TOTAL];
    counters[DRG_C_MOVEGEN] = R.movegen;
    counters[DRG_C_OVERFLOW] = h[K_OVF];
    counters[DRG_C_ILLEGAL] = h[K_ILLEGAL];
  }
  if (h[K_OVF]) return fail(DRG_E_OVERFLOW, "%llu boards had a capture path longer than %d squares", h[K_OVF], DRG_MAX_PATH);
  return DRG_OK;
}
}  // namespace

int drg_perft_sharded(int variant, int64_t n_roots, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                      const uint64_t* bk, int turn, int depth, int rank, int world, uint64_t* nodes, uint64_t* counters,
                      uint64_t* counters_dev, void* stream) {
  int rc = need_init();
  if (rc) return rc;
  if (family_of(variant) == F_BAD) return fail(DRG_E_ARG, "unknown variant id %d", variant);
  if (n_roots < 0 || depth < 0 || !nodes || (n_roots > 0 && (!wm || !wk || !bm || !bk))) return fail(DRG_E_ARG, "bad arguments");
  if (world < 1 || rank < 0 || rank >= world) return fail(DRG_E_ARG, "rank %d of %d", rank, world);
  std::lock_guard<std::mutex> lk(g.mu);
  return perft_locked(variant, n_roots, wm, wk, bm, bk, turn, depth, rank, world, nodes, counters, counters_dev,
                      (cudaStream_t)stream);
}

int drg_perft(int variant, int64_t n_roots, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
              const uint64_t* bk, int turn, int depth, uint64_t* nodes, uint64_t* counters, void* stream) {
  return drg_perft_sharded(variant, n_roots, wm, wk, bm, bk, turn, depth, 0, 1, nodes, counters, nullptr, stream);
}

// ---- host-buffer entry points -------------------------------------------------------------------------

namespace {

// upload SoA boards into g.in: [wm | wk | bm | bk] uint64[n] then turn uint8[n] then clock uint8[n]
struct DevBoards {
  uint64_t *wm, *wk, *bm, *bk;
  uint8_t *turn, *clock;
};

int upload_boards(int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm, const uint64_t* bk,
                  const uint8_t* turn, const uint8_t* clock, DevBoards& d, cudaStream_t st) {
  const size_t nn = (size_t)(n ? n : 1);
  int rc = g.in.ensure(nn * 34 + 64);
  if (rc) return rc;
  uint64_t* b = g.in.as<uint64_t>();
  d.wm = b; d.wk = b + nn; d.bm = b + 2 * nn; d.bk = b + 3 * nn;
  d.turn = reinterpret_cast<uint8_t*>(b + 4 * nn);
  d.clock = d.turn + nn;
  if (n == 0) return DRG_OK;
  CUDA_TRY(cudaMemcpyAsync(d.wm, wm, n * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d.wk, wk, n * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d.bm, bm, n * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d.bk, bk, n * 8, cudaMemcpyHostToDevice, st));
  if (turn) CUDA_TRY(cudaMemcpyAsync(d.turn, turn, n, cudaMemcpyHostToDevice, st));
  else CUDA_TRY(cudaMemsetAsync(d.turn, 0, n, st));
  if (clock) CUDA_TRY(cudaMemcpyAsync(d.clock, clock, n, cudaMemcpyHostToDevice, st));
  else CUDA_TRY(cudaMemsetAsync(d.clock, 0, n, st));
  return DRG_OK;
}

int download_boards(int64_t n, uint64_t* wm, uint64_t* wk, uint64_t* bm, uint64_t* bk, uint8_t* turn, uint8_t* clock,
                    const DevBoards& d, cudaStream_t st) {
  if (n == 0) return DRG_OK;
  CUDA_TRY(cudaMemcpyAsync(wm, d.wm, n * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(wk, d.wk, n * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(bm, d.bm, n * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(bk, d.bk, n * 8, cudaMemcpyDeviceToHost, st));
  if (turn) CUDA_TRY(cudaMemcpyAsync(turn, d.turn, n, cudaMemcpyDeviceToHost, st));
  if (clock) CUDA_TRY(cudaMemcpyAsync(clock, d.clock, n, cudaMemcpyDeviceToHost, st));
  return DRG_OK;
}

// ---- small batches: one launch, no copies (k_small on mapped pinned memory) -----------------------------
constexpr int64_t kSmallN = kBlock;
constexpr int kFallThrough = 1;  // result does not fit the small block: use the batched pipeline

inline size_t pad16(size_t b) { return (b + 15) & ~(size_t)15; }

int small_path(int variant, int S, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
               const uint64_t* bk, const uint8_t* turn, uint32_t* counts, uint64_t* offsets, DrgMove* moves,
               int64_t moves_cap, int64_t* total, uint8_t* mask, float* tensor) {
  cudaStream_t st = 0;
  const bool emit = moves != nullptr;
  // result block: info | counts | offsets | mask | tensor | records
  const size_t o_counts = 16, o_off = o_counts + pad16(n * 4), o_mask = o_off + pad16((n + 1) * 8);
  const size_t o_tensor = o_mask + (emit && mask ? pad16((size_t)n * S * S) : 0);
  const size_t o_moves = o_tensor + (emit && tensor ? pad16((size_t)n * 4 * S * sizeof(float)) : 0);
  const size_t cap_small = emit ? (size_t)(64 * n + 448) : 0;
  const size_t out_bytes = o_moves + cap_small * sizeof(DrgMove);
  const size_t in_bytes = pad16((size_t)n * 33);
  int rc;
  if (g.h_small_cap < in_bytes + out_bytes) {
    if (g.h_small) cudaFreeHost(g.h_small);
    g.h_small = nullptr;
    g.h_small_cap = 0;
    const size_t want = in_bytes + out_bytes + (256 << 10);
    CUDA_TRY(cudaMallocHost(&g.h_small, want));
    g.h_small_cap = want;
  }
  uint8_t* h_in = g.h_small;
  uint8_t* h_out = g.h_small + in_bytes;
  memcpy(h_in, wm, n * 8);
  memcpy(h_in + n * 8, wk, n * 8);
  memcpy(h_in + n * 16, bm, n * 8);
  memcpy(h_in + n * 24, bk, n * 8);
  memcpy(h_in + n * 32, turn, n);
  void* d_in = nullptr;
  CUDA_TRY(cudaHostGetDevicePointer(&d_in, h_in, 0));  // pinned memory is mapped: the kernel reads it in place
  // results are written in place too (posted writes over the link; stream completion makes them visible): the
  // staged variant -- device block + one copy -- cost 46 us per call against 27 us
  uint8_t* d_out = static_cast<uint8_t*>(d_in) + in_bytes;
  SmallArgs a;
  a.n = (int)n;
  a.emit = emit;
  const uint64_t* in64 = static_cast<const uint64_t*>(d_in);
  a.wm = in64; a.wk = in64 + n; a.bm = in64 + 2 * n; a.bk = in64 + 3 * n;
  a.turn = reinterpret_cast<const uint8_t*>(in64 + 4 * n);
  a.info = reinterpret_cast<unsigned long long*>(d_out);
  a.counts = reinterpret_cast<uint32_t*>(d_out + o_counts);
  a.offsets = reinterpret_cast<uint64_t*>(d_out + o_off);
  a.mask = emit && mask ? d_out + o_mask : nullptr;
  a.tensor = emit && tensor ? reinterpret_cast<float*>(d_out + o_tensor) : nullptr;
  a.moves = reinterpret_cast<DrgMove*>(d_out + o_moves);
  a.moves_cap = (uint32_t)cap_small;
  SmallFn f{a, st};
  if ((rc = dispatch(variant, f))) return rc;
  CUDA_TRY(cudaStreamSynchronize(st));
  const unsigned long long* info = reinterpret_cast<const unsigned long long*>(h_out);
  const uint64_t tot = info[1];
  if (total) *total = (int64_t)tot;
  memcpy(counts, h_out + o_counts, n * 4);
  if (offsets) memcpy(offsets, h_out + o_off, (n + 1) * 8);
  if (!emit) return DRG_OK;
  if ((int64_t)tot > moves_cap) return fail(DRG_E_OVERFLOW, "moves buffer holds %lld records, %llu needed", (long long)moves_cap, (unsigned long long)tot);
  if (tot > cap_small) return kFallThrough;
  memcpy(moves, h_out + o_moves, tot * sizeof(DrgMove));
  if (mask) memcpy(mask, h_out + o_mask, (size_t)n * S * S);
  if (tensor) memcpy(tensor, h_out + o_tensor, (size_t)n * 4 * S * sizeof(float));
  if (info[0]) return fail(DRG_E_OVERFLOW, "%llu boards had a capture path longer than %d squares", info[0], DRG_MAX_PATH);
  return DRG_OK;
}

}  // namespace

namespace {
// mask: one byte per index (the reference's layout) or NULL; mask_bits: the same rows as a flat little-endian bit
// stream (bit j = byte j of the [n, S*S] array) or NULL -- at most one of the two
int legal_moves_host_impl(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                          const uint64_t* bk, const uint8_t* turn, uint32_t* counts, uint64_t* offsets, DrgMove* moves,
                          int64_t moves_cap, int64_t* total, uint8_t* mask, uint8_t* mask_bits, float* tensor) {
  int rc = need_init();
  if (rc) return rc;
  const int S = drg_squares(variant);
  if (S < 0) return S;
  if (n < 0 || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !counts))) return fail(DRG_E_ARG, "bad arguments");
  if ((mask || mask_bits || tensor) && !moves) return fail(DRG_E_ARG, "mask / tensor export needs a moves buffer");
  if (mask && mask_bits) return fail(DRG_E_ARG, "ask for the mask rows as bytes or as bits, not both");
  const bool want_mask = mask || mask_bits;
  std::lock_guard<std::mutex> lk(g.mu);
  cudaStream_t st = 0;
  if (total) *total = 0;
  if (n == 0) {
    if (offsets) offsets[0] = 0;
    return DRG_OK;
  }
  if (n <= kSmallN && !mask_bits && !getenv("DRG_NO_SMALL_PATH")) {
    rc = small_path(variant, S, n, wm, wk, bm, bk, turn, counts, offsets, moves, moves_cap, total, mask, tensor);
    if (rc != kFallThrough) return rc;
  }
  const bool dbg = getenv("DRG_DEBUG_TIMING") != nullptr;
  const auto T0 = std::chrono::steady_clock::now();
  auto stamp = [&](const char* what) {
    if (dbg) fprintf(stderr, "[drg] %-22s +%.2f ms\n", what,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count());
  };
  DrainOnExit drain{st};
  DevBoards d;
  if ((rc = upload_boards(n, wm, wk, bm, bk, turn, nullptr, d, st))) return rc;
  if ((rc = g.counts.ensure(n * sizeof(uint32_t)))) return rc;
  if ((rc = g.offsets.ensure((n + 1) * sizeof(uint64_t)))) return rc;
  uint32_t* d_counts = g.counts.as<uint32_t>();
  uint64_t* d_off = g.offsets.as<uint64_t>();
  unsigned long long* ctr = g.ctr.as<unsigned long long>();
  CUDA_TRY(cudaMemsetAsync(ctr + K_OVF, 0, sizeof(unsigned long long), st));
  // With a record buffer of known capacity nothing has to wait for the host: the fused device call (mask rows from
  // t = 0, counts -> offsets -> records beside them) runs into device buffers of that capacity, the packed rows start
  // crossing the link chunk by chunk right behind it, and the host only waits for the 8-byte total before it
  // queues the record copy of the exact size.  (The two-step form below reads the total back BEFORE the emit.)
  if (moves && moves_cap > 0 && (size_t)moves_cap * sizeof(DrgMove) <= ((size_t)8 << 30) && !getenv("DRG_HOST_TWO_STEP")) {
    if ((rc = g.moves.ensure((size_t)moves_cap * sizeof(DrgMove)))) return rc;
    uint8_t* d_mask = nullptr;
    float* d_tensor = nullptr;
    if (want_mask) { if ((rc = g.mask.ensure((size_t)n * S * S))) return rc; d_mask = g.mask.as<uint8_t>(); }
    if (tensor) { if ((rc = g.tensor.ensure((size_t)n * 4 * S * sizeof(float)))) return rc; d_tensor = g.tensor.as<float>(); }
    if ((rc = drg_movegen(variant, n, d.wm, d.wk, d.bm, d.bk, d.turn, d_counts, d_off, g.moves.as<DrgMove>(), moves_cap,
                          d_mask, d_tensor, reinterpret_cast<uint64_t*>(ctr + K_OVF), st)))
      return rc;
    CUDA_TRY(cudaMemcpyAsync(g.h_word, d_off + n, sizeof *g.h_word, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(g.h_word + 1, ctr + K_OVF, sizeof *g.h_word, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(g.ev_total, st));
    const size_t mask_bytes = want_mask ? (size_t)n * S * S : 0;
    const bool packed = mask_bits || (mask_bytes >= ((size_t)1 << 20) && !getenv("DRG_NO_PACKED_MASK"));
    int nchunks = 0;
    size_t chunk = 0;
    if (packed) {
      const size_t nwords = (mask_bytes + 31) / 32;
      if ((rc = g.maskbits.ensure(nwords * 4))) return rc;
      if (!mask_bits && g.h_bits_cap < nwords * 4) {
        if (g.h_bits) cudaFreeHost(g.h_bits);
        g.h_bits = nullptr;
        g.h_bits_cap = 0;
        CUDA_TRY(cudaMallocHost(&g.h_bits, nwords * 4 + 4096));
        g.h_bits_cap = nwords * 4 + 4096;
      }
      nchunks = (int)(mask_bytes / ((size_t)64 << 20)) + 1;  // ~64 MB of mask (8 MB on the wire) per chunk
      if (nchunks > kMaxChunks) nchunks = kMaxChunks;
      chunk = ((mask_bytes / nchunks) + 2047) & ~(size_t)2047;
      nchunks = (int)((mask_bytes + chunk - 1) / chunk);
      for (int c = 0; c < nchunks; c++) {  // pack + copy chunk by chunk: the first one lands while the rest is packed
        if (!g.chunk_ev[c]) CUDA_TRY(cudaEventCreateWithFlags(&g.chunk_ev[c], cudaEventDisableTiming));
        const size_t b_lo = (size_t)c * chunk, b_hi = b_lo + chunk < mask_bytes ? b_lo + chunk : mask_bytes;
        const size_t w = (b_hi - b_lo + 31) / 32;
        k_pack_mask<<<(unsigned)((w + 255) / 256), 256, 0, st>>>((int64_t)(b_hi - b_lo), d_mask + b_lo,
                                                                 g.maskbits.as<uint32_t>() + b_lo / 32);
        if ((rc = launch_check("k_pack_mask"))) return rc;
        // packed rows asked for: they land in the caller's buffer as they are (the last word may carry padding bits)
        const size_t wire = mask_bits ? (b_hi - b_lo + 7) / 8 : w * 4;
        CUDA_TRY(cudaMemcpyAsync((mask_bits ? mask_bits : g.h_bits) + b_lo / 8, g.maskbits.as<uint8_t>() + b_lo / 8, wire,
                                 cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaEventRecord(g.chunk_ev[c], st));
      }
    } else if (mask) {
      CUDA_TRY(cudaMemcpyAsync(mask, d_mask, mask_bytes, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaMemcpyAsync(counts, d_counts, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (offsets) CUDA_TRY(cudaMemcpyAsync(offsets, d_off, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (tensor) CUDA_TRY(cudaMemcpyAsync(tensor, d_tensor, (size_t)n * 4 * S * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventSynchronize(g.ev_total));
    const uint64_t tot = g.h_word[0];
    const unsigned long long ovf = g.h_word[1];
    stamp("total on host");
    if (total) *total = (int64_t)tot;
    if ((int64_t)tot > moves_cap) {
      CUDA_TRY(cudaStreamSynchronize(st));
      return fail(DRG_E_OVERFLOW, "moves buffer holds %lld records, %llu needed", (long long)moves_cap, (unsigned long long)tot);
    }
    if (tot) CUDA_TRY(cudaMemcpyAsync(moves, g.moves.p, tot * sizeof(DrgMove), cudaMemcpyDeviceToHost, st));
    if (packed && !mask_bits) {
      struct Waiter {
        static void wait(void* ctx, int c) { cudaEventSynchronize(static_cast<cudaEvent_t*>(ctx)[c]); }
      };
      drg_host::expand_pipelined(g.h_bits, mask, mask_bytes, chunk, nchunks, &Waiter::wait, g.chunk_ev,
                                 drg_host::default_threads());
    }
    stamp("mask rows widened");
    CUDA_TRY(cudaStreamSynchronize(st));
    stamp("all copies landed");
    if (ovf) return fail(DRG_E_OVERFLOW, "%llu boards had a capture path longer than %d squares", ovf, DRG_MAX_PATH);
    return DRG_OK;
  }
  CountFn cf{n, d.wm, d.wk, d.bm, d.bk, d.turn, d_counts, ctr + K_OVF, st};
  if ((rc = dispatch(variant, cf))) return rc;
  if ((rc = scan_counts(n, d_counts, d_off, st))) return rc;
  CUDA_TRY(cudaMemcpyAsync(counts, d_counts, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (offsets) CUDA_TRY(cudaMemcpyAsync(offsets, d_off, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(g.h_word, d_off + n, sizeof *g.h_word, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const uint64_t tot = g.h_word[0];
  stamp("counts on host");
  if (total) *total = (int64_t)tot;
  if (!moves) return DRG_OK;
  if ((int64_t)tot > moves_cap) return fail(DRG_E_OVERFLOW, "moves buffer holds %lld records, %llu needed", (long long)moves_cap, (unsigned long long)tot);
  if ((rc = g.moves.ensure((size_t)(tot ? tot : 1) * sizeof(DrgMove)))) return rc;
  uint8_t* d_mask = nullptr;
  float* d_tensor = nullptr;
  if (want_mask) { if ((rc = g.mask.ensure((size_t)n * S * S))) return rc; d_mask = g.mask.as<uint8_t>(); }
  if (tensor) { if ((rc = g.tensor.ensure((size_t)n * 4 * S * sizeof(float)))) return rc; d_tensor = g.tensor.as<float>(); }
  EmitFn ef{n, d.wm, d.wk, d.bm, d.bk, d.turn, d_off, g.moves.as<DrgMove>(), d_mask, d_tensor, nullptr, ctr + K_OVF, st,
            ~0ull, true};
  if ((rc = dispatch(variant, ef))) return rc;
  // Mask rows cross PCIe bit-packed (8:1) when they are large: k_pack_mask on the device, chunked D2H into a
  // pinned staging buffer, widened back to one byte per index by host threads while later chunks (and the
  // move records) are still in flight (drg_host.cpp).  Small masks are copied as they are.
  const size_t mask_bytes = want_mask ? (size_t)n * S * S : 0;
  const bool packed = mask_bits || (mask_bytes >= ((size_t)1 << 20) && !getenv("DRG_NO_PACKED_MASK"));
  int nchunks = 0;
  size_t chunk = 0;
  if (packed) {
    const size_t nwords = (mask_bytes + 31) / 32;
    if ((rc = g.maskbits.ensure(nwords * 4))) return rc;
    if (!mask_bits && g.h_bits_cap < nwords * 4) {
      if (g.h_bits) cudaFreeHost(g.h_bits);
      g.h_bits = nullptr;
      g.h_bits_cap = 0;
      CUDA_TRY(cudaMallocHost(&g.h_bits, nwords * 4 + 4096));
      g.h_bits_cap = nwords * 4 + 4096;
    }
    k_pack_mask<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>((int64_t)mask_bytes, d_mask, g.maskbits.as<uint32_t>());
    if ((rc = launch_check("k_pack_mask"))) return rc;
    nchunks = (int)(mask_bytes / ((size_t)64 << 20)) + 1;  // ~64 MB of mask (8 MB on the wire) per chunk
    if (nchunks > kMaxChunks) nchunks = kMaxChunks;
    chunk = ((mask_bytes / nchunks) + 2047) & ~(size_t)2047;
    nchunks = (int)((mask_bytes + chunk - 1) / chunk);
    for (int c = 0; c < nchunks; c++) {
      if (!g.chunk_ev[c]) CUDA_TRY(cudaEventCreateWithFlags(&g.chunk_ev[c], cudaEventDisableTiming));
      const size_t lo = (size_t)c * chunk / 8;
      const size_t hi = ((size_t)(c + 1) * chunk < mask_bytes ? (size_t)(c + 1) * chunk / 8
                                                             : (mask_bits ? (mask_bytes + 7) / 8 : nwords * 4));
      CUDA_TRY(cudaMemcpyAsync((mask_bits ? mask_bits : g.h_bits) + lo, g.maskbits.as<uint8_t>() + lo, hi - lo, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaEventRecord(g.chunk_ev[c], st));
    }
  } else if (mask) {
    CUDA_TRY(cudaMemcpyAsync(mask, d_mask, mask_bytes, cudaMemcpyDeviceToHost, st));
  }
  if (tot) CUDA_TRY(cudaMemcpyAsync(moves, g.moves.p, tot * sizeof(DrgMove), cudaMemcpyDeviceToHost, st));
  if (tensor) CUDA_TRY(cudaMemcpyAsync(tensor, d_tensor, (size_t)n * 4 * S * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(g.h_word + 1, ctr + K_OVF, sizeof *g.h_word, cudaMemcpyDeviceToHost, st));
  stamp("emit + copies queued");
  if (packed && !mask_bits) {
    struct Waiter {
      static void wait(void* ctx, int c) { cudaEventSynchronize(static_cast<cudaEvent_t*>(ctx)[c]); }
    };
    const auto t0 = std::chrono::steady_clock::now();
    drg_host::expand_pipelined(g.h_bits, mask, mask_bytes, chunk, nchunks, &Waiter::wait, g.chunk_ev,
                               drg_host::default_threads());
    if (getenv("DRG_DEBUG_TIMING")) {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      fprintf(stderr, "[drg] packed mask: %zu bytes, %d chunks, %d threads, transfer+widen %.2f ms\n", mask_bytes,
              nchunks, drg_host::default_threads(), ms);
    }
  }
  stamp("mask rows widened");
  CUDA_TRY(cudaStreamSynchronize(st));
  stamp("all copies landed");
  const unsigned long long ovf = g.h_word[1];
  if (ovf) return fail(DRG_E_OVERFLOW, "%llu boards had a capture path longer than %d squares", ovf, DRG_MAX_PATH);
  return DRG_OK;
}

}  // namespace

int drg_legal_moves_host(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                         const uint64_t* bk, const uint8_t* turn, uint32_t* counts, uint64_t* offsets, DrgMove* moves,
                         int64_t moves_cap, int64_t* total, uint8_t* mask, float* tensor) {
  return legal_moves_host_impl(variant, n, wm, wk, bm, bk, turn, counts, offsets, moves, moves_cap, total, mask, nullptr, tensor);
}

int drg_legal_moves_host_bits(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                              const uint64_t* bk, const uint8_t* turn, uint32_t* counts, uint64_t* offsets, DrgMove* moves,
                              int64_t moves_cap, int64_t* total, uint8_t* mask_bits, float* tensor) {
  return legal_moves_host_impl(variant, n, wm, wk, bm, bk, turn, counts, offsets, moves, moves_cap, total, nullptr, mask_bits,
                               tensor);
}

int drg_evaluate_host(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                      const uint64_t* bk, const uint8_t* turn, double* score) {
  int rc = need_init();
  if (rc) return rc;
  if (family_of(variant) == F_BAD) return fail(DRG_E_ARG, "unknown variant id %d", variant);
  if (n < 0 || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !score))) return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return DRG_OK;
  std::lock_guard<std::mutex> lk(g.mu);
  cudaStream_t st = 0;
  DrainOnExit drain{st};
  DevBoards d;
  if ((rc = upload_boards(n, wm, wk, bm, bk, turn, nullptr, d, st))) return rc;
  if ((rc = g.misc.ensure((size_t)n * sizeof(double)))) return rc;
  if ((rc = drg_evaluate(variant, n, d.wm, d.wk, d.bm, d.bk, d.turn, g.misc.as<double>(), st))) return rc;
  CUDA_TRY(cudaMemcpyAsync(score, g.misc.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DRG_OK;
}

int drg_apply_host(int variant, int64_t n, uint64_t* wm, uint64_t* wk, uint64_t* bm, uint64_t* bk, uint8_t* turn,
                   uint8_t* clock, const DrgMove* chosen, uint8_t* status) {
  int rc = need_init();
  if (rc) return rc;
  if (family_of(variant) == F_BAD) return fail(DRG_E_ARG, "unknown variant id %d", variant);
  if (n < 0 || (n > 0 && (!wm || !wk || !bm || !bk || !turn || !clock || !chosen))) return fail(DRG_E_ARG, "bad arguments");
  if (n == 0) return DRG_OK;
  std::lock_guard<std::mutex> lk(g.mu);
  cudaStream_t st = 0;
  if (n <= kSmallN && !getenv("DRG_NO_SMALL_PATH")) {
    // Board.push is a batch of 1: k_apply works in place on mapped pinned words (one launch, no copies queued)
    const size_t o_chosen = pad16((size_t)n * 34), o_status = o_chosen + (size_t)n * sizeof(DrgMove);
    const size_t bytes = o_status + pad16((size_t)n);
    if (g.h_small_cap < bytes) {
      if (g.h_small) cudaFreeHost(g.h_small);
      g.h_small = nullptr;
      g.h_small_cap = 0;
      CUDA_TRY(cudaMallocHost(&g.h_small, bytes + (256 << 10)));
      g.h_small_cap = bytes + (256 << 10);
    }
    uint8_t* h = g.h_small;
    uint64_t* hw = reinterpret_cast<uint64_t*>(h);
    memcpy(hw, wm, n * 8); memcpy(hw + n, wk, n * 8); memcpy(hw + 2 * n, bm, n * 8); memcpy(hw + 3 * n, bk, n * 8);
    memcpy(h + n * 32, turn, n); memcpy(h + n * 33, clock, n);
    memcpy(h + o_chosen, chosen, (size_t)n * sizeof(DrgMove));
    void* dv = nullptr;
    CUDA_TRY(cudaHostGetDevicePointer(&dv, h, 0));
    uint8_t* dp = static_cast<uint8_t*>(dv);
    uint64_t* dw = reinterpret_cast<uint64_t*>(dp);
    ApplyFn f{n, dw, dw + n, dw + 2 * n, dw + 3 * n, dp + n * 32, dp + n * 33,
              reinterpret_cast<const DrgMove*>(dp + o_chosen), dp + o_status, st};
    if ((rc = dispatch(variant, f))) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(wm, hw, n * 8); memcpy(wk, hw + n, n * 8); memcpy(bm, hw + 2 * n, n * 8); memcpy(bk, hw + 3 * n, n * 8);
    memcpy(turn, h + n * 32, n); memcpy(clock, h + n * 33, n);
    if (status) memcpy(status, h + o_status, n);
    return DRG_OK;
  }
  DrainOnExit drain{st};
  DevBoards d;
  if ((rc = upload_boards(n, wm, wk, bm, bk, turn, clock, d, st))) return rc;
  if ((rc = g.moves.ensure((size_t)n * sizeof(DrgMove)))) return rc;
  if ((rc = g.misc.ensure((size_t)n))) return rc;
  CUDA_TRY(cudaMemcpyAsync(g.moves.p, chosen, (size_t)n * sizeof(DrgMove), cudaMemcpyHostToDevice, st));
  ApplyFn f{n, d.wm, d.wk, d.bm, d.bk, d.turn, d.clock, g.moves.as<DrgMove>(), g.misc.as<uint8_t>(), st};
  if ((rc = dispatch(variant, f))) return rc;
  if ((rc = download_boards(n, wm, wk, bm, bk, turn, clock, d, st))) return rc;
  if (status) CUDA_TRY(cudaMemcpyAsync(status, g.misc.p, (size_t)n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DRG_OK;
}

int drg_perft_host(int variant, int64_t n_roots, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                   const uint64_t* bk, int turn, int depth, uint64_t* nodes, uint64_t* counters) {
  int rc = need_init();
  if (rc) return rc;
  if (family_of(variant) == F_BAD) return fail(DRG_E_ARG, "unknown variant id %d", variant);
  if (n_roots < 0 || !nodes || (n_roots > 0 && (!wm || !wk || !bm || !bk))) return fail(DRG_E_ARG, "bad arguments");
  if (n_roots < 0 || depth < 0) return fail(DRG_E_ARG, "bad arguments");
  DevBoards d;
  std::lock_guard<std::mutex> lk(g.mu);  // one lock over the upload into the shared staging buffer AND the run that reads it
  if ((rc = upload_boards(n_roots, wm, wk, bm, bk, nullptr, nullptr, d, 0))) return rc;
  return perft_locked(variant, n_roots, d.wm, d.wk, d.bm, d.bk, turn, depth, 0, 1, nodes, counters, nullptr, nullptr);
}

int drg_rollout_host(int variant, int64_t n_games, uint64_t seed, uint64_t game_id0, int max_plies,
                     const uint16_t* stop_ply, uint32_t flags, uint64_t* wm, uint64_t* wk, uint64_t* bm, uint64_t* bk,
                     uint8_t* turn, uint8_t* clock, uint8_t* result, uint16_t* plies, uint64_t* counters) {
  int rc = need_init();
  if (rc) return rc;
  if (family_of(variant) == F_BAD) return fail(DRG_E_ARG, "unknown variant id %d", variant);
  if (n_games < 0 || (n_games > 0 && (!wm || !wk || !bm || !bk || !turn || !clock))) return fail(DRG_E_ARG, "bad arguments");
  if (n_games == 0) return DRG_OK;
  std::lock_guard<std::mutex> lk(g.mu);
  cudaStream_t st = 0;
  const int64_t n = n_games;
  DrainOnExit drain{st};
  DevBoards d;
  if ((rc = upload_boards(n, wm, wk, bm, bk, turn, clock, d, st))) return rc;
  // misc: result u8[n] | plies u16[n] | stop u16[n] | counters u64[8]
  const size_t o_res = 0, o_pl = ((size_t)n + 1) & ~(size_t)1, o_stop = o_pl + 2 * (size_t)n;
  const size_t o_ctr = (o_stop + 2 * (size_t)n + 7) & ~(size_t)7;
  if ((rc = g.misc.ensure(o_ctr + DRG_N_COUNTERS * 8))) return rc;
  uint8_t* m = g.misc.as<uint8_t>();
  uint16_t* d_stop = nullptr;
  if (stop_ply) {
    d_stop = reinterpret_cast<uint16_t*>(m + o_stop);
    CUDA_TRY(cudaMemcpyAsync(d_stop, stop_ply, 2 * (size_t)n, cudaMemcpyHostToDevice, st));
  }
  CUDA_TRY(cudaMemsetAsync(m + o_ctr, 0, DRG_N_COUNTERS * 8, st));
  rc = drg_rollout(variant, n, seed, game_id0, max_plies, d_stop, flags, d.wm, d.wk, d.bm, d.bk, d.turn, d.clock,
                   m + o_res, reinterpret_cast<uint16_t*>(m + o_pl), reinterpret_cast<uint64_t*>(m + o_ctr), st);
  if (rc) return rc;
  if ((rc = download_boards(n, wm, wk, bm, bk, turn, clock, d, st))) return rc;
  if (result) CUDA_TRY(cudaMemcpyAsync(result, m + o_res, (size_t)n, cudaMemcpyDeviceToHost, st));
  if (plies) CUDA_TRY(cudaMemcpyAsync(plies, m + o_pl, 2 * (size_t)n, cudaMemcpyDeviceToHost, st));
  uint64_t hc[DRG_N_COUNTERS];
  CUDA_TRY(cudaMemcpyAsync(hc, m + o_ctr, sizeof hc, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (counters) memcpy(counters, hc, sizeof hc);
  if (hc[DRG_C_OVERFLOW]) return fail(DRG_E_OVERFLOW, "%llu games met a capture path longer than %d squares", (unsigned long long)hc[DRG_C_OVERFLOW], DRG_MAX_PATH);
  return DRG_OK;
}

}  // extern "C"
