/*
 * drg.h -- C-ABI of the B200-native draughts hot path (libdrg_b200.so).
 *
 * The reference (miskibin/py-draughts v1.8.3) is pure Python and exposes no
 * FFI; the interface this library replaces is the Python class API of
 * draughts/boards/base.py (BaseBoard) and the per-variant generators.  Every
 * entry point below names the reference function(s) it stands in for
 * (file:line relative to the reference root).  The ctypes binding a maintainer
 * would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; no C++/torch types cross this boundary;
 *   - return value 0 = ok, <0 = error (DRG_E_*); drg_last_error() gives the
 *     thread-local message; no exception crosses the ABI;
 *   - "device" entry points take DEVICE pointers owned by the caller and are
 *     enqueued on `stream` (a cudaStream_t passed as void*, NULL = default
 *     stream); they return without synchronising unless stated;
 *   - "_host" entry points take HOST pointers, do H2D + kernels + D2H inside
 *     and return when the results are in the host buffers;
 *   - there is no CPU fallback: with no usable CUDA device every compute entry
 *     point fails with DRG_E_CUDA;
 *   - one context per process, bound to one device at a time (drg_init with
 *     another device releases everything held on the old one first); the
 *     device entry points share
 *     scratch buffers (tree-search lists, scan tiles) and one internal
 *     high-priority stream, so calls must be issued in ONE stream order: one
 *     host thread at a time, successive calls on the same stream (or on
 *     streams the caller has ordered).  drg_movegen with a mask forks onto the
 *     internal stream and joins before it returns control of `stream`: to the
 *     caller it is an ordinary stream-ordered call.  The _host entry points
 *     serialise themselves with a mutex.
 *
 * Position layout (structure of arrays, one element per board):
 *   wm, wk, bm, bk : uint64 bitboards white men / white kings / black men /
 *                    black kings; bit i = square i (0-based, square 0 = top
 *                    left dark square = black's back rank), base.py:53-54,
 *                    standard.py:14-18.
 *   turn           : uint8, 0 = WHITE to move, 1 = BLACK (models.py:12-14
 *                    uses -1/+1; the shim converts).
 *   clock          : uint8 halfmove_clock (base.py:277-280).
 */
#ifndef DRG_H_
#define DRG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRG_ABI_VERSION 3 /* 3: + drg_selfplay_compact, drg_nccl_unique_id / drg_comm_init / drg_allreduce_counters / drg_comm_destroy */

/* variant ids, order of test/_test_helpers.py:19-28 */
enum {
  DRG_STANDARD = 0,     /* draughts/boards/standard.py     */
  DRG_AMERICAN = 1,     /* draughts/boards/american.py     */
  DRG_FRISIAN = 2,      /* draughts/boards/frisian.py      */
  DRG_RUSSIAN = 3,      /* draughts/boards/russian.py      */
  DRG_BRAZILIAN = 4,    /* draughts/boards/brazilian.py    */
  DRG_ANTIDRAUGHTS = 5, /* draughts/boards/antidraughts.py */
  DRG_BREAKTHROUGH = 6, /* draughts/boards/breakthrough.py */
  DRG_FRYSK = 7,        /* draughts/boards/frysk.py        */
  DRG_N_VARIANTS = 8
};

enum {
  DRG_OK = 0,
  DRG_E_ARG = -1,      /* bad argument (variant id, null pointer, n < 0)    */
  DRG_E_CUDA = -2,     /* CUDA runtime error / no device                    */
  DRG_E_OVERFLOW = -3, /* an output buffer / path record was too small      */
  DRG_E_NOMEM = -4
};

/*
 * Packed fixed-width move record (replaces draughts/move.py:8-65 on device).
 * 32 bytes, 16-byte aligned, little endian.
 *   captured : bit mask of the victim squares (move.py captured_list as a set;
 *              the ORDER is recovered on the host: victim k lies between
 *              path[k] and path[k+1]).  0 for a quiet move.
 *   n_sq     : number of visited squares = len(square_list) (2 for a quiet
 *              move, captures + 1 for a capture).
 *   flags    : DRG_MF_*.
 *   path     : square_list, 0-based, path[0] = from, path[n_sq-1] = to.
 */
#define DRG_MAX_PATH 22
#define DRG_MF_PROMO 0x01u    /* Move.is_promotion as GENERATED (Russian mid-capture, russian.py:281-282) */
#define DRG_MF_KINGMOVE 0x02u /* mover was a king at the start of the move          */
#define DRG_MF_CAPTURE 0x04u  /* captured != 0                                       */
#define DRG_MF_TRUNC 0x80u    /* path longer than DRG_MAX_PATH: record truncated     */

typedef struct DrgMove {
  uint64_t captured;
  uint8_t n_sq;
  uint8_t flags;
  uint8_t path[DRG_MAX_PATH];
} DrgMove;

/* counters written by the perft / rollout drivers (uint64 each) */
enum {
  DRG_C_NODES = 0,     /* perft leaf count / rollout plies played            */
  DRG_C_MOVEGEN = 1,   /* positions on which legal moves were generated      */
  DRG_C_WHITE_WINS = 2,
  DRG_C_BLACK_WINS = 3,
  DRG_C_DRAWS = 4,
  DRG_C_UNFINISHED = 5, /* rollouts stopped by max_plies / stop_ply          */
  DRG_C_OVERFLOW = 6,   /* truncated records / capacity overflows (must be 0) */
  DRG_C_ILLEGAL = 7,    /* apply() rejected a move (base.py:241-245)          */
  DRG_N_COUNTERS = 8
};

/* rollout flags */
#define DRG_RF_DRAW_RULES 0x1u /* stop on is_draw (base.py:369-376); off = play until no moves (test/_test_helpers.py:44-63) */

/* ---- context ------------------------------------------------------------ */
int drg_abi_version(void);
/* bind the calling process to CUDA device `device`, upload geometry tables */
int drg_init(int device);
int drg_shutdown(void);
const char* drg_last_error(void);
/* number of squares (50 / 32) of a variant, <0 on bad id (BaseBoard.SQUARES_COUNT, base.py:72) */
int drg_squares(int variant);
/* how many kernels this library launched since drg_init (bench.py: gpu_launches) */
int64_t drg_launch_count(void);
/* host copy of the neighbour table NB[64][8] (4 diagonal + 4 orthogonal directions, 0xFF = off
 * board) the kernels use in place of MOVE_TGT / JUMP_TGT / KING_RAYS / ORTHO_* (standard.py:29-70,
 * frisian.py:110-204).  Needs no GPU. */
int drg_geometry_table(int squares, uint8_t* nb /* [512] */);
/* measurement hook: while enabled, CUDA events bracket the k_emit launch (the dominant kernel) on its stream;
 * k_emit_ms (may be NULL) receives the duration of the most recent bracketed launch.  bench.py's roofline. */
int drg_profile(int enable, float* k_emit_ms);
/* cudaMemsetAsync on `stream` (bench.py measures the pure-write ceiling of the GPU with it) */
int drg_memset_async(void* dev_ptr, int value, size_t bytes, void* stream);
/* pinned host memory for the _host entry points (optional; any host pointer works) */
void* drg_host_alloc(size_t bytes);
int drg_host_free(void* p);

/* ---- device entry points ------------------------------------------------ */

/* len(board.legal_moves) per board.
 * Replaces: StandardBoard.legal_moves standard.py:99-105, american.py:93-96,
 * russian.py:119-129, brazilian.py:33-39, frisian.py:256-269 (count only). */
int drg_movegen_count(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk,
                      const uint64_t* bm, const uint64_t* bk, const uint8_t* turn,
                      uint32_t* counts, void* stream);

/* exclusive prefix sum of counts into offsets[0..n] (offsets[n] = total). */
int drg_scan_counts(int64_t n, const uint32_t* counts, uint64_t* offsets, void* stream);

/* board.legal_moves in canonical order for every board, written at
 * moves[offsets[i] .. offsets[i+1]) (at most offsets[i+1]-offsets[i] records;
 * a board with more moves bumps *overflow if overflow != NULL).
 * mask   : NULL or uint8[n][S*S]  -- legal_moves_mask, base.py:807-838.
 * tensor : NULL or float[n][4][S] -- to_tensor(perspective=turn), base.py:753-805.
 * counts : NULL or uint32[n]      -- len(legal_moves) (same as drg_movegen_count). */
int drg_movegen_emit(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk,
                     const uint64_t* bm, const uint64_t* bk, const uint8_t* turn,
                     const uint64_t* offsets, DrgMove* moves, uint8_t* mask, float* tensor,
                     uint32_t* counts, uint64_t* overflow, void* stream);

/* count -> scan -> emit in one stream-ordered call without any host synchronisation: counts[n], offsets[n+1]
 * and the records are all written by this call; `moves` holds moves_cap records -- boards whose slice would
 * pass the end are clipped and bump *overflow (DEVICE uint64, may be NULL).  With a mask the rows are written
 * by their own kernel from the start of the call while counts, offsets and records are produced beside it on
 * the library's second stream (same results as drg_movegen_count + drg_scan_counts + drg_movegen_emit). */
int drg_movegen(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                const uint64_t* bk, const uint8_t* turn, uint32_t* counts, uint64_t* offsets,
                DrgMove* moves, int64_t moves_cap, uint8_t* mask, float* tensor, uint64_t* overflow,
                void* stream);

/* board.push(move) for every board (base.py:215-293), state updated in place.
 * status : NULL or uint8[n]; 1 where the reference would raise ValueError
 *          (base.py:241-245) -- that board is left unchanged. */
int drg_apply(int variant, int64_t n, uint64_t* wm, uint64_t* wk, uint64_t* bm, uint64_t* bk,
              uint8_t* turn, uint8_t* clock, const DrgMove* chosen, uint8_t* status,
              void* stream);

/* pure-movegen perft (recursion over legal_moves/push, bulk count
 * at depth 1, no draw/game_over test) summed over the n_roots boards.
 * All roots must have the same side to move.  Synchronises; nodes is a HOST
 * pointer receiving the total; counters (HOST, DRG_N_COUNTERS uint64, may be
 * NULL) receives DRG_C_NODES / DRG_C_MOVEGEN / DRG_C_OVERFLOW / DRG_C_ILLEGAL. */
int drg_perft(int variant, int64_t n_roots, const uint64_t* wm, const uint64_t* wk,
              const uint64_t* bm, const uint64_t* bk, int turn, int depth, uint64_t* nodes,
              uint64_t* counters, void* stream);

/* drg_perft for one of `world` GPUs: every rank calls it with the same roots; the plies above the first frontier
 * that is wide enough (16 Ki boards per rank, or the last but one ply) are expanded identically everywhere, that
 * frontier is dealt by a hash of the board's content (so the deal does not depend on the order in which a GPU
 * happened to write it), and the rank walks its share.  *nodes / counters are THIS rank's; the sum over the ranks is
 * the perft value.  counters_dev (DEVICE uint64[DRG_N_COUNTERS], may be NULL) receives the same vector in device
 * memory, ready for one NCCL all-reduce without a host round trip. */
int drg_perft_sharded(int variant, int64_t n_roots, const uint64_t* wm, const uint64_t* wk,
                      const uint64_t* bm, const uint64_t* bk, int turn, int depth, int rank, int world,
                      uint64_t* nodes, uint64_t* counters, uint64_t* counters_dev, void* stream);

/* random self-play (examples/custom_agents.py:230-239 pattern): game g plays
 * legal_moves[idx] with idx = ((splitmix64(seed + (game_id0+g)*0x9E3779B97F4A7C15
 * + ply*0xBF58476D1CE4E5B9) >> 32) * n_moves) >> 32 until game_over
 * (base.py:369-376, breakthrough.py:28-31) when DRG_RF_DRAW_RULES is set, or
 * until no legal move otherwise; never more than stop_ply[g] (or max_plies if
 * stop_ply == NULL) plies.  Start position: the arrays (in/out, device) hold
 * the start boards on entry and the final boards on return.
 * result : NULL or uint8[n] 0 '-' unfinished, 1 '1-0', 2 '0-1', 3 '1/2-1/2' (base.py:378-393,
 *          antidraughts.py:31-37, breakthrough.py:34-39).
 * plies  : NULL or uint16[n] plies played.
 * counters: NULL or DEVICE uint64[DRG_N_COUNTERS], accumulated with atomics. */
int drg_rollout(int variant, int64_t n_games, uint64_t seed, uint64_t game_id0, int max_plies,
                const uint16_t* stop_ply, uint32_t flags, uint64_t* wm, uint64_t* wk,
                uint64_t* bm, uint64_t* bk, uint8_t* turn, uint8_t* clock, uint8_t* result,
                uint16_t* plies, uint64_t* counters, void* stream);

/* ONE ply of self-play for every running game, using the move lists drg_movegen_emit just wrote for the
 * current positions (counts / offsets / moves), so the mask / tensor exported by that emit call describe the
 * position the move is chosen from (examples/reinforcement_learning.py:103-112).  Same order of checks and
 * RNG as drg_rollout.  choice : NULL (counter RNG) or int32[n] move index per game (policy sampling).
 * game_ids: NULL (game i has id game_id0 + i) or uint64[n] explicit ids (after finished games were compacted away).
 * plies  : uint16[n] plies played so far (in/out, start at 0).
 * state  : uint8[n] in/out: 0 running, 1 '1-0', 2 '0-1', 3 '1/2-1/2', 4 stopped at the ply limit.
 * hist   : uint32[54][n] scratch holding the last nine pushed moves (threefold rule, base.py:355-366).
 * running: DEVICE uint64, incremented once per game that made a move in this call.
 * moves_cap: records `moves` holds; a game whose chosen record lies beyond it (drg_movegen clipped the board) is
 *          stopped with state 4 and counted like a rejected move instead of reading past the buffer. */
int drg_selfplay_step(int variant, int64_t n, uint64_t seed, uint64_t game_id0, int max_plies,
                      const uint16_t* stop_ply, uint32_t flags, uint64_t* wm, uint64_t* wk,
                      uint64_t* bm, uint64_t* bk, uint8_t* turn, uint8_t* clock, uint16_t* plies,
                      uint8_t* state, uint32_t* hist, const uint32_t* counts, const uint64_t* offsets,
                      const DrgMove* moves, int64_t moves_cap, const int32_t* choice,
                      const uint64_t* game_ids, uint64_t* running, void* stream);

/* The one collective of the path (SURVEY 8(e), 8(b) drg_allreduce_counters): the ranks' uint64 counter vectors (perft
 * nodes, game results ...) summed over NCCL / NVLink.  One process per GPU; the work itself is sharded with no data
 * exchange.  The library binds at run time to the NCCL the process already has loaded (torch's, when
 * torch.distributed runs beside it) or to libnccl.so.2, and owns its communicator:
 *   drg_nccl_unique_id  rank 0 creates the 128-byte id, the caller hands it to every rank (any side channel);
 *   drg_comm_init       collective over the `world` ranks, the context's device must be the rank's GPU;
 *   drg_allreduce_counters  in-place sum of DEVICE uint64[n], stream-ordered on `stream`;
 *   drg_comm_destroy.
 * DRG_E_CUDA with a message when no NCCL can be bound -- callers then use their own process group. */
int drg_nccl_unique_id(uint8_t* id128);
int drg_comm_init(const uint8_t* id128, int rank, int world, void** comm);
int drg_allreduce_counters(void* comm, uint64_t* counters, int n, void* stream);
int drg_comm_destroy(void* comm);

/* Between two plies of the self-play loop (examples/reinforcement_learning.py:95-137 keeps one Board per game and
 * drops it when game_over is reached): the working set of games shrinks.  Games that are over (state != 0) are
 * written to the caller's full-size result arrays at position ids[i]; running games (state == 0) are packed, in
 * their order, to the front of the destination working set, with their 9-move history ([54][n] -> [54][n_running])
 * and their ids.  One launch chain (flag, scan, gather) instead of a gather per array.
 *   src : n games; dst : room for n games, dst == NULL: last call, EVERY game is written to `fin`, nothing is packed;
 *   fin : wm, wk, bm, bk, turn, clock, plies, state of all games of the job (hist / ids / game_ids unused).
 *   n_running : host int64 out (may be NULL when dst is NULL); the call waits for the stream to deliver it. */
typedef struct DrgGames {
  uint64_t *wm, *wk, *bm, *bk;
  uint8_t *turn, *clock;
  uint16_t* plies;
  uint8_t* state;
  uint32_t* hist;
  int64_t* ids;
  uint64_t* game_ids;
} DrgGames;
int drg_selfplay_compact(int64_t n, const DrgGames* src, const DrgGames* dst, const DrgGames* fin,
                         int64_t* n_running, void* stream);

/* The policy-head end of the RL loop (examples/reinforcement_learning.py:103-112: logits[~mask] = -1e9;
 * softmax(logits / temp); multinomial; board.index_to_move(action)) for a batch, on the device, from the move lists
 * drg_movegen just wrote (counts / offsets / moves; `moves` holds moves_cap records).
 *   logits  : float[n][S*S] policy head, index = from * S + to (BaseBoard.move_to_index, base.py:840-859).
 *   mode    : 0 sample from softmax(logits / temperature) restricted to the legal actions (the set bytes of the
 *             board's legal_moves_mask row); counter RNG keyed by (seed, game id, plies[i]) like drg_rollout.
 *             1 argmax over the legal actions (ties: lowest action index).
 *   action  : int32[n] out, from * S + to; -1 for a board without legal moves.
 *   choice  : int32[n] out, index into board i's move list of the FIRST record with that (from, to) -- the move
 *             BaseBoard.index_to_move returns (base.py:861-891); feeds drg_selfplay_step's `choice`.
 * game_ids (uint64[n]) / plies (uint16[n]) may be NULL (game i has id game_id0 + i / ply 0).
 * drg_action_to_choice is the index_to_move half alone, for callers that sample with their own code:
 * choice[i] = -1 where the reference raises ValueError (no legal move with these squares). */
int drg_sample_action(int variant, int64_t n, const float* logits, const uint32_t* counts,
                      const uint64_t* offsets, const DrgMove* moves, int64_t moves_cap, float temperature,
                      int mode, uint64_t seed, uint64_t game_id0, const uint64_t* game_ids,
                      const uint16_t* plies, int32_t* action, int32_t* choice, void* stream);
int drg_action_to_choice(int variant, int64_t n, const int32_t* action, const uint32_t* counts,
                         const uint64_t* offsets, const DrgMove* moves, int64_t moves_cap, int32_t* choice,
                         void* stream);

/* ---- host-buffer entry points (H2D + kernels + D2H inside) --------------- */

/* legal_moves (+ optional mask / tensor) for n boards given as HOST arrays.
 * counts[n] always written; offsets (NULL or uint64[n+1]) receives the exclusive
 * prefix sum of counts, i.e. board i's records are moves[offsets[i] .. offsets[i+1]).
 * moves/mask/tensor may be NULL.  moves_cap = number of records `moves` can hold;
 * *total receives sum(counts).  Returns DRG_E_OVERFLOW (with counts, offsets and
 * *total valid) when total > moves_cap.  Large masks cross PCIe bit-packed and are
 * widened to one byte per index by host threads (DRG_HOST_THREADS, default <= 16). */
int drg_legal_moves_host(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk,
                         const uint64_t* bm, const uint64_t* bk, const uint8_t* turn,
                         uint32_t* counts, uint64_t* offsets, DrgMove* moves, int64_t moves_cap,
                         int64_t* total, uint8_t* mask, float* tensor);

/* The same call with the legal-move-mask rows delivered PACKED: mask_bits (uint8[(n*S*S + 7) / 8], may be NULL) is the
 * [n, S*S] bool array as one flat little-endian bit stream -- mask[i][a] is bit ((i*S*S + a) & 7) of byte
 * ((i*S*S + a) >> 3); numpy: np.unpackbits(mask_bits, bitorder="little")[: n*S*S].reshape(n, S*S).  The rows are 99.7 %
 * zeros; this is how they cross PCIe in either call, and here nothing widens them on the host, so the call is bound
 * by the link instead of by the host's memory system (bench.py reports both). */
int drg_legal_moves_host_bits(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk,
                              const uint64_t* bm, const uint64_t* bk, const uint8_t* turn,
                              uint32_t* counts, uint64_t* offsets, DrgMove* moves, int64_t moves_cap,
                              int64_t* total, uint8_t* mask_bits, float* tensor);

int drg_apply_host(int variant, int64_t n, uint64_t* wm, uint64_t* wk, uint64_t* bm,
                   uint64_t* bk, uint8_t* turn, uint8_t* clock, const DrgMove* chosen,
                   uint8_t* status);

int drg_perft_host(int variant, int64_t n_roots, const uint64_t* wm, const uint64_t* wk,
                   const uint64_t* bm, const uint64_t* bk, int turn, int depth,
                   uint64_t* nodes, uint64_t* counters);

int drg_rollout_host(int variant, int64_t n_games, uint64_t seed, uint64_t game_id0,
                     int max_plies, const uint16_t* stop_ply, uint32_t flags, uint64_t* wm,
                     uint64_t* wk, uint64_t* bm, uint64_t* bk, uint8_t* turn, uint8_t* clock,
                     uint8_t* result, uint16_t* plies, uint64_t* counters);

/* ---- search-leaf service ------------------------------------------------------------------------
 * What a batched search (MCTS leaf expansion, the frontier of AlphaBetaEngine.negamax / quiescence,
 * engines/alpha_beta.py:330-465) asks of the move generator: the children of N leaves in move-list order and a
 * static score for every one of them.  The search itself stays with the caller.
 *
 * drg_children: child j = copy of its parent with moves[j] pushed (board.copy(); board.push(move),
 * base.py:215-293, 672-700), the parent of j being the board i with offsets[i] <= j < offsets[i+1] (the layout
 * drg_movegen / drg_movegen_emit produce; offsets has n+1 entries).  Outputs hold `total` child slots; the
 * min(total, offsets[n]) real children are written and the remaining slots are left untouched, so a caller that sized its
 * buffers by capacity never has to read the exact total back (no host synchronisation in a movegen ->
 * children -> evaluate chain).  oclock, parent (index of the parent board) and status may be NULL.  Where the
 * reference raises ValueError (base.py:241-245) the child is the unchanged parent and status[j] = 1.
 *
 * drg_evaluate: AlphaBetaEngine.evaluate (alpha_beta.py:205-244) per board: material (man 1.0, king 2.5) plus the
 * four piece-square tables of alpha_beta.py:27-57,166-177 (10 rows for S = 50, 8 for S = 32), relative to the side
 * to move, float64, bit-exact with the reference (additions in numpy's pairwise order, no FMA contraction).
 * drg_pst_table: the tables themselves (host doubles, out[4*S]: man black, man white, king black, king white);
 * returns 4*S. */
int drg_children(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                 const uint64_t* bk, const uint8_t* turn, const uint8_t* clock, const uint64_t* offsets,
                 int64_t total, const DrgMove* moves, uint64_t* owm, uint64_t* owk, uint64_t* obm,
                 uint64_t* obk, uint8_t* oturn, uint8_t* oclock, uint32_t* parent, uint8_t* status,
                 void* stream);
int drg_evaluate(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                 const uint64_t* bk, const uint8_t* turn, double* score, void* stream);
int drg_pst_table(int variant, double* out);
/* host-buffer form of drg_evaluate (H2D + kernel + D2H, synchronous) */
int drg_evaluate_host(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk,
                      const uint64_t* bm, const uint64_t* bk, const uint8_t* turn, double* score);

/* ---- bulk FEN codec (host only, threaded; no device work) -----------------------------------------------------
 * Replaces per-position loops over BaseBoard.from_fen (base.py:435-495) and BaseBoard.fen (base.py:408-433).
 * Text is one flat buffer; row i is text[off[i] .. off[i+1]) (no terminators).
 *
 * drg_fen_decode: canonical rows -- optional [FEN "..."] wrapper around <W|B>:W<list>:B<list>, upper case, list
 * tokens empty, <digits> or K<digits> with 1 <= square <= S; white list written first, then black, the last
 * token owns a square (from_fen's cell-array order) -- are decoded and get status[i] = 0.  Any other spelling
 * from_fen tolerates or rejects (lower case, G/P tokens, legacy four-field form, :H0:F2 tails, square 0 / > S,
 * malformed tokens) gets status[i] = 1 and zero bitboards; the caller runs its full parser on those rows.
 *
 * drg_fen_encode: text == NULL -> sizing pass, off[0..n] = exclusive scan of the row lengths; otherwise rows are
 * written at off[i] (text_cap >= off[n], else DRG_E_OVERFLOW).  Output is byte-identical to BaseBoard.fen,
 * including squares held by both colours (listed on both sides) and men shadowing kings of the same colour. */
#define DRG_FEN_MAX 448 /* upper bound of one encoded row, any variant */
int drg_fen_decode(int variant, int64_t n, const char* text, const int64_t* off, uint64_t* wm,
                   uint64_t* wk, uint64_t* bm, uint64_t* bk, uint8_t* turn, uint8_t* status);
int drg_fen_encode(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk,
                   const uint64_t* bm, const uint64_t* bk, const uint8_t* turn, char* text,
                   int64_t text_cap, int64_t* off);

#ifdef __cplusplus
}
#endif
#endif /* DRG_H_ */
