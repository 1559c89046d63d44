// drg_geom.h -- board geometry of the 10x10 (50 squares) and 8x8 (32 squares) draughts boards.
//
// Replaces the import-time tables of the reference: MOVE_TGT / JUMP_TGT / KING_RAYS
// (draughts/boards/standard.py:29-70, american.py:26-56, russian.py:36-79) and the Frisian
// ORTHO_RAYS / ORTHO_JUMP (frisian.py:110-204).  Built differently on purpose: everything is
// derived from board coordinates, as (a) a per-square neighbour table NB[sq][8] (4 diagonal +
// 4 orthogonal directions; a jump target is NB[NB[sq][d]][d], a ray is the NB chain) that the
// capture DFS keeps in shared memory, and (b) whole-bitboard step functions used for the
// bit-parallel quiet moves and capture pre-checks.  tests/test_host_abi.py checks the table
// against the reference's (tests/golden/tables.json) through drg_geometry_table().
//
// Square numbering (standard.py:14-18): sq = row * W + col_in_row, row 0 = top (black's back
// rank); even rows hold board columns 1,3,5,...; odd rows 0,2,4,....  White moves toward row 0.
// Directions: 0 up-right, 1 up-left, 2 down-right, 3 down-left (standard.py:30);
//             4 up, 5 right, 6 down, 7 left (frisian.py:113; two board squares per step).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DRG_HD __host__ __device__ __forceinline__
#else
#define DRG_HD inline
#endif

namespace drg {

constexpr int kNone = 0xFF;

template <int S>
struct Geo;

template <>
struct Geo<50> {
  using BB = uint64_t;
  static constexpr int NSQ = 50, W = 5, ROWS = 10;
  static constexpr BB MASK = (1ull << 50) - 1;
  static constexpr BB EVEN = 0x7C1F07C1Full | (0x1Full << 40);  // rows 0,2,4,6,8
  static constexpr BB ODD = MASK & ~EVEN;
  static constexpr BB EVEN_RIGHT = (1ull << 4) | (1ull << 14) | (1ull << 24) | (1ull << 34) | (1ull << 44);
  static constexpr BB ODD_LEFT = (1ull << 5) | (1ull << 15) | (1ull << 25) | (1ull << 35) | (1ull << 45);
  static constexpr BB COL_FIRST = 0x0084210842108421ull & MASK;  // col_in_row == 0
  static constexpr BB COL_LAST = (COL_FIRST << 4) & MASK;        // col_in_row == W-1
  static constexpr BB ROW_FIRST = 0x1Full;
  static constexpr BB ROW_LAST = 0x1Full << 45;
};

template <>
struct Geo<32> {
  using BB = uint32_t;
  static constexpr int NSQ = 32, W = 4, ROWS = 8;
  static constexpr BB MASK = 0xFFFFFFFFu;
  static constexpr BB EVEN = 0x0F0F0F0Fu;
  static constexpr BB ODD = 0xF0F0F0F0u;
  static constexpr BB EVEN_RIGHT = 0x08080808u;
  static constexpr BB ODD_LEFT = 0x10101010u;
  static constexpr BB COL_FIRST = 0x11111111u;
  static constexpr BB COL_LAST = 0x88888888u;
  static constexpr BB ROW_FIRST = 0x0000000Fu;
  static constexpr BB ROW_LAST = 0xF0000000u;
};

// One step of every set square in direction D (squares that would leave the board vanish).
template <int S, int D>
DRG_HD typename Geo<S>::BB step(typename Geo<S>::BB x) {
  using G = Geo<S>;
  using BB = typename G::BB;
  constexpr int W = G::W;
  if (D == 0) return BB(((x & G::EVEN & ~G::EVEN_RIGHT) >> (W - 1)) | ((x & G::ODD) >> W));
  if (D == 1) return BB(((x & G::EVEN) >> W) | ((x & G::ODD & ~G::ODD_LEFT) >> (W + 1)));
  if (D == 2) return BB((((x & G::EVEN & ~G::EVEN_RIGHT) << (W + 1)) | ((x & G::ODD) << W)) & G::MASK);
  if (D == 3) return BB((((x & G::EVEN) << W) | ((x & G::ODD & ~G::ODD_LEFT) << (W - 1))) & G::MASK);
  if (D == 4) return BB(x >> (2 * W));
  if (D == 5) return BB(((x & ~G::COL_LAST) << 1) & G::MASK);
  if (D == 6) return BB((x << (2 * W)) & G::MASK);
  return BB((x & ~G::COL_FIRST) >> 1);
}

// opposite direction (used to walk a landing square back to the piece that jumps)
DRG_HD constexpr int opposite(int d) { return d < 4 ? 3 - d : 4 + ((d - 4 + 2) & 3); }

// neighbour of one square from board coordinates (host-side table construction)
template <int S>
inline int neighbour(int sq, int d) {
  constexpr int W = Geo<S>::W, ROWS = Geo<S>::ROWS;
  const int row = sq / W, c = sq % W;
  const int bcol = 2 * c + ((row & 1) ? 0 : 1);
  static const int dr[8] = {-1, -1, 1, 1, -2, 0, 2, 0};
  static const int dc[8] = {1, -1, 1, -1, 0, 2, 0, -2};
  const int r2 = row + dr[d], b2 = bcol + dc[d];
  if (r2 < 0 || r2 >= ROWS || b2 < 0 || b2 >= 2 * W) return kNone;
  if (((r2 & 1) ? 0 : 1) != (b2 & 1)) return kNone;  // not a dark square (cannot happen for these deltas)
  return r2 * W + b2 / 2;
}

// NB table, S rows x 8 directions, kNone where the step leaves the board.
template <int S>
inline void build_nb(uint8_t* nb /*[S*8]*/) {
  for (int i = 0; i < S * 8; i++) nb[i] = kNone;
  for (int sq = 0; sq < S; sq++)
    for (int d = 0; d < 8; d++) nb[sq * 8 + d] = (uint8_t)neighbour<S>(sq, d);
}

// ray masks and jump landings derived from the NB table (same information as KING_RAYS / ORTHO_RAYS /
// JUMP_TGT / ORTHO_JUMP of the reference, as bit masks)
template <int S>
inline void build_ray_jmp(typename Geo<S>::BB* ray /*[S*8]*/, uint8_t* jmp /*[S*8]*/, const uint8_t* nb) {
  using BB = typename Geo<S>::BB;
  for (int i = 0; i < S * 8; i++) {
    ray[i] = 0;
    jmp[i] = kNone;
  }
  for (int sq = 0; sq < S; sq++)
    for (int d = 0; d < 8; d++) {
      BB m = 0;
      for (int t = nb[sq * 8 + d]; t != kNone; t = nb[t * 8 + d]) m |= BB(1) << t;
      ray[sq * 8 + d] = m;
      const int mid = nb[sq * 8 + d];
      if (mid != kNone) jmp[sq * 8 + d] = nb[mid * 8 + d];
    }
}

}  // namespace drg
