// drg_fen.cpp -- bulk FEN codec between text and the SoA bitboards the kernels read (include/drg.h, drg_fen_*).
//
// The reference converts one position at a time (`BaseBoard.fen`, base.py:408-433; `BaseBoard.from_fen`,
// base.py:435-495; 18 us and 27 us per call).  Feeding a million-board batch from a FEN file, or writing the
// children of a batch back as FENs, needs the same conversion at memory speed, so it lives here as
// threaded host code over one flat text buffer + offsets.
//
// Decoding restates from_fen for the CANONICAL spellings only -- what `fen` itself emits and what the reference's
// test files contain: an optional [FEN "..."] wrapper around  <turn>:W<list>:B<list>, upper case, every list token
// empty, <digits> or K<digits> with 1 <= square <= S.  from_fen writes the white list, then the black list, into one
// cell array (a later token overwrites an earlier one for the same square); that order is kept.  Every other
// spelling (lower case, G/P tokens, the legacy four-field form, :H0:F2 tails, square 0 or > S, malformed tokens)
// is NOT guessed at: the row gets status 1 and the caller's full parser (boards.parse_fen, the regex restatement)
// decides, including which exception to raise.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/drg.h"

namespace drg_host {
int default_threads();
}

namespace {

template <class F>
void parallel_rows(int64_t n, F&& body) {
  int threads = drg_host::default_threads();
  if (n < 4096) threads = 1;
  if (threads <= 1) {
    body((int64_t)0, n);
    return;
  }
  std::vector<std::thread> pool;
  const int64_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) {
    const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
    if (lo >= hi) break;
    pool.emplace_back([&body, lo, hi] { body(lo, hi); });
  }
  for (auto& th : pool) th.join();
}

struct Cells {
  uint64_t wm = 0, wk = 0, bm = 0, bk = 0;
  void put(int sq, int code) {  // code: 0 WM, 1 WK, 2 BM, 3 BK; the last writer owns the square
    const uint64_t bit = 1ull << sq;
    wm &= ~bit; wk &= ~bit; bm &= ~bit; bk &= ~bit;
    (code == 0 ? wm : code == 1 ? wk : code == 2 ? bm : bk) |= bit;
  }
};

// one piece list; p points behind ":W" / ":B".  Returns the first character that is not part of the list, or nullptr
// when a token is not canonical.
const char* decode_list(const char* p, const char* end, int S, int man_code, Cells& c) {
  while (p < end) {
    if (*p == ',') { ++p; continue; }  // empty tokens are skipped by from_fen as well
    if (*p != 'K' && (*p < '0' || *p > '9')) break;
    const int code = man_code + (*p == 'K');
    if (*p == 'K') ++p;
    if (p >= end || *p < '0' || *p > '9') return nullptr;
    int sq = 0, digits = 0;
    while (p < end && *p >= '0' && *p <= '9') {
      sq = sq * 10 + (*p++ - '0');
      if (++digits > 6) return nullptr;
    }
    if (sq < 1 || sq > S) return nullptr;
    if (p < end && *p == 'K') return nullptr;  // "12K3": not a token from_fen understands
    c.put(sq - 1, code);
  }
  return p;
}

bool decode_row(const char* p, const char* end, int S, Cells& c, uint8_t& turn) {
  static const char kOpen[] = "[FEN \"";
  if (end - p >= 8 && std::memcmp(p, kOpen, 6) == 0) {
    if (end[-1] != ']' || end[-2] != '"') return false;
    p += 6;
    end -= 2;
  }
  if (end - p < 5) return false;  // shortest canonical body: "W:W:B"
  if ((p[0] != 'W' && p[0] != 'B') || p[1] != ':' || p[2] != 'W') return false;
  turn = p[0] == 'B';
  p = decode_list(p + 3, end, S, 0, c);
  if (!p || end - p < 2 || p[0] != ':' || p[1] != 'B') return false;
  p = decode_list(p + 2, end, S, 2, c);
  return p == end;
}

inline int put_number(char* out, int v) {  // 1..64
  if (v >= 10) {
    out[0] = (char)('0' + v / 10);
    out[1] = (char)('0' + v % 10);
    return 2;
  }
  out[0] = (char)('0' + v);
  return 1;
}

// `fen` (base.py:408-433): per square ascending, white list takes the man before the king, black list likewise;
// a square held by both colours (Russian overlap) appears in both lists.
int encode_row(int S, uint64_t wm, uint64_t wk, uint64_t bm, uint64_t bk, int turn, char* out) {
  char* q = out;
  std::memcpy(q, "[FEN \"", 6);
  q += 6;
  *q++ = turn ? 'B' : 'W';
  *q++ = ':';
  for (int side = 0; side < 2; side++) {
    *q++ = side ? 'B' : 'W';
    const uint64_t men = side ? bm : wm, kings = (side ? bk : wk) & ~men;
    uint64_t occ = (men | kings) & (S == 64 ? ~0ull : ((1ull << S) - 1));
    bool first = true;
    while (occ) {
      const int sq = __builtin_ctzll(occ);
      occ &= occ - 1;
      if (!first) *q++ = ',';
      first = false;
      if (kings >> sq & 1) *q++ = 'K';
      q += put_number(q, sq + 1);
    }
    if (!side) *q++ = ':';
  }
  *q++ = '"';
  *q++ = ']';
  return (int)(q - out);
}

}  // namespace

extern "C" {

int drg_fen_decode(int variant, int64_t n, const char* text, const int64_t* off, uint64_t* wm, uint64_t* wk,
                   uint64_t* bm, uint64_t* bk, uint8_t* turn, uint8_t* status) {
  const int S = drg_squares(variant);
  if (S < 0) return S;
  if (n < 0 || (n > 0 && (!text || !off || !wm || !wk || !bm || !bk || !turn || !status))) return DRG_E_ARG;
  parallel_rows(n, [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; i++) {
      Cells c;
      uint8_t t = 0;
      const bool ok = off[i + 1] >= off[i] && decode_row(text + off[i], text + off[i + 1], S, c, t);
      if (!ok) c = Cells(), t = 0;
      wm[i] = c.wm; wk[i] = c.wk; bm[i] = c.bm; bk[i] = c.bk;
      turn[i] = t;
      status[i] = ok ? 0 : 1;
    }
  });
  return DRG_OK;
}

int drg_fen_encode(int variant, int64_t n, const uint64_t* wm, const uint64_t* wk, const uint64_t* bm,
                   const uint64_t* bk, const uint8_t* turn, char* text, int64_t text_cap, int64_t* off) {
  const int S = drg_squares(variant);
  if (S < 0) return S;
  if (n < 0 || !off || (n > 0 && (!wm || !wk || !bm || !bk || !turn))) return DRG_E_ARG;
  if (!text) {  // sizing pass: off[0..n] = exclusive scan of the row lengths
    parallel_rows(n, [=](int64_t lo, int64_t hi) {
      char tmp[DRG_FEN_MAX];
      for (int64_t i = lo; i < hi; i++) off[i + 1] = encode_row(S, wm[i], wk[i], bm[i], bk[i], turn[i] & 1, tmp);
    });
    off[0] = 0;
    for (int64_t i = 0; i < n; i++) off[i + 1] += off[i];
    return DRG_OK;
  }
  if (n > 0 && text_cap < off[n]) return DRG_E_OVERFLOW;
  parallel_rows(n, [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; i++) encode_row(S, wm[i], wk[i], bm[i], bk[i], turn[i] & 1, text + off[i]);
  });
  return DRG_OK;
}

}  // extern "C"
