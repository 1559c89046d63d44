// Host-side widening of bit-packed mask rows (drg_host.cpp): which store strategy is fastest on the GPU box's CPU?
//   g++ -O3 -march=native -pthread tools/micro/widen_bw.cpp -o tools/micro/widen_bw && tools/micro/widen_bw [threads]
// Output bytes 2.62 GB (1 Mi rows x 2500), 0.33 % of the bits set -- the e2e path's real shape.
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static const size_t kTotal = (size_t)1048576 * 2500;

typedef void (*Fn)(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi);

static void nt_maskz(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi) {
  const __m512i one = _mm512_set1_epi8(1);
  for (size_t j = lo; j + 64 <= hi; j += 64) {
    uint64_t m;
    memcpy(&m, bits + (j >> 3), 8);
    _mm512_stream_si512((__m512i*)(out + j), _mm512_maskz_mov_epi8((__mmask64)m, one));
  }
  _mm_sfence();
}
static void st_maskz(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi) {
  const __m512i one = _mm512_set1_epi8(1);
  for (size_t j = lo; j + 64 <= hi; j += 64) {
    uint64_t m;
    memcpy(&m, bits + (j >> 3), 8);
    _mm512_store_si512((__m512i*)(out + j), _mm512_maskz_mov_epi8((__mmask64)m, one));
  }
}
static void nt_maskz_prefetch(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi) {
  const __m512i one = _mm512_set1_epi8(1);
  for (size_t j = lo; j + 512 <= hi; j += 512) {  // one cache line of bits -> 8 output lines
    _mm_prefetch((const char*)(bits + (j >> 3) + 1024), _MM_HINT_T0);
    uint64_t m[8];
    memcpy(m, bits + (j >> 3), 64);
    for (int k = 0; k < 8; k++)
      _mm512_stream_si512((__m512i*)(out + j + 64 * k), _mm512_maskz_mov_epi8((__mmask64)m[k], one));
  }
  _mm_sfence();
}
static void nt_maskz_blocked(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi) {
  const __m512i one = _mm512_set1_epi8(1);
  const size_t BLK = 1 << 20;  // 1 MB of output = 128 KB of bits, pulled into cache first, then a pure store burst
  for (size_t b = lo; b < hi; b += BLK) {
    const size_t e = b + BLK < hi ? b + BLK : hi;
    uint64_t acc = 0;
    for (size_t j = b; j < e; j += 512) acc += *(const volatile uint64_t*)(bits + (j >> 3));
    if (acc == 0x123456789abcdefull) out[b] = 0;
    for (size_t j = b; j + 64 <= e; j += 64) {
      uint64_t m;
      memcpy(&m, bits + (j >> 3), 8);
      _mm512_stream_si512((__m512i*)(out + j), _mm512_maskz_mov_epi8((__mmask64)m, one));
    }
  }
  _mm_sfence();
}
static void nt_zero_only(const uint8_t*, uint8_t* out, size_t lo, size_t hi) {
  const __m512i z = _mm512_setzero_si512();
  for (size_t j = lo; j + 64 <= hi; j += 64) _mm512_stream_si512((__m512i*)(out + j), z);
  _mm_sfence();
}
static void memset_only(const uint8_t*, uint8_t* out, size_t lo, size_t hi) { memset(out + lo, 0, hi - lo); }
template <size_t BLK>
static void memset_patch(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi) {
  for (size_t b = lo; b < hi; b += BLK) {
    const size_t e = b + BLK < hi ? b + BLK : hi;
    memset(out + b, 0, e - b);
    for (size_t j = b; j + 64 <= e; j += 64) {
      uint64_t m;
      memcpy(&m, bits + (j >> 3), 8);
      while (m) {
        out[j + __builtin_ctzll(m)] = 1;
        m &= m - 1;
      }
    }
  }
}
static void stosb_patch(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi) {
  const size_t BLK = 65536;
  for (size_t b = lo; b < hi; b += BLK) {
    const size_t e = b + BLK < hi ? b + BLK : hi;
    void* d = out + b;
    size_t n = e - b;
    __asm__ volatile("rep stosb" : "+D"(d), "+c"(n) : "a"(0) : "memory");
    for (size_t j = b; j + 64 <= e; j += 64) {
      uint64_t m;
      memcpy(&m, bits + (j >> 3), 8);
      while (m) {
        out[j + __builtin_ctzll(m)] = 1;
        m &= m - 1;
      }
    }
  }
}

static double run(Fn fn, const uint8_t* bits, uint8_t* out, int threads) {
  double best = 1e9;
  for (int rep = 0; rep < 4; rep++) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    const size_t per = ((kTotal / threads) + 4095) & ~(size_t)4095;
    for (int t = 0; t < threads; t++) {
      const size_t lo = (size_t)t * per, hi = lo + per < kTotal ? lo + per : kTotal;
      if (lo < hi) pool.emplace_back(fn, bits, out, lo, hi);
    }
    for (auto& th : pool) th.join();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 16;
  uint8_t* out = (uint8_t*)aligned_alloc(4096, kTotal + 4096);
  uint8_t* bits = (uint8_t*)aligned_alloc(4096, kTotal / 8 + 4096);
  memset(out, 1, kTotal);
  memset(bits, 0, kTotal / 8);
  uint64_t x = 88172645463325252ull;
  for (size_t k = 0; k < kTotal / 300; k++) {  // ~0.33 % of the bits
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    const size_t b = x % kTotal;
    bits[b >> 3] |= 1 << (b & 7);
  }
  struct { const char* name; Fn fn; } v[] = {
      {"nt_maskz (current)", nt_maskz}, {"nt_maskz_prefetch", nt_maskz_prefetch}, {"nt_maskz_blocked", nt_maskz_blocked},
      {"st_maskz (cached stores)", st_maskz}, {"nt_zero_only", nt_zero_only},
      {"memset_only", memset_only}, {"memset64K+patch", memset_patch<65536>}, {"memset512K+patch", memset_patch<524288>},
      {"rep_stosb64K+patch", stosb_patch}};
  for (auto& e : v) {
    const double ms = run(e.fn, bits, out, threads);
    printf("%-28s %2d threads  %7.2f ms  %7.1f GB/s\n", e.name, threads, ms, kTotal / ms / 1e6);
  }
  return 0;
}
