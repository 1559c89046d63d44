// drg_host.cpp -- host side of the packed legal-move-mask transport used by drg_legal_moves_host.
//
// The mask rows the kernels produce (bool [n, S*S], base.py:807-838) are 99.7 % zeros; sending them over
// PCIe as bytes makes the host-buffer path link-bound (2.6 GB per Mi Standard boards).  The device therefore
// packs them 8:1 (k_pack_mask) and the host widens the bit stream back to the reference's one-byte-per-index
// layout while later chunks are still in flight.  Nothing here knows draughts: it is a transport decoder.
#if defined(__linux__)
#include <sched.h>
#endif

#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace drg_host {

static uint64_t g_lut[256];
static std::atomic<bool> g_lut_ready{false};

static void init_lut() {
  if (g_lut_ready.load(std::memory_order_acquire)) return;
  for (int b = 0; b < 256; b++) {
    uint64_t v = 0;
    for (int k = 0; k < 8; k++)
      if (b & (1 << k)) v |= 1ull << (8 * k);
    g_lut[b] = v;
  }
  g_lut_ready.store(true, std::memory_order_release);
}

// out[j] = bit j of `bits`, for j in [lo, hi); lo and hi are multiples of 8 except possibly the global end
static void expand_scalar(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi) {
  size_t j = lo;
  for (; j + 8 <= hi; j += 8) {
    const uint64_t v = g_lut[bits[j >> 3]];
    std::memcpy(out + j, &v, 8);
  }
  for (; j < hi; j++) out[j] = (bits[j >> 3] >> (j & 7)) & 1;
}

#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw"))) static void expand_avx512(const uint8_t* bits, uint8_t* out, size_t lo,
                                                                       size_t hi) {
  size_t j = lo;
  const __m512i one = _mm512_set1_epi8(1);
  if (((reinterpret_cast<uintptr_t>(out) + j) & 63) == 0) {
    // the rows are written once and not read back here: non-temporal stores skip the read-for-ownership
    for (; j + 64 <= hi; j += 64) {
      uint64_t m;
      std::memcpy(&m, bits + (j >> 3), 8);
      _mm512_stream_si512(reinterpret_cast<__m512i*>(out + j), _mm512_maskz_mov_epi8((__mmask64)m, one));
    }
    _mm_sfence();
  } else {
    for (; j + 64 <= hi; j += 64) {
      uint64_t m;
      std::memcpy(&m, bits + (j >> 3), 8);
      _mm512_storeu_si512(reinterpret_cast<void*>(out + j), _mm512_maskz_mov_epi8((__mmask64)m, one));
    }
  }
  expand_scalar(bits, out, j, hi);
}
#endif

static void expand_range(const uint8_t* bits, uint8_t* out, size_t lo, size_t hi) {
#if defined(__x86_64__)
  static const bool has512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f");
  if (has512) {
    expand_avx512(bits, out, lo, hi);
    return;
  }
#endif
  expand_scalar(bits, out, lo, hi);
}

int default_threads() {
  if (const char* e = std::getenv("DRG_HOST_THREADS")) {
    const int v = std::atoi(e);
    if (v > 0) return v > 256 ? 256 : v;
  }
  // the cores this process may run on (a container's cpuset, a torchrun rank's binding), not the machine's
  unsigned hw = 0;
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) hw = (unsigned)CPU_COUNT(&set);
#endif
  if (hw == 0) hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) {  // one process per GPU (torchrun): share the cores
    const int w = std::atoi(e);
    if (w > 1) hw = hw / (unsigned)w > 0 ? hw / (unsigned)w : 1;
  }
  return hw > 32 ? 32 : (int)hw;  // the widening is bound by memory bandwidth long before that
}

// Widen `total` bits into bytes with `threads` workers.  The bit stream arrives in `nchunks` equal chunks of
// `chunk` output bytes (multiple of 64); worker 0 calls wait_chunk(ctx, c) (blocks until the device-to-host copy
// of chunk c has landed) and then releases the chunk to everybody, so the expansion overlaps the transfer.
void expand_pipelined(const uint8_t* bits, uint8_t* out, size_t total, size_t chunk, int nchunks,
                      void (*wait_chunk)(void*, int), void* ctx, int threads) {
  init_lut();
  if (threads < 1) threads = 1;
  std::atomic<int> ready_count{0};
  std::atomic<int>* ready = &ready_count;
  auto work = [&](int tid) {
    for (int c = 0; c < nchunks; c++) {
      if (tid == 0) {
        wait_chunk(ctx, c);
        ready->store(c + 1, std::memory_order_release);
      }
      while (ready->load(std::memory_order_acquire) <= c) std::this_thread::yield();
      const size_t c_lo = (size_t)c * chunk;
      const size_t c_hi = c_lo + chunk < total ? c_lo + chunk : total;
      if (c_lo >= c_hi) continue;
      const size_t len = c_hi - c_lo;
      size_t per = (len / threads + 63) & ~(size_t)63;  // slices are multiples of 64 output bytes
      if (per == 0) per = 64;
      const size_t lo = c_lo + (size_t)tid * per;
      if (lo >= c_hi) continue;
      const size_t hi = lo + per < c_hi ? lo + per : c_hi;
      expand_range(bits, out, lo, hi);
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(threads - 1);
  for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
}

}  // namespace drg_host
