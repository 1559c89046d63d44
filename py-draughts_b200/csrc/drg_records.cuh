// drg_records.cuh -- construction of the packed 32-byte move records (include/drg.h DrgMove) from the capture search's
// state, and the sink that keeps a tree search's best leaves.  No CUDA runtime dependency: the kernels include it, and
// so does the host replay of the per-thread algorithm (tools/devsim).
#pragma once
#include "drg_core.cuh"

namespace drg {

// ---------------------------------------------------------------------------------------------
// move record construction (include/drg.h DrgMove; 8 little-endian words)
// ---------------------------------------------------------------------------------------------
struct Rec {
  uint32_t w[8];
};

__device__ __forceinline__ void rec_quiet(Rec& r, int from, int to, bool king) {
  r.w[0] = r.w[1] = 0;
  r.w[2] = 2u | ((king ? DRG_MF_KINGMOVE : 0u) << 8) | ((uint32_t)from << 16) | ((uint32_t)to << 24);
  r.w[3] = r.w[4] = r.w[5] = r.w[6] = r.w[7] = 0;
}

__device__ __forceinline__ void rec_jump(Rec& r, int from, int victim, int land) {
  const uint64_t c64 = 1ull << victim;
  r.w[0] = (uint32_t)c64;
  r.w[1] = (uint32_t)(c64 >> 32);
  r.w[2] = 2u | ((uint32_t)DRG_MF_CAPTURE << 8) | ((uint32_t)from << 16) | ((uint32_t)land << 24);
  r.w[3] = r.w[4] = r.w[5] = r.w[6] = r.w[7] = 0;
}

template <class BB>
__device__ __forceinline__ void rec_capture(Rec& r, int nsq, BB capt, bool promo, bool root_king, int cur,
                                            const Frame* stack, int stride) {
  const uint64_t c64 = (uint64_t)capt;
  r.w[0] = (uint32_t)c64;
  r.w[1] = (uint32_t)(c64 >> 32);
  const uint32_t flags = DRG_MF_CAPTURE | (promo ? DRG_MF_PROMO : 0u) | (root_king ? DRG_MF_KINGMOVE : 0u);
  r.w[2] = (uint32_t)nsq | (flags << 8);
  r.w[3] = r.w[4] = r.w[5] = r.w[6] = r.w[7] = 0;
#pragma unroll
  for (int i = 0; i < DRG_MAX_PATH; i++) {
    if (i < nsq) {
      const uint32_t sq = (i == nsq - 1) ? (uint32_t)cur : (uint32_t)stack[i * stride].sq;
      r.w[(10 + i) >> 2] |= sq << (8 * ((10 + i) & 3));
    }
  }
}

__device__ __forceinline__ void rec_store(DrgMove* dst, const Rec& r) {
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
  d[1] = make_uint4(r.w[4], r.w[5], r.w[6], r.w[7]);
}

__device__ __forceinline__ int rec_from(const Rec& r) { return (r.w[2] >> 16) & 0xFF; }
__device__ __forceinline__ int rec_nsq(const Rec& r) { return r.w[2] & 0xFF; }
__device__ __forceinline__ int rec_to(const Rec& r) {
  const int b = 10 + rec_nsq(r) - 1;
  uint32_t v = 0;
#pragma unroll
  for (int k = 2; k < 8; k++)
    if ((b >> 2) == k) v = r.w[k];
  return (v >> (8 * (b & 3))) & 0xFF;
}

struct KeepSink {  // PASS 2 of the tree search: the leaves of the running best, into the board's pool slot
  DrgMove* out;
  uint32_t n, room;
  __device__ __forceinline__ void restart() { n = 0; }
  __device__ __forceinline__ void quiet(int, int, bool) {}
  __device__ __forceinline__ void jump(int, int, int) {}
  template <class BB>
  __device__ __forceinline__ void capture(int nsq, BB capt, bool promo, bool root_king, int cur, const Frame* stack,
                                          int stride) {
    if (n < room) { Rec r; rec_capture(r, nsq, capt, promo, root_king, cur, stack, stride); rec_store(out + n, r); }
    n++;
  }
};

}  // namespace drg
