// emit_pattern.cu -- memory-system model of k_emit without any move generation: per warp, zero-fill 32 mask rows
// (80 000 B, coalesced 16-byte stores), then every lane stores 8 single bytes into its own row and 8 move records
// (32 B each) into a compact record array.  Variants isolate which store pattern costs what.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE_REC, bool BYTES, bool ZERO>
__global__ void __launch_bounds__(128) k(uint8_t* mask, uint8_t* rec, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t wb0 = ((int64_t)blockIdx.x * 128 + (threadIdx.x & ~31));
  const int64_t i = wb0 + lane;
  if (wb0 >= n) return;
  if (ZERO) {
    uint4* d4 = reinterpret_cast<uint4*>(mask + wb0 * 2500);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int q = lane; q < 5000; q += 32) d4[q] = z;
    __syncwarp();
  }
  uint32_t h = (uint32_t)i * 2654435761u;
#pragma unroll 1
  for (int k8 = 0; k8 < 8; k8++) {
    h = h * 1664525u + 1013904223u;
    if (BYTES) mask[i * 2500 + (h >> 8) % 2500] = 1;
    uint8_t* r = rec + (i * 8 + k8) * 32;
    if (MODE_REC == 1) {
      reinterpret_cast<uint4*>(r)[0] = make_uint4(h, 0, 2, 3);
      reinterpret_cast<uint4*>(r)[1] = make_uint4(0, 0, 0, 0);
    } else if (MODE_REC == 2) {
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%2,%2,%2,%2,%2,%2};" ::"l"(r), "r"(h), "r"(0u) : "memory");
    }
  }
}

int main() {
  const int64_t n = 1 << 20;
  uint8_t *mask, *rec;
  CK(cudaMalloc(&mask, n * 2500));
  CK(cudaMalloc(&rec, n * 8 * 32));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
#define RUN(name, ...)                                                       \
  {                                                                          \
    float best = 1e9;                                                        \
    for (int r = 0; r < 5; r++) {                                            \
      cudaEventRecord(a); __VA_ARGS__; cudaEventRecord(b); cudaEventSynchronize(b); \
      float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms; \
    }                                                                        \
    CK(cudaGetLastError());                                                  \
    printf("%-52s %.3f ms\n", name, best);                                  \
  }
  const int grid = (int)(n / 128);
  RUN("zero fill only", (k<0, false, true><<<grid, 128>>>(mask, rec, n)));
  RUN("zero fill + 8 byte stores", (k<0, true, true><<<grid, 128>>>(mask, rec, n)));
  RUN("zero fill + 8 records as 2x16B", (k<1, false, true><<<grid, 128>>>(mask, rec, n)));
  RUN("zero fill + 8 records as 1x32B (v8.b32)", (k<2, false, true><<<grid, 128>>>(mask, rec, n)));
  RUN("zero fill + bytes + records 2x16B", (k<1, true, true><<<grid, 128>>>(mask, rec, n)));
  RUN("zero fill + bytes + records 1x32B", (k<2, true, true><<<grid, 128>>>(mask, rec, n)));
  RUN("records 2x16B only", (k<1, false, false><<<grid, 128>>>(mask, rec, n)));
  RUN("records 1x32B only", (k<2, false, false><<<grid, 128>>>(mask, rec, n)));
  RUN("bytes only (rows not in L2)", (k<0, true, false><<<grid, 128>>>(mask, rec, n)));
  RUN("cudaMemset rows", cudaMemsetAsync(mask, 0, n * 2500));
  return 0;
}
