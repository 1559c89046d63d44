// write_bw.cu -- how fast can a kernel WRITE HBM on this GPU?  (the legal-move-mask export is a pure writer)
// variants: plain 16-byte stores, streaming (.cs) stores, 32-byte (v8.b32 / 256-bit) stores, TMA bulk stores from
// shared memory.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void w_plain(uint4* p, size_t n16) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = z;
}
__global__ void w_cs(uint4* p, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p + i), "r"(0u) : "memory");
}
__global__ void w_v8(uint4* p, size_t n16) {  // 256-bit stores
  const size_t n32 = n16 / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n32; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + 2 * i), "r"(0u) : "memory");
}
// contiguous chunk per warp (like k_emit's zero fill: each warp owns 80 000 B)
__global__ void w_warpchunk(uint4* p, size_t n16, int chunk16) {
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t c = warp; c * chunk16 < n16; c += nwarps) {
    uint4* q = p + c * chunk16;
    for (int i = lane; i < chunk16 && c * chunk16 + i < n16; i += 32) q[i] = z;
  }
}
// TMA bulk store: one elected thread per block streams a zeroed shared-memory tile to global memory
__global__ void w_bulk(uint8_t* p, size_t bytes, int tile) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < tile / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    int inflight = 0;
    for (size_t off = (size_t)blockIdx.x * tile; off + tile <= bytes; off += (size_t)gridDim.x * tile) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(s), "r"(tile) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const size_t bytes = (size_t)2621440000;  // 1 Mi Standard mask rows
  uint8_t* d;
  CK(cudaMalloc(&d, bytes));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const size_t n16 = bytes / 16;
  auto report = [&](const char* name, float best) { printf("%-34s %8.1f GB/s  (%.3f ms)\n", name, bytes / (best * 1e-3) / 1e9, best); };
#define RUN(name, launch)                                                    \
  {                                                                          \
    float best = 1e9;                                                        \
    for (int r = 0; r < 6; r++) {                                            \
      cudaEventRecord(a); launch; cudaEventRecord(b); cudaEventSynchronize(b); \
      float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms; \
    }                                                                        \
    CK(cudaGetLastError());                                                  \
    report(name, best);                                                      \
  }
  RUN("cudaMemsetAsync", cudaMemsetAsync(d, 0, bytes));
  for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
    char nm[64];
    snprintf(nm, 64, "plain st.v4  grid=%d x256", blocks); RUN(nm, (w_plain<<<blocks, 256>>>((uint4*)d, n16)));
    snprintf(nm, 64, "st.cs.v4     grid=%d x256", blocks); RUN(nm, (w_cs<<<blocks, 256>>>((uint4*)d, n16)));
    snprintf(nm, 64, "st.v8.b32    grid=%d x256", blocks); RUN(nm, (w_v8<<<blocks, 256>>>((uint4*)d, n16)));
    snprintf(nm, 64, "warp chunks 80000B grid=%d x128", blocks * 2); RUN(nm, (w_warpchunk<<<blocks * 2, 128>>>((uint4*)d, n16, 5000)));
  }
  for (int tile : {8192, 16384, 32768, 65536}) {
    char nm[64];
    cudaFuncSetAttribute(w_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, tile);
    snprintf(nm, 64, "TMA bulk store tile=%d grid=%d", tile, 148 * 4); RUN(nm, (w_bulk<<<148 * 4, 128, tile>>>(d, bytes, tile)));
  }
  return 0;
}
