// cuda_shim.h -- lets g++ compile the per-thread device algorithm of csrc/drg_core.cuh on the host.
// DEVELOPMENT / TEST AID ONLY (tools/devsim): the product library never includes this file and has no CPU
// path.  It exists because the authoring container has no GPU: new per-thread logic (tree-search splitting,
// statistics merging) is replayed here on one CPU thread and compared with the oracle before GPU time is spent.
#pragma once
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __align__(n) __attribute__((aligned(n)))

static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }

struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
