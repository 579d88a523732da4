// TEST INFRASTRUCTURE ONLY.
// Lets g++ compile starflate_b200/csrc/inflate_lanes.cuh so the kernel's lane-local logic can
// be unit-tested in the GPU-less authoring container (SURVEY.md §7 "hard parts": every kernel
// change otherwise costs a round-trip to the GPU box).  One emulated lane per "warp": every
// warp primitive degenerates to the identity, which is faithful because each lane decodes its
// own stream and never exchanges data with another lane.  This is never built into, loaded by,
// or reachable from the product library — it is not a fallback.
#pragma once
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__ inline __attribute__((noinline))
#define __constant__ static const
#define __shared__
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct emu_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emu_dim3 threadIdx, blockIdx;
static thread_local emu_dim3 blockDim, gridDim;

static inline unsigned __brev(unsigned v)
{
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_down_sync(unsigned, T, int) { return T{}; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline int __any_sync(unsigned, int p) { return p != 0; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __all_sync(unsigned, int p) { return p != 0; }
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
  const unsigned long long old = *p;
  *p = old + v;
  return old;
}
static inline unsigned atomicOr(unsigned* p, unsigned v)
{
  const unsigned old = *p;
  *p = old | v;
  return old;
}
