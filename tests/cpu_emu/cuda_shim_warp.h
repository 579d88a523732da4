// TEST INFRASTRUCTURE ONLY.
// Lets g++ compile starflate_b200/csrc/lz_warp.cuh (a warp-cooperative kernel) for the host: the
// 32 lanes of one warp run as 32 host threads and every warp primitive (__shfl_sync,
// __any_sync, __all_sync, __syncwarp ...) is a rendezvous on a std::barrier, so the exchange
// pattern, the divergence structure and the memory ordering the kernel relies on
// (stores -> __syncwarp -> loads by other lanes) are exercised as written.  Never built into,
// loaded by, or reachable from the product library — it is not a fallback.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__ inline __attribute__((noinline))
#define __launch_bounds__(...)
#define __constant__ static const
#define __shared__
#define __align__(n) __attribute__((aligned(n)))

struct emu_dim3 { unsigned x = 0, y = 0, z = 0; };
struct uint4 { unsigned x, y, z, w; };
static thread_local emu_dim3 threadIdx, blockIdx;
static thread_local emu_dim3 blockDim, gridDim;

struct EmuWarp {
  std::barrier<> bar{32};
  uint64_t slot[32];
};
static EmuWarp* emu_warp = nullptr;

template <class T> static inline T emu_exchange(T v, unsigned src)
{
  const unsigned lane = threadIdx.x & 31u;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  emu_warp->slot[lane] = raw;
  emu_warp->bar.arrive_and_wait();
  raw = emu_warp->slot[src & 31u];
  emu_warp->bar.arrive_and_wait();
  T r;
  std::memcpy(&r, &raw, sizeof(T));
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, static_cast<unsigned>(src)); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d)
{
  const unsigned lane = threadIdx.x & 31u;
  const unsigned src = lane + static_cast<unsigned>(d);
  return emu_exchange(v, src > 31u ? lane : src);
}
static inline unsigned __ballot_sync(unsigned, int p)
{
  const unsigned lane = threadIdx.x & 31u;
  emu_warp->slot[lane] = p ? 1u : 0u;
  emu_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (unsigned i = 0; i < 32; ++i) m |= static_cast<unsigned>(emu_warp->slot[i]) << i;
  emu_warp->bar.arrive_and_wait();
  return m;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
  const unsigned lane = threadIdx.x & 31u;
  emu_warp->slot[lane] = v;
  emu_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (unsigned i = 0; i < 32; ++i) m = emu_warp->slot[i] > m ? static_cast<unsigned>(emu_warp->slot[i]) : m;
  emu_warp->bar.arrive_and_wait();
  return m;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
  const unsigned lane = threadIdx.x & 31u;
  emu_warp->slot[lane] = v;
  emu_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (unsigned i = 0; i < 32; ++i) m += static_cast<unsigned>(emu_warp->slot[i]);
  emu_warp->bar.arrive_and_wait();
  return m;
}
static inline unsigned __match_any_sync(unsigned, unsigned v)
{
  const unsigned lane = threadIdx.x & 31u;
  emu_warp->slot[lane] = v;
  emu_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (unsigned i = 0; i < 32; ++i) m |= (static_cast<unsigned>(emu_warp->slot[i]) == v ? 1u : 0u) << i;
  emu_warp->bar.arrive_and_wait();
  return m;
}
static inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0; }
static inline int __all_sync(unsigned mask, int p) { return __ballot_sync(mask, p) == 0xffffffffu; }
static inline void __syncwarp() { emu_warp->bar.arrive_and_wait(); }
static inline void __syncthreads() { emu_warp->bar.arrive_and_wait(); }  // (one warp per block here)
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
  const unsigned lane = threadIdx.x & 31u;
  return emu_exchange(v, lane >= d ? lane - d : lane);
}
static inline unsigned __brev(unsigned v)
{
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
}
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
  return static_cast<unsigned>(((static_cast<uint64_t>(hi) << 32) | lo) >> (s & 31u));
}
static inline unsigned atomicCAS(unsigned* p, unsigned expect, unsigned v)
{
  __atomic_compare_exchange_n(p, &expect, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
  return expect;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
