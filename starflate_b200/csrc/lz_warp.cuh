// Pass 2 of the batched raw-DEFLATE decoder for sm_100a: LZ77 back-references, ONE WARP PER
// STREAM, one output byte per lane, 32-byte chunks aligned with the match-head bitmap.
//
// Replaces the copy half of the reference's decompress_length_distance and copy_from_before
// (src/decompress.cpp:157-187,388-398): pass 1 (huff_lanes.cuh) has already placed literals and
// stored payload at their final positions and left, for every match, a 3-byte descriptor at
// the match start plus a bit in the match-head bitmap; this pass turns each descriptor into the
// bytes the reference would have copied.  All range/room checks were made in pass 1.
//
// Per 32-byte chunk [P, P+32) of a stream's output (P a multiple of 32 from dst_base):
//   1. one bitmap word M says which lanes sit on a match head; those lanes assemble their
//      descriptor from their own byte and the next two (the chunk after is prefetched, so a
//      head in lanes 30/31 finds its bytes there);
//   2. every lane finds the nearest head at or below it (clz on M below the lane), or falls
//      back to the match carried in from earlier chunks, and so learns whether it lies inside
//      a match and at which offset k;
//   3. a covered lane's byte is the byte at  start - distance + (k mod distance)  — the
//      forward-overlapping copy of copy_from_before() is periodic with period `distance`, so
//      the source always lies BEFORE the match start.  Sources below P are final in memory
//      (earlier chunks, ordered by __syncwarp) and are gathered; sources inside this chunk are
//      lower lanes, resolved by pointer doubling with shuffles (chains only pass through
//      different matches, so their depth is at most the number of matches in the chunk);
//   4. the whole chunk is stored back (one 32-byte sector per warp store).
// Streams are pulled from a global counter.  A stream's chunks are processed in order, so the
// 32 KiB window a match may reach into was written by this same warp a few thousand
// instructions earlier: with ~5k streams in flight the windows stay in the 126 MB L2.
#pragma once

#include <cstdint>
#ifndef SFB_CPU_EMU
#include <cuda_runtime.h>
#endif

namespace sfb {

struct ResolveArgs {
  uint8_t* dst_base;
  const uint64_t* dst_off;
  const uint64_t* written;
  const uint32_t* match_bits;  // bit k <-> dst_base[k]
  uint64_t n;
  unsigned long long* stream_counter;  // zeroed before launch
};

constexpr int LZ_THREADS = 256;

// byte permute (PRMT): result byte i = byte sel.nibble[i] of the 8 bytes b:a
__device__ __forceinline__ uint32_t lz_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#ifdef SFB_CPU_EMU
  const uint64_t v = (static_cast<uint64_t>(b) << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= static_cast<uint32_t>((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
  return r;
#else
  return __byte_perm(a, b, sel);
#endif
}

__global__ void __launch_bounds__(LZ_THREADS) lz_resolve_kernel(const ResolveArgs a)
{
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u;
  // loop invariants of the descriptor assembly: lane l takes bytes l+1 and l+2 of the chunk;
  // for lanes 31 / 30,31 they are bytes 0 / 0,1 of the NEXT chunk, which travel in byte 1 of
  // the shuffled value (v = cur | nxt << 8)
  const int src1 = static_cast<int>((lane + 1u) & 31u), src2 = static_cast<int>((lane + 2u) & 31u);
  const uint32_t sel1 = lane == 31u ? 0x7750u : 0x7740u;  // {cur.b0, t1.b0 | t1.b1, 0(b3 of cur: 0), ...}
  const uint32_t sel2 = lane >= 30u ? 0x7510u : 0x7410u;  // {d.b0, d.b1, t2.b0 | t2.b1, 0}
  const uint32_t le_mask = 0xffffffffu >> (31u - lane);   // lanes at or below this one
  for (;;) {
    unsigned long long si = 0;
    if (lane == 0) si = atomicAdd(a.stream_counter, 1ull);
    si = __shfl_sync(FULL, si, 0);
    if (si >= a.n) break;
    const uint64_t off = a.dst_off[si];
    const uint64_t wr = a.written[si];
    if (wr == 0) continue;
    // virtual positions: byte v of the view sits at base[v]; the stream occupies [q, end)
    uint8_t* const base = a.dst_base + (off & ~31ull);
    const uint32_t* bmw = a.match_bits + (off >> 5);
    const uint32_t q = static_cast<uint32_t>(off & 31u);
    const uint32_t end = q + static_cast<uint32_t>(wr);
    // the most recent match seen so far: [c_o, c_end) at distance c_d (none yet: empty range)
    uint32_t c_o = 0, c_end = 0, c_d = 1;
    // two chunks of pass-1 bytes and bitmap words are kept in flight ahead of the one being
    // resolved (nothing a chunk's resolution stores can change them: it only writes its own
    // 32 bytes)
    uint8_t* pb = base + lane;  // this lane's byte of the current chunk
    uint32_t cur = 0, nxt = 0, M, nM = 0;
    if (lane >= q && lane < end) cur = pb[0];
    if (lane + 32u < end) nxt = pb[32];
    M = bmw[0] & (0xffffffffu << q);
    if (32u < end) nM = bmw[1];
    for (uint32_t P = 0; P < end; P += 32, pb += 32, ++bmw) {
      const uint32_t p = P + lane;
      uint32_t nn = 0, nnM = 0;
      if (p + 64u < end) nn = pb[64];
      if (P + 64u < end) nnM = bmw[2];
      const uint32_t hi = end - P;  // >= 1: valid lanes are those below hi (and, in chunk 0, from q)
      const bool valid = lane < hi && p >= q;
      const uint32_t Mv = hi < 32u ? M & ((1u << hi) - 1u) : M;
      // descriptor of a match that starts at this lane (only meaningful on head lanes)
      const uint32_t v = cur | (nxt << 8);
      const uint32_t t1 = __shfl_sync(FULL, v, src1);
      const uint32_t t2 = __shfl_sync(FULL, v, src2);
      const uint32_t desc = lz_prmt(lz_prmt(cur, t1, sel1), t2, sel2);
      // the match this lane may lie in: nearest head at or below the lane, else the carry
      const uint32_t below = Mv & le_mask;
      const uint32_t hb = 31u - static_cast<uint32_t>(__clz(static_cast<int>(below | 1u)));
      const uint32_t hdesc = __shfl_sync(FULL, desc, static_cast<int>(hb));
      uint32_t t_o = c_o, t_end = c_end, t_d = c_d;
      if (below) {
        t_o = P + hb;
        t_end = t_o + (hdesc & 0xffu) + 3u;
        t_d = (hdesc >> 8) + 1u;
      }
      if (Mv) {  // warp-uniform: the last head of this chunk is carried into the next ones
        const uint32_t top = 31u - static_cast<uint32_t>(__clz(static_cast<int>(Mv)));
        const uint32_t tdesc = __shfl_sync(FULL, desc, static_cast<int>(top));
        c_o = P + top;
        c_end = c_o + (tdesc & 0xffu) + 3u;
        c_d = (tdesc >> 8) + 1u;
      }
      const bool covered = valid && p < t_end;
      uint32_t val = cur;
      uint32_t ptr = lane;
      if (covered) {
        uint32_t k = p - t_o;
        if (k >= t_d) k %= t_d;
        const uint32_t src = t_o - t_d + k;  // >= q: pass 1 checked distance <= written
        if (src >= P) ptr = src - P;         // a lower lane of this chunk
        else val = base[src];                // final since an earlier chunk
      }
      if (__any_sync(FULL, ptr != lane)) {
        for (;;) {
          const uint32_t pp = __shfl_sync(FULL, ptr, static_cast<int>(ptr));
          const bool fix = pp == ptr;
          ptr = pp;
          if (__all_sync(FULL, fix)) break;
        }
        val = __shfl_sync(FULL, val, static_cast<int>(ptr));
      }
      if (valid) pb[0] = static_cast<uint8_t>(val);
      __syncwarp();
      cur = nxt;
      nxt = nn;
      M = nM;
      nM = nnM;
    }
  }
}

// Position-weighted checksum (see starflate_b200.h): one warp per stream.
__global__ void checksum_kernel(const uint8_t* base, const uint64_t* off, const uint64_t* len,
                                uint64_t* out, uint64_t n)
{
  const uint64_t w = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (w >= n) return;
  const uint8_t* p = base + off[w];
  const uint64_t L = len[w];
  uint64_t acc = 0;
  for (uint64_t j = lane; j < L; j += 32)
    acc += (static_cast<uint64_t>(p[j]) + 1ull) * ((0x9E3779B97F4A7C15ull * (j + 1ull)) | 1ull);
  for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) out[w] = acc;
}

}  // namespace sfb
