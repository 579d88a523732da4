// Pass 2 of the batched raw-DEFLATE decoder for sm_100a, second generation: LZ77 back-references,
// ONE WARP PER STREAM, driven by 1 KiB WINDOWS of the match-head bitmap.
//
// Replaces the copy half of the reference's decompress_length_distance and copy_from_before
// (src/decompress.cpp:157-187,388-398) exactly like lz_warp.cuh (same input: literals and 3-byte
// match descriptors in place, one bitmap bit per match head; all range / room checks were made by
// pass 1) — what changes is how the work is cut:
//
//   * A lane loads ONE bitmap word per window (32 lanes x 32 bits = 1024 output bytes); the head
//     count of the window (popc + warp reduction) picks the mode.
//   * SPARSE windows (<= LZW_SPARSE_MAX heads; none at all for stored data): matches are taken one
//     at a time, in order.  A maximum-length match is checked, 32 candidates at once, for being the
//     first of a RUN of back-to-back matches with the same distance (what a compressor makes of a
//     long repeat: byte p = byte p - D over the whole run), and the run is FILLED in one go:
//       - D in {1,2,4,8,16}: the 16-byte pattern is built once and stored with 16-byte vector
//         stores, 512 bytes per warp iteration, no loads at all (memset-like);
//       - D >= 512: 16 bytes per lane copied from p - D, iterations separated by __syncwarp;
//       - anything else: 4 bytes per lane from the period [p0 - D, p0), phase kept incrementally.
//   * DENSE windows (text): 128-byte chunks, one aligned word per lane.  DEFLATE's minimum match
//     length of 3 means a word holds at most two heads and its bytes fall into at most TWO
//     segments: bytes [0, nA) continue match A (the incoming one, or a head at byte 0), bytes
//     [s, 4) belong to the word's last head B.  Non-overlapping segments (distance >= length, 99.9 %
//     of text matches) are gathered as ONE unaligned word each; a chunk none of whose sources lies
//     inside the chunk itself (two thirds of text chunks) is done in a single shot without the
//     frontier machinery.  Overlapping matches and the ragged first / last chunk of a stream take
//     the per-byte walk of lz_warp.cuh.
#pragma once

#include "lz_warp.cuh"
#include "deflate_lane.cuh"

namespace sfb {

constexpr int LZW_THREADS = 256;
constexpr uint32_t LZW_SPARSE_MAX = 6;
constexpr uint32_t LZW_NONE = 0xffffffffu;
constexpr uint32_t LZW_PF_MIN = 1024u;   // sources nearer than this are not prefetched (lzw_run<.., PF>)

struct LzwView {
  uint8_t* base;          // view byte v lives at base[v]; base is 128-byte aligned, windows start at v % 1024 == 0
  const uint32_t* bits;   // bitmap word of view bytes [32 w, 32 w + 32) at bits[w]
  uint32_t q, end;        // the stream occupies view bytes [q, end)
  uint32_t lane;
};

__device__ __forceinline__ uint32_t lzw_popc(uint32_t v)
{
#ifdef SFB_CPU_EMU
  return static_cast<uint32_t>(__builtin_popcount(v));
#else
  return static_cast<uint32_t>(__popc(v));
#endif
}
// CG = loads go to the L2 (ld.global.cg), not through the L1: the variant that runs WHILE pass 1 is
// still writing other parts of dst and of the bitmap from other SMs (lz_window_queue_kernel) must
// never see a line its SM cached before those bytes were published.
template <bool CG>
__device__ __forceinline__ uint32_t lzw_ld32(const void* p)
{
#ifndef SFB_CPU_EMU
  if constexpr (CG) return __ldcg(reinterpret_cast<const uint32_t*>(p));
#endif
  return *reinterpret_cast<const uint32_t*>(p);
}
template <bool CG>
__device__ __forceinline__ uint32_t lzw_ld8(const uint8_t* p)
{
#ifndef SFB_CPU_EMU
  if constexpr (CG) return __ldcg(p);
#endif
  return *p;
}
template <bool CG>
__device__ __forceinline__ uint32_t lzw_ldw(const uint8_t* base, uint32_t v)  // v % 4 == 0
{
  return lzw_ld32<CG>(base + v);
}
__device__ __forceinline__ void lzw_st16(uint8_t* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w)  // p % 16 == 0
{
  uint4 v;
  v.x = x;
  v.y = y;
  v.z = z;
  v.w = w;
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void lzw_prefetch(const void* p)
{
#ifndef SFB_CPU_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// the 3-byte descriptor at view position h (any alignment; h + 2 is inside the stream)
template <bool CG>
__device__ __forceinline__ uint32_t lzw_desc(const uint8_t* base, uint32_t h)
{
  const uint32_t a = h & ~3u, r = h & 3u;
  const uint32_t lo = lzw_ldw<CG>(base, a);
  const uint32_t hi = r >= 2u ? lzw_ldw<CG>(base, a + 4u) : 0u;
  return lz_funnel(lo, hi, 8u * r) & 0xffffffu;
}
// n source bytes (1 <= n <= 4) starting at view position f, as the low bytes of a word; only aligned
// words that hold at least one of them are read
template <bool CG>
__device__ __forceinline__ uint32_t lzw_gather(const uint8_t* base, uint32_t f, uint32_t n)
{
  const uint32_t a = f & ~3u, r = f & 3u;
  const uint32_t lo = lzw_ldw<CG>(base, a);
  const uint32_t hi = r + n > 4u ? lzw_ldw<CG>(base, a + 4u) : 0u;
  return lz_funnel(lo, hi, 8u * r);
}

// ---------------------------------------------------------------------------------------------
// FILL: bytes [p0, p1) := the periodic continuation of the d bytes before p0 (byte p = byte p - d,
// front to back — what copy_from_before() produces for one match, or for a run of back-to-back
// matches with the same distance).  All bytes below p0 are final.  Warp-uniform arguments.

// 4 bytes per lane; reads only from the period [p0 - d, p0)
template <bool CG>
__device__ __forceinline__ void lzw_fill_periodic(const LzwView v, uint32_t p0, uint32_t p1, uint32_t d)
{
  uint8_t* const base = v.base;
  uint32_t a = (p0 & ~3u) + 4u * v.lane;
  if (a >= p1) return;
  // phase of byte a inside the period (a may lie up to 3 bytes before p0: add whole periods first)
  uint32_t m = (a + 4u * d - p0) % d;
  const uint32_t step = 128u % d;
  const uint32_t s0 = p0 - d;
  for (; a < p1; a += 128u) {
    const bool full = a >= p0 && a + 4u <= p1;
    uint32_t val;
    if (full && d >= 4u && m + 3u < d) {  // four consecutive sources
      val = lzw_gather<CG>(base, s0 + m, 4u);
      *reinterpret_cast<uint32_t*>(base + a) = val;
    } else {
      uint32_t mb = m;
#pragma unroll
      for (uint32_t b = 0; b < 4; ++b) {
        while (mb >= d) mb -= d;
        const uint32_t p = a + b;
        if (p >= p0 && p < p1) base[p] = static_cast<uint8_t>(lzw_ld8<CG>(base + s0 + mb));
        ++mb;
      }
    }
    m += step;
    if (m >= d) m -= d;
  }
}

// The same for fills of 48 bytes and more (runs): 4 bytes per lane; reads only from the period [p0 - d, p0).
// A short period (d < 16) is widened first: once the first k d - d bytes exist, the fill is just as
// periodic with k d >= 16, and a word then rarely straddles the end of the period; a word that does
// is put together from the end and the start of the period (two gathers) instead of byte by byte.
template <bool CG>
__device__ __forceinline__ void lzw_fill_periodic_long(const LzwView v, uint32_t p0, uint32_t p1, uint32_t d)
{
  uint8_t* const base = v.base;
  if (d < 16u) {
    const uint32_t d2 = ((15u + d) / d) * d;   // 16 .. 30
    const uint32_t pre = d2 - d;               // < 32: one step
    if (v.lane < pre) base[p0 + v.lane] = static_cast<uint8_t>(lzw_ld8<CG>(base + p0 - d + v.lane % d));
    __syncwarp();
    p0 += pre;
    d = d2;
  }
  uint32_t a = (p0 & ~3u) + 4u * v.lane;
  if (a >= p1) return;
  // phase of byte a inside the period (a may lie up to 3 bytes before p0: add whole periods first)
  uint32_t m = (a + 4u * d - p0) % d;
  const uint32_t step = 128u % d;
  const uint32_t s0 = p0 - d;
  for (; a < p1; a += 128u) {
    const bool full = a >= p0 && a + 4u <= p1;
    if (full && d >= 4u) {
      const uint32_t rem = d - m;              // bytes up to the end of the period (>= 1)
      uint32_t val = lzw_gather<CG>(base, s0 + m, rem < 4u ? rem : 4u);
      if (rem < 4u) {
        const uint32_t g2 = lzw_gather<CG>(base, s0, 4u - rem);
        val = (val & ~(0xffffffffu << (8u * rem))) | (g2 << (8u * rem));
      }
      *reinterpret_cast<uint32_t*>(base + a) = val;
    } else {
      uint32_t mb = m;
#pragma unroll
      for (uint32_t b = 0; b < 4; ++b) {
        while (mb >= d) mb -= d;
        const uint32_t p = a + b;
        if (p >= p0 && p < p1) base[p] = static_cast<uint8_t>(lzw_ld8<CG>(base + s0 + mb));
        ++mb;
      }
    }
    m += step;
    if (m >= d) m -= d;
  }
}

// d in {1, 2, 4, 8, 16}: the 16 bytes of every aligned 16-byte slot are the same — build them once
template <bool CG>
__device__ __forceinline__ void lzw_fill_pattern16(const LzwView v, uint32_t p0, uint32_t p1, uint32_t d)
{
  constexpr unsigned FULL = 0xffffffffu;
  uint8_t* const base = v.base;
  const uint32_t r = p0 & 15u;
  // slot byte t (position (p0 & ~15) + t) = period byte ((t - r) mod d); lane k < 4 assembles word k
  uint32_t w = 0;
  {
    const uint32_t k = v.lane & 3u;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
      const uint32_t t = 4u * k + j;
      w |= lzw_ld8<CG>(base + p0 - d + ((t - r) & (d - 1u))) << (8u * j);
    }
  }
  const uint32_t r0 = __shfl_sync(FULL, w, 0), r1 = __shfl_sync(FULL, w, 1), r2 = __shfl_sync(FULL, w, 2),
                 r3 = __shfl_sync(FULL, w, 3);
  const uint32_t h16 = (p0 + 15u) & ~15u;   // first full slot
  const uint32_t t16 = p1 & ~15u;           // end of the last full slot
  // ragged head [p0, h16) and tail [t16, p1): lane t < 16 owns slot byte t
  if (v.lane < 16u) {
    const uint32_t t = v.lane;
    const uint32_t word = (t & 8u) ? ((t & 4u) ? r3 : r2) : ((t & 4u) ? r1 : r0);
    const uint8_t byte = static_cast<uint8_t>(word >> (8u * (t & 3u)));
    const uint32_t ph = (p0 & ~15u) + t;
    if (ph >= p0 && ph < p1 && ph < h16) base[ph] = byte;
    const uint32_t pt = t16 + t;
    if (t16 >= h16 && pt < p1) base[pt] = byte;
  }
  for (uint32_t a = h16 + 16u * v.lane; a + 16u <= t16 + 0u && a < t16; a += 512u) lzw_st16(base + a, r0, r1, r2, r3);
}

// d >= 512: 16 bytes per lane from p - d; a 512-byte iteration only reads what earlier iterations
// (or earlier matches) wrote
template <bool CG>
__device__ __forceinline__ void lzw_fill_far(const LzwView v, uint32_t p0, uint32_t p1, uint32_t d)
{
  uint8_t* const base = v.base;
  const uint32_t h16 = (p0 + 15u) & ~15u, t16 = p1 & ~15u;
  if (h16 >= t16) {
    lzw_fill_periodic<CG>(v, p0, p1, d);
    return;
  }
  if (p0 < h16) lzw_fill_periodic<CG>(v, p0, h16, d);  // (reads [p0 - d, p0) only)
  __syncwarp();
  for (uint32_t a0 = h16; a0 < t16; a0 += 512u) {
    const uint32_t a = a0 + 16u * v.lane;
    if (a < t16) {
      const uint32_t s = a - d, sa = s & ~3u, sh = 8u * (s & 3u);
      const uint32_t w0 = lzw_ldw<CG>(base, sa), w1 = lzw_ldw<CG>(base, sa + 4u), w2 = lzw_ldw<CG>(base, sa + 8u),
                     w3 = lzw_ldw<CG>(base, sa + 12u);
      const uint32_t w4 = sh ? lzw_ldw<CG>(base, sa + 16u) : 0u;
      lzw_st16(base + a, lz_funnel(w0, w1, sh), lz_funnel(w1, w2, sh), lz_funnel(w2, w3, sh), lz_funnel(w3, w4, sh));
    }
    __syncwarp();
  }
  if (t16 < p1) lzw_fill_periodic<CG>(v, t16, p1, d);
}

template <bool CG>
__device__ __noinline__ void lzw_fill(const LzwView v, uint32_t p0, uint32_t p1, uint32_t d)
{
  const uint32_t n = p1 - p0;
  if (n >= 48u && d <= 16u && (d & (d - 1u)) == 0u) lzw_fill_pattern16<CG>(v, p0, p1, d);
  else if (n >= 48u && d >= 512u) lzw_fill_far<CG>(v, p0, p1, d);
  else if (n >= 48u) lzw_fill_periodic_long<CG>(v, p0, p1, d);
  else lzw_fill_periodic<CG>(v, p0, p1, d);
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// One 128-byte chunk of a dense window.  `lo`/`hi`: only bytes in [lo, hi) are this call's to
// produce (stream edges; bytes a fill has already produced).  EDGE = the chunk is not wholly
// inside [lo, hi).  c_*: the most recent match seen so far (in/out).
template <bool EDGE, bool CG>
__device__ __forceinline__ void lzw_chunk(const LzwView& v, uint32_t P, uint32_t lo, uint32_t hi, uint32_t cw,
                                          uint32_t ncw, uint32_t hb4, uint32_t& c_o, uint32_t& c_end,
                                          uint32_t& c_d)
{
  constexpr unsigned FULL = 0xffffffffu;
  uint8_t* const base = v.base;
  const uint32_t lane = v.lane;
  const uint32_t wp = P + 4u * lane;
  uint32_t vm = 15u;
  if (EDGE) {
    const uint32_t l = lo > wp ? (lo - wp < 4u ? lo - wp : 4u) : 0u;
    const uint32_t h = hi > wp ? (hi - wp < 4u ? hi - wp : 4u) : 0u;
    vm = h > l ? ((1u << h) - 1u) & ~((1u << l) - 1u) : 0u;
  }
  const uint32_t hb = hb4 & vm;
  // the word after mine (lane 31: first word of the next chunk)
  const uint32_t nx = __shfl_sync(FULL, lane == 0u ? ncw : cw, static_cast<int>((lane + 1u) & 31u));
  const uint32_t hl = 31u - static_cast<uint32_t>(__clz(static_cast<int>(hb | 1u)));  // my last head (0: none or byte 0)
  const uint32_t dl = lz_funnel(cw, nx, 8u * hl) & 0xffffffu;                         // its descriptor
  const uint32_t own_pack = (4u * lane + hl) | (dl << 7);
  const uint32_t hm = __ballot_sync(FULL, hb != 0);
  const uint32_t below = hm & ((1u << lane) - 1u);
  const uint32_t sl = 31u - static_cast<uint32_t>(__clz(static_cast<int>(below | 1u)));
  const uint32_t in_pack = __shfl_sync(FULL, own_pack, static_cast<int>(sl));
  uint32_t in_o = c_o, in_end = c_end, in_d = c_d;
  if (below) {
    in_o = P + (in_pack & 127u);
    in_end = in_o + ((in_pack >> 7) & 255u) + 3u;
    in_d = (in_pack >> 15) + 1u;
  }
  if (hm) {  // (warp-uniform) the last head of this chunk is carried into the next ones
    const uint32_t top = 31u - static_cast<uint32_t>(__clz(static_cast<int>(hm)));
    const uint32_t pk = __shfl_sync(FULL, own_pack, static_cast<int>(top));
    c_o = P + (pk & 127u);
    c_end = c_o + ((pk >> 7) & 255u) + 3u;
    c_d = (pk >> 15) + 1u;
  }
  // my bytes: [0, nA) continue match A, [s, 4) belong to my last head B (length >= 3 covers them all)
  const bool has0 = (hb & 1u) != 0;
  const bool hasB = (hb & 0xeu) != 0;
  const uint32_t s = hasB ? hl : 4u;
  const uint32_t A_o = has0 ? wp : in_o;
  const uint32_t A_e = has0 ? wp + (cw & 255u) + 3u : in_end;
  const uint32_t A_d = has0 ? ((cw >> 8) & 0xffffu) + 1u : in_d;
  const uint32_t B_d = (dl >> 8) + 1u;
  uint32_t nA = A_e > wp ? A_e - wp : 0u;
  nA = nA < s ? nA : s;
  bool bytewise = EDGE;
  if (!EDGE) {
    const bool ovl = (nA != 0u && wp + nA - 1u - A_o >= A_d) || (hasB && 3u - s >= B_d);
    bytewise = __any_sync(FULL, ovl) != 0;
  }
  if (!bytewise) {
    // ---- segment mode: one unaligned word per segment ------------------------------------------
    // (all loads of a round are issued before the first one is used: one memory latency per round)
    bool pendA = nA != 0u, pendB = hasB;
    const uint32_t fA = wp - A_d, lastA = fA + nA - 1u;
    const uint32_t fB = wp + s - B_d, lastB = wp + 3u - B_d;
    const uint32_t mA = 0xffffffffu >> ((32u - 8u * nA) & 31u);  // (nA >= 1 where it is used)
    const uint32_t shB = (8u * s) & 31u;                         // (s < 4 where it is used)
    const uint32_t mB = 0xffffffffu << shB;
    const uint8_t* const pA = base + (fA & ~3u);
    const uint8_t* const pB = base + (fB & ~3u);
    const uint32_t rA = fA & 3u, rB = fB & 3u;
    const bool hiA = rA + nA > 4u, hiB = rB + 4u - s > 4u;       // the segment reaches into a second word
    uint32_t* const out = reinterpret_cast<uint32_t*>(base + wp);
    uint32_t res = cw;
    const bool dep = (pendA && lastA >= P) || (pendB && lastB >= P);
    if (!__any_sync(FULL, dep)) {  // every source lies before the chunk: one shot
      const uint32_t a0 = pendA ? lzw_ld32<CG>(pA) : 0u;
      const uint32_t a1 = (pendA && hiA) ? lzw_ld32<CG>(pA + 4) : 0u;
      const uint32_t b0 = pendB ? lzw_ld32<CG>(pB) : 0u;
      const uint32_t b1 = (pendB && hiB) ? lzw_ld32<CG>(pB + 4) : 0u;
      const uint32_t gA = lz_funnel(a0, a1, 8u * rA), gB = lz_funnel(b0, b1, 8u * rB) << shB;
      if (pendA) res = (res & ~mA) | (gA & mA);
      if (pendB) res = (res & ~mB) | (gB & mB);
      if (pendA | pendB) *out = res;
      __syncwarp();
      return;
    }
    for (;;) {
      // F = first byte of the chunk that is still unresolved
      const uint32_t first = pendA ? wp : (pendB ? wp + s : LZW_NONE);
      const uint32_t bal = __ballot_sync(FULL, first != LZW_NONE);
      if (bal == 0) break;
      const uint32_t F = __shfl_sync(FULL, first, __ffs(static_cast<int>(bal)) - 1);
      const bool doA = pendA && lastA < F, doB = pendB && lastB < F;
      const uint32_t a0 = doA ? lzw_ld32<CG>(pA) : 0u;
      const uint32_t a1 = (doA && hiA) ? lzw_ld32<CG>(pA + 4) : 0u;
      const uint32_t b0 = doB ? lzw_ld32<CG>(pB) : 0u;
      const uint32_t b1 = (doB && hiB) ? lzw_ld32<CG>(pB + 4) : 0u;
      const uint32_t gA = lz_funnel(a0, a1, 8u * rA), gB = lz_funnel(b0, b1, 8u * rB) << shB;
      if (doA) res = (res & ~mA) | (gA & mA);
      if (doB) res = (res & ~mB) | (gB & mB);
      pendA = pendA && !doA;
      pendB = pendB && !doB;
      if (doA | doB) *out = res;
      __syncwarp();
    }
    return;
  }
  // ---- per-byte mode (overlapping matches, ragged chunks): sources by the period rule ---------
  uint32_t src[4];
  uint32_t pend = 0;
#pragma unroll
  for (uint32_t b = 0; b < 4; ++b) {
    const bool useB = b >= s;
    const uint32_t o = useB ? wp + s : A_o;
    const uint32_t d = useB ? B_d : A_d;
    const bool cov = ((vm >> b) & 1u) && (useB || b < nA);
    uint32_t k = wp + b - o;
    if (cov && k >= d) k = lz_mod_small(k, d);  // overlapping copy: period d
    src[b] = o - d + k;
    if (cov) pend |= 1u << b;
  }
  uint32_t res = cw;
  for (;;) {
    const uint32_t pm_any = __ballot_sync(FULL, pend != 0);
    if (pm_any == 0) break;
    const int fl = __ffs(static_cast<int>(pm_any)) - 1;
    const uint32_t fp = __shfl_sync(FULL, pend, fl);
    const uint32_t F = P + 4u * static_cast<uint32_t>(fl) + static_cast<uint32_t>(__ffs(static_cast<int>(fp)) - 1);
#pragma unroll
    for (uint32_t b = 0; b < 4; ++b) {
      if (((pend >> b) & 1u) && src[b] < F) {
        const uint32_t byte = lzw_ld8<CG>(base + src[b]);
        res = lz_prmt(res, byte, 0x3210u ^ ((0x4u ^ b) << (4u * b)));  // byte b <- `byte`
        pend &= ~(1u << b);
      }
    }
    if (vm == 15u) {
      *reinterpret_cast<uint32_t*>(base + wp) = res;
    } else {  // a word this call does not own alone: only our bytes
#pragma unroll
      for (uint32_t b = 0; b < 4; ++b)
        if ((vm >> b) & 1u) base[wp + b] = static_cast<uint8_t>(res >> (8u * b));
    }
    __syncwarp();
  }
}

// this lane's bitmap word of window W, restricted to the stream's own bytes
template <bool CG>
__device__ __forceinline__ uint32_t lzw_window_bits(const LzwView& v, uint32_t W)
{
  const uint32_t pos0 = W + 32u * v.lane;
  if (pos0 >= v.end || pos0 + 32u <= v.q) return 0u;
  uint32_t w = lzw_ld32<CG>(v.bits + (pos0 >> 5));
  if (pos0 < v.q) w &= 0xffffffffu << (v.q - pos0);
  if (pos0 + 32u > v.end) w &= 0xffffffffu >> (pos0 + 32u - v.end);
  return w;
}

// What one warp carries along a stream, and what lz_window_queue_kernel hands from one segment of a
// stream to the next.
struct LzwState {
  uint32_t W;                  // the window being worked on (view position, multiple of 1024)
  uint32_t cur;                // every byte below cur is final (dense mode: see c_*)
  uint32_t c_o, c_end, c_d;    // dense mode: the most recent match, may reach past cur
};

// Windows W <= ... < w_stop of the stream in view `v` (v.end bounds what may be read at all; `look`
// bounds how far ahead a run of matches may be looked for — both the end of the stream when all of
// it is there).
// PF (EXPERIMENT, measured slower, off unless SFB200_LZW_PF=1): in dense windows every head lane
// asks the L2 for the line its match's SOURCE starts in one chunk before the chunk is worked on (the
// descriptors of a stream are all there before pass 2 starts, only the source bytes may not be
// final yet — a prefetch does not care).  Half of this kernel's stall samples on text are gathers
// that miss the L2, but the kernel is also issue-bound and moves 2.9 TB/s of DRAM
// traffic: C2 pass 2 10.3 ms with the prefetch against 8.5 ms without (profiles/ab/r02_ab_pf1.jsonl).
template <bool CG, bool PF = false>
__device__ __forceinline__ void lzw_run(const LzwView& v, LzwState& s, uint32_t w_stop, uint32_t look)
{
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = v.lane;
  const uint32_t end = v.end;
  uint32_t cur = s.cur, c_o = s.c_o, c_end = s.c_end, c_d = s.c_d;
  uint32_t pre_P = LZW_NONE, pre_w = 0, pre_w1 = 0;  // the words of chunks pre_P and pre_P + 128, loaded ahead
                                                     // (valid while nothing was written there)
  uint32_t W = s.W;
  uint32_t wbits = (W < end && W < w_stop) ? lzw_window_bits<CG>(v, W) : 0u;
  while (W < end && W < w_stop) {
    if (cur >= W + 1024u) {  // a run went past this window
      W = cur & ~1023u;
      if (W >= end || W >= w_stop) break;
      wbits = lzw_window_bits<CG>(v, W);
      continue;
    }
    const uint32_t pos0 = W + 32u * lane;
    uint32_t mine = wbits;  // heads at or after cur
    if (cur > pos0) mine = cur - pos0 >= 32u ? 0u : mine & (0xffffffffu << (cur - pos0));
    uint32_t cnt = lzw_popc(mine);
    cnt = __reduce_add_sync(FULL, cnt);
    const uint32_t wend = W + 1024u < end ? W + 1024u : end;
    if (c_end > cur && (cnt <= LZW_SPARSE_MAX || c_end - cur >= 128u)) {
      // the match carried out of a dense chunk: finish it in one go
      const uint32_t e = c_end;
      lzw_fill<CG>(v, cur, e, c_d);
      cur = e;
      c_end = 0;
      pre_P = LZW_NONE;
      continue;
    }
    if (cnt <= LZW_SPARSE_MAX && c_end <= cur) {
      // ---- sparse window: one match (or run of matches) at a time ---------------------------
      if (cnt) pre_P = LZW_NONE;
      while (cnt) {
        uint32_t m2 = wbits;
        if (cur > pos0) m2 = cur - pos0 >= 32u ? 0u : m2 & (0xffffffffu << (cur - pos0));
        const uint32_t bal = __ballot_sync(FULL, m2 != 0);
        if (bal == 0) break;
        const int fl = __ffs(static_cast<int>(bal)) - 1;
        const uint32_t fm = __shfl_sync(FULL, m2, fl);
        const uint32_t h = W + 32u * static_cast<uint32_t>(fl) + static_cast<uint32_t>(__ffs(static_cast<int>(fm)) - 1);
        const uint32_t desc = lzw_desc<CG>(v.base, h);
        const uint32_t L = (desc & 255u) + 3u, D = (desc >> 8) + 1u;
        uint32_t run_end = h + L;
        if (L == 258u) {
          // is this the first of a run of back-to-back matches with the same distance?  Lane i
          // looks at h + i L (lane 0 at the head itself).
          const uint32_t g = h + lane * L;
          bool okD = false, ok = false;
          uint32_t Lg = 0;
          if (g + 3u <= look && ((lzw_ld32<CG>(v.bits + (g >> 5)) >> (g & 31u)) & 1u)) {
            const uint32_t dg = lzw_desc<CG>(v.base, g);
            Lg = (dg & 255u) + 3u;
            okD = (dg >> 8) + 1u == D;
            ok = okD && Lg == L;
          }
          const uint32_t okm = __ballot_sync(FULL, ok), okDm = __ballot_sync(FULL, okD);
          const uint32_t f = okm == FULL ? 32u : static_cast<uint32_t>(__ffs(static_cast<int>(~okm)) - 1);  // >= 1
          run_end = h + f * L;
          const uint32_t Lf = __shfl_sync(FULL, Lg, static_cast<int>(f & 31u));
          if (f < 32u && ((okDm >> f) & 1u)) run_end += Lf;  // a shorter last one
        }
        __syncwarp();  // every lane has read the descriptor(s) before the fill overwrites them
        lzw_fill<CG>(v, h, run_end, D);
        cur = run_end;
        if (cur >= wend) break;
      }
      if (cur < wend) cur = wend;
      if (cur >= W + 1024u || cur >= end) {
        W += 1024u;
        if (W < end && W < w_stop) wbits = lzw_window_bits<CG>(v, W);
      }
      continue;
    }
    // ---- dense window: 128-byte chunks --------------------------------------------------------
    // (chunk words are loaded TWO chunks ahead and the next window's bitmap word one window ahead:
    //  a chunk's own loads never wait for memory; chunk P only writes inside [P, P + 128))
    const uint32_t nwb = lzw_window_bits<CG>(v, W + 1024u);
    if (W + 2048u < end) lzw_prefetch(v.base + W + 2048u + 32u * lane);
    {
      uint32_t P = cur & ~127u;
      const uint32_t lo = cur;  // (bytes below are final or not ours)
      auto load_w = [&](uint32_t PP) -> uint32_t {
        const uint32_t wp = PP + 4u * lane;
        return (wp + 4u > v.q && wp < end) ? lzw_ldw<CG>(v.base, wp) : 0u;
      };
      uint32_t cw, ncw;
      if (pre_P == P) {
        cw = pre_w;
        ncw = pre_w1;
      } else {
        cw = load_w(P);
        ncw = load_w(P + 128u);
      }
      bool stopped = false;
      for (; P < wend; P += 128u) {
        if (c_end >= P + 128u && c_end > cur && P >= lo) {  // the carried match covers the whole chunk: fill it
          stopped = true;
          break;
        }
        const uint32_t nncw = load_w(P + 256u);
        const uint32_t mw = __shfl_sync(FULL, wbits, static_cast<int>(((P - W) >> 5) + (lane >> 3)));
        const uint32_t hb4 = (mw >> (4u * (lane & 7u))) & 15u;
        if constexpr (PF) {
          // the last head in my word of the NEXT chunk (its words are in ncw, the bitmap bits in
          // this window's word or the next one's): far sources only, near ones are in the L1 / L2
          const uint32_t P1 = P + 128u;
          const uint32_t mw1 = __shfl_sync(FULL, P1 < W + 1024u ? wbits : nwb,
                                           static_cast<int>((((P1 - W) >> 5) & 31u) + (lane >> 3)));
          const uint32_t hb1 = (mw1 >> (4u * (lane & 7u))) & 15u;
          const uint32_t nx1 = __shfl_sync(FULL, lane == 0u ? nncw : ncw, static_cast<int>((lane + 1u) & 31u));
          const uint32_t hl1 = 31u - static_cast<uint32_t>(__clz(static_cast<int>(hb1 | 1u)));
          const uint32_t d1 = (lz_funnel(ncw, nx1, 8u * hl1) >> 8) & 0xffffu;   // distance - 1
          const uint32_t h1 = P1 + 4u * lane + hl1;
          // (a word that is not a descriptor any more — a fill went over it — gives some d1: the
          //  source is only asked for when it lies inside the stream)
          if (hb1 != 0u && d1 >= LZW_PF_MIN && d1 < h1 - v.q && h1 < end) lzw_prefetch(v.base + h1 - d1 - 1u);
        }
        if (P >= lo && P + 128u <= end) lzw_chunk<false, CG>(v, P, lo, end, cw, ncw, hb4, c_o, c_end, c_d);
        else lzw_chunk<true, CG>(v, P, lo, end, cw, ncw, hb4, c_o, c_end, c_d);
        cw = ncw;
        ncw = nncw;
        cur = P + 128u < end ? P + 128u : end;
      }
      pre_P = P;
      pre_w = cw;
      pre_w1 = ncw;
      if (stopped) continue;  // (the fill branch above takes it from here)
    }
    W += 1024u;
    wbits = nwb;
  }
  s.W = W;
  s.cur = cur;
  s.c_o = c_o;
  s.c_end = c_end;
  s.c_d = c_d;
}

// MINB: resident CTAs per SM the register allocation aims at (capi.cu picks the instantiation)
template <int MINB, bool PF = false>
__global__ void __launch_bounds__(LZW_THREADS, MINB) lz_window_kernel(const ResolveArgs a)
{
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u;
  for (;;) {
    unsigned long long si = 0;
    if (lane == 0) si = atomicAdd(a.stream_counter, 1ull);
    si = __shfl_sync(FULL, si, 0);
    if (si >= a.n) break;
    si = a.todo_list ? a.todo_list[si] : a.idx_base + si;
    const uint64_t off = a.dst_off[si] + a.dst_delta;
    const uint64_t wr = a.written[si];
    if (wr == 0) continue;
    LzwView v;
    v.base = a.dst_base + (off & ~1023ull);
    v.bits = a.match_bits + ((off & ~1023ull) >> 5);
    v.q = static_cast<uint32_t>(off & 1023u);
    v.end = v.q + static_cast<uint32_t>(wr);
    v.lane = lane;
    LzwState s;
    s.W = 0;
    s.cur = v.q;
    s.c_o = 0;
    s.c_end = 0;
    s.c_d = 1;
    lzw_run<false, PF>(v, s, 0xffffffffu, v.end);
  }
}

// ---------------------------------------------------------------------------------------------
// Pass 2 BESIDE pass 1: the same work, taken from the queue pass 1 fills (deflate_lane.cuh:
// QueueArgs) — item (stream, k) = "the first (k + 1) SEG + LAG bytes of this stream's pass-1 output
// are in memory", plus one final item per stream.  A warp pops an item, waits until the stream's
// earlier segments are done (they were popped before it, by warps that wait for nothing but their
// own predecessors), takes the five words of state the previous segment left, resolves the windows
// below (k + 1) SEG — never reading a byte at or above (k + 1) SEG + LAG - 16, never writing one at
// or above (k + 1) SEG + 258 — and leaves its state for the next.  All loads go to the L2 (CG): an
// L1 line may have been filled before a neighbouring stream's bytes in it were published.
// Launched twice per batch: a few CTAs per SM on a second stream while pass 1 runs (whatever fits
// beside its CTAs), and the rest of the GPU's worth behind pass 1 on its own stream; both drain the
// one queue and leave when pass 1 has said how many items there are and all have been taken.
#ifndef SFB_CPU_EMU
__device__ __forceinline__ unsigned long long lzq_ld64(const unsigned long long* p)
{
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ unsigned int lzq_ld32(const unsigned int* p)
{
  return *reinterpret_cast<const volatile unsigned int*>(p);
}

// Every wait of a consumer is bounded (Q_SPIN_CYCLES): a consumer that gives up records what it was
// waiting for in q.dbg and raises the abort flag, which ends every other consumer too — the batch
// call then fails (capi.cu checks the flag) instead of hanging the GPU.
constexpr long long Q_SPIN_CYCLES = 3000000000ll;   // ~1.5 s at 1.9 GHz (a turn: waits for other consumers only)
constexpr long long Q_IDLE_CYCLES = 8000000ll;      // ~4 ms (an item: the first ones take pass 1 ~2 ms)
__device__ __forceinline__ void lzq_give_up(const QueueArgs& q, unsigned long long what, unsigned long long a0,
                                            unsigned long long a1, unsigned long long a2)
{
  if (atomicCAS(q.dbg, 0ull, what) == 0ull) {
    q.dbg[1] = a0;
    q.dbg[2] = a1;
    q.dbg[3] = a2;
    __threadfence();
  }
}

template <int MINB>
__global__ void __launch_bounds__(LZW_THREADS, MINB) lz_window_queue_kernel(const ResolveArgs a, const QueueArgs q)
{
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u;
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) {
      // Nothing is claimed before pass 1 has been seen alive (tail != 0): the hardware may have placed
      // this kernel first and have no room left for pass 1, and waiting for it then would be a
      // deadlock — a warp that has waited Q_IDLE_CYCLES for the first sign leaves (whatever is left
      // over is taken by the launch behind pass 1).  Once pass 1 runs it needs nothing from here to
      // finish, so a claimed slot is either filled or known to stay empty (done).
      long long t0 = clock64();
      bool alive = true;
      while (lzq_ld64(q.tail) == 0ull && lzq_ld32(q.done + 1) == 0u) {
        if (clock64() - t0 > Q_IDLE_CYCLES || lzq_ld64(q.dbg) != 0ull) {
          alive = false;
          break;
        }
        __nanosleep(1000);
      }
      if (alive) {
        const unsigned long long i = atomicAdd(q.head, 1ull);
        t0 = clock64();
        for (; i < q.cap;) {   // (a slot past the end: more consumers than items — nothing will come)
          item = lzq_ld64(q.items + i);
          if (item) break;
          if (lzq_ld32(q.done + 1) != 0u && i >= lzq_ld64(q.final_tail)) {
            item = lzq_ld64(q.items + i);   // (nothing is pushed after `done`)
            break;
          }
          if (lzq_ld64(q.dbg) != 0ull) break;
          if (clock64() - t0 > Q_SPIN_CYCLES) {
            lzq_give_up(q, 1ull, i, lzq_ld64(q.tail), lzq_ld32(q.done));
            break;
          }
          __nanosleep(200);
        }
      }
    }
    item = __shfl_sync(FULL, item, 0);
    if (item == 0) break;
    const uint64_t si = (item >> 24) - 1ull;
    const uint32_t k = static_cast<uint32_t>(item >> 1) & 0x7fffffu;
    const bool fin = (item & 1ull) != 0;
    bool bail = false;
    if (lane == 0) {
      const long long t0 = clock64();
      while (lzq_ld32(q.turn + si) != k) {
        if (lzq_ld64(q.dbg) != 0ull) { bail = true; break; }
        if (clock64() - t0 > Q_SPIN_CYCLES) {
          lzq_give_up(q, 2ull, si, k, lzq_ld32(q.turn + si));
          bail = true;
          break;
        }
        __nanosleep(100);
      }
    }
    if (__shfl_sync(FULL, bail, 0)) break;
    __threadfence();   // (what the producers of the item and of the turn stored is visible from here on)
    const uint64_t off = a.dst_off[si] + a.dst_delta;
    LzwView v;
    v.base = a.dst_base + (off & ~1023ull);
    v.bits = a.match_bits + ((off & ~1023ull) >> 5);
    v.q = static_cast<uint32_t>(off & 1023u);
    v.lane = lane;
    LzwState s;
    if (k == 0) {
      s.W = 0;
      s.cur = v.q;
      s.c_o = 0;
      s.c_end = 0;
      s.c_d = 1;
    } else {
      const uint32_t* st = q.state + 5ull * si;
      s.W = __ldcg(st + 0);
      s.cur = __ldcg(st + 1);
      s.c_o = __ldcg(st + 2);
      s.c_end = __ldcg(st + 3);
      s.c_d = __ldcg(st + 4);
    }
    if (fin) {
      const uint64_t wr = __ldcg(a.written + si);
      v.end = v.q + static_cast<uint32_t>(wr);
      if (wr) lzw_run<true>(v, s, 0xffffffffu, v.end);
    } else {
      const uint32_t lim = v.q + ((k + 1u) << q.seg_shift);
      v.end = lim + q.lag - 16u;
      lzw_run<true>(v, s, lim & ~1023u, v.end);
      if (lane == 0) {
        uint32_t* st = q.state + 5ull * si;
        st[0] = s.W;
        st[1] = s.cur;
        st[2] = s.c_o;
        st[3] = s.c_end;
        st[4] = s.c_d;
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(q.turn + si) = k + 1u;
      }
    }
    __syncwarp();
  }
}
#endif

}  // namespace sfb
