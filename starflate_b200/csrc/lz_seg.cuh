// Pass 2 of the batched raw-DEFLATE decoder for sm_100a, second generation: LZ77 back-references
// resolved SEGMENT-wise.  One warp per stream, 128-byte chunks, one aligned 32-bit word per lane —
// the layout of lz_warp.cuh (whose header describes the in-place token format) — but a lane no
// longer walks its four bytes one by one.
//
// Replaces the copy half of the reference's decompress_length_distance and copy_from_before
// (src/decompress.cpp:157-187,388-398); all range / room checks were made in pass 1.
//
// A match is at least 3 bytes long, so the 4 bytes of a lane's word belong to at most three
// matches: the one that reaches into the word from the left (segment A: bytes [0, nA)), the one
// that starts at its first head (segment B) and — only when B is a 3-byte match at byte 0 — the
// one that starts at byte 3 (segment C).  Inside a segment the copy_from_before() source of lane
// byte j is  start - D + ((k0 + j - o) mod D)  (start: match start, D: distance, o: first lane
// byte of the segment, k0: how far into the match that byte is): consecutive positions as long
// as the period does not wrap.  So a segment is gathered as ONE unaligned word — the two aligned
// words around  a = start - D + (k0 mod D) - o  funnel-shifted, whose byte j is then exactly the
// source of lane byte j — and merged under the segment's byte mask: two loads, a funnel shift
// and one LOP3 per segment instead of a load, a compare and a byte permute per byte.  The rare
// shapes take side branches that are uniform across the warp in text: a period that wraps
// inside the word (one more unaligned word), distances 1..3 (the period's bytes replicated with
// one byte permute from a 16-entry selector table), k0 >= D (k0 mod D by a reciprocal table).
//
// Resolution order inside a chunk is the round scheme of lz_warp.cuh at segment granularity:
// F = first byte of the chunk that is still pending; a segment whose LAST source byte lies below
// F is gathered; the words are stored; __syncwarp; F advances.  The first pending segment is
// always ready (its sources precede its first byte), so every round makes progress.
//
// ncu on the first-generation kernel (profiles/r01_two_pass_c2_end.md) showed the ALU pipe 83 %
// busy (LOP3 / SHF / PRMT / ISETP issue at one warp instruction per two cycles per scheduler)
// and 21 % of the stall samples on the per-byte gather: this kernel is organised around fewer
// ALU-pipe instructions per chunk.
#pragma once

#include "lz_warp.cuh"

namespace sfb {

// floor(2^20 / d) + 1 for d < 264: q = (k * inv) >> 20 equals floor(k / d) for every k < 2^9
// (the excess k * (inv - 2^20 / d) / 2^20 is below 2^-11 < 1 / d).
struct LzInvTab {
  uint32_t v[264];
};
constexpr LzInvTab lz_make_inv()
{
  LzInvTab t{};
  t.v[0] = 0;
  for (uint32_t d = 1; d < 264; ++d) t.v[d] = (1u << 20) / d + 1u;
  return t;
}
__constant__ LzInvTab c_lz_inv = lz_make_inv();
// byte-permute selectors for distances 1..3, index 4 * D + m0: result byte i = period byte (m0 + i) mod D
__constant__ uint32_t c_lz_sel[16] = {0, 0, 0, 0, 0x0000u, 0, 0, 0, 0x1010u, 0x0101u, 0, 0, 0x0210u, 0x1021u, 0x2102u, 0};

// k mod d for d <= k < 512, d < 264
__device__ __forceinline__ uint32_t lz_mod_tab(uint32_t k, uint32_t d)
{
  return k - ((k * c_lz_inv.v[d]) >> 20) * d;
}

// The stream's output as the resolve pass sees it: byte v of the "view" (v = 0 at a 128-byte
// aligned address at or below the stream's first byte).
struct LzGlobalView {
  static constexpr bool GUARD_LOW = true;  // nothing readable below the first stream of a buffer
  uint8_t* base;
  __device__ __forceinline__ uint32_t ldw(uint32_t v) const { return *reinterpret_cast<const uint32_t*>(base + v); }
  __device__ __forceinline__ void stw(uint32_t v, uint32_t x) const { *reinterpret_cast<uint32_t*>(base + v) = x; }
  __device__ __forceinline__ void stb(uint32_t v, uint32_t x) const { base[v] = static_cast<uint8_t>(x); }
};

// bytes [a, a + 4) of the view.  `a` may be as low as -3 when the source sits in the first bytes of
// the view (only the bytes at non-negative positions are then used by the caller): a view whose
// memory does not extend below position 0 (M::GUARD_LOW) does not read the word at -4.
template <class M>
__device__ __forceinline__ uint32_t lz_uw(const M& mem, uint32_t a)
{
  const uint32_t aw = a & ~3u;
  uint32_t lo = 0;
  if (!M::GUARD_LOW || static_cast<int32_t>(aw) >= 0) lo = mem.ldw(aw);
  const uint32_t hi = mem.ldw(aw + 4u);
  return lz_funnel(lo, hi, a << 3);
}

// The bytes copy_from_before() would put into lane bytes [o, o + n) of the word at wp, for the
// match that starts at `start` with distance D, whose byte at lane byte o is m0 = k0 mod D bytes
// into its period.  Bytes outside [o, o + n) of the result are unspecified.
template <class M>
__device__ __forceinline__ uint32_t lz_gather(const M& mem, uint32_t start, uint32_t D, uint32_t o, uint32_t n, uint32_t m0)
{
  uint32_t v;
  if (D >= 4u) {
    const uint32_t a = start - D + m0 - o;
    v = lz_uw(mem, a);
    if (m0 + n > D) {  // the period wraps inside my bytes (once: n <= 4 <= D)
      const uint32_t v2 = lz_uw(mem, a - D);
      const uint32_t wm = 0xffffffffu << (8u * (o + D - m0));  // lane bytes from the wrap on
      v = (v & ~wm) | (v2 & wm);
    }
  } else {
    const uint32_t w = lz_uw(mem, start - D);  // its first D bytes are the period
    v = lz_prmt(w, 0u, c_lz_sel[4u * D + m0] << (4u * o));
  }
  return v;
}

// byte mask of lane bytes [o, o + n), 1 <= n, o + n <= 4
__device__ __forceinline__ uint32_t lz_bytes_mask(uint32_t o, uint32_t n)
{
  return (0xffffffffu >> (32u - 8u * n)) << (8u * o);
}

// One 128-byte chunk, one word per lane.  `cw` my raw word, `nx` the raw word after it, `hb4`
// the match heads among my bytes, `vm` my valid bytes (EDGE chunks: the first / last of a stream),
// (t_o, t_end, t_d) the match that may reach into my word from the left.
//
// Control flow is uniform across the warp wherever it matters: which lanes have which segments
// differs from lane to lane in every chunk, so the gathers are PREDICATED straight-line code (a
// divergent branch would run both sides anyway and pay for the reconvergence on top); only the
// rare shapes — a period that wraps inside a word, distances 1..3 — switch the whole warp to the
// general loop for that chunk.
template <bool EDGE, class M>
__device__ __forceinline__ void lz_chunk_segments(const M& mem, uint32_t wp, uint32_t cw, uint32_t nx,
                                                  uint32_t hb4, uint32_t hl, uint32_t d_last, uint32_t vm,
                                                  uint32_t t_o, uint32_t t_end, uint32_t t_d)
{
  constexpr unsigned FULL = 0xffffffffu;
  // ---- segment A: the match that reaches in from the left ------------------------------------
  const bool hasA = t_end > wp;
  const uint32_t nA = hasA ? min(t_end - wp, 4u) : 0u;
  const uint32_t k0A = wp - t_o;
  uint32_t m0A = k0A;
  if (__any_sync(FULL, hasA && k0A >= t_d)) {
    if (hasA && k0A >= t_d) m0A = lz_mod_tab(k0A, t_d);  // (then t_d <= k0A < 258 + 128)
  }
  const uint32_t eA = t_o - t_d + min(t_d, k0A + nA) - 1u;  // last source byte it needs
  // ---- segment B: my first head; segment C: a second one (B is then 3 bytes at byte 0) -------
  const bool hasB = hb4 != 0u;
  const uint32_t h1 = static_cast<uint32_t>(__ffs(static_cast<int>(hb4 | 16u))) - 1u;
  const uint32_t d1 = lz_funnel(cw, nx, 8u * h1) & 0xffffffu;
  const uint32_t DB = (d1 >> 8) + 1u;
  const uint32_t nB = min(4u - h1, (d1 & 255u) + 3u);
  const uint32_t eB = wp + h1 - DB + min(DB, nB) - 1u;
  const bool hasC = hasB && hl != h1;
  const uint32_t DC = (d_last >> 8) + 1u;
  const uint32_t nC = min(4u - hl, (d_last & 255u) + 3u);
  const uint32_t eC = wp + hl - DC + min(DC, nC) - 1u;
  const bool anyC = __any_sync(FULL, hasC);
  uint32_t pend = (hasA ? 1u : 0u) | (hasB ? 2u : 0u) | (hasC ? 4u : 0u);
  // a shape the plain gather (one unaligned word, no wrap) does not cover?
  const bool odd = (hasA && (t_d < 4u || m0A + nA > t_d)) || (hasB && DB < 4u) || (hasC && DC < 4u);
  uint32_t res = cw;
  auto store = [&](bool changed) {
    if (changed) {
      if (!EDGE || vm == 15u) {
        mem.stw(wp, res);
      } else {  // first / last word of the stream: only our bytes
#pragma unroll
        for (uint32_t b = 0; b < 4; ++b)
          if ((vm >> b) & 1u) mem.stb(wp + b, res >> (8u * b));
      }
    }
  };
  // first byte of this lane that is still pending
  auto first_pending = [&]() { return wp + ((pend & 1u) ? 0u : (pend & 2u) ? h1 : hl); };
  if (!__any_sync(FULL, odd)) {
    // ---- the plain loop: every segment is one unaligned word --------------------------------
    const uint32_t aA = t_o - t_d + m0A, aB = wp - DB, aC = wp - DC;   // (a = start - D + m0 - o)
    const uint32_t mA = 0xffffffffu >> ((32u - 8u * nA) & 31u);
    const uint32_t mB = lz_bytes_mask(h1 & 3u, nB | (hasB ? 0u : 1u));
    const uint32_t mC = lz_bytes_mask(hl, nC);
    for (;;) {
      const uint32_t pm_any = __ballot_sync(FULL, pend != 0u);
      if (pm_any == 0u) break;
      const uint32_t F = __shfl_sync(FULL, first_pending(), __ffs(static_cast<int>(pm_any)) - 1);
      const bool doA = (pend & 1u) && eA < F;
      const bool doB = (pend & 2u) && eB < F;
      if (doA) {
        const uint32_t v = lz_uw(mem, aA);
        res = (res & ~mA) | (v & mA);
      }
      if (doB) {
        const uint32_t v = lz_uw(mem, aB);
        res = (res & ~mB) | (v & mB);
      }
      bool doC = false;
      if (anyC) {
        doC = (pend & 4u) && eC < F;
        if (doC) {
          const uint32_t v = lz_uw(mem, aC);
          res = (res & ~mC) | (v & mC);
        }
      }
      pend &= ~((doA ? 1u : 0u) | (doB ? 2u : 0u) | (doC ? 4u : 0u));
      store(doA | doB | doC);
      __syncwarp();
    }
    return;
  }
  // ---- the general loop ------------------------------------------------------------------------
  for (;;) {
    const uint32_t pm_any = __ballot_sync(FULL, pend != 0u);
    if (pm_any == 0u) break;
    const uint32_t F = __shfl_sync(FULL, first_pending(), __ffs(static_cast<int>(pm_any)) - 1);
    bool changed = false;
    if ((pend & 1u) && eA < F) {
      const uint32_t v = lz_gather(mem, t_o, t_d, 0u, nA, m0A);
      const uint32_t m = 0xffffffffu >> (32u - 8u * nA);
      res = (res & ~m) | (v & m);
      pend &= ~1u;
      changed = true;
    }
    if ((pend & 2u) && eB < F) {
      const uint32_t v = lz_gather(mem, wp + h1, DB, h1, nB, 0u);
      const uint32_t m = lz_bytes_mask(h1, nB);
      res = (res & ~m) | (v & m);
      pend &= ~2u;
      changed = true;
    }
    if ((pend & 4u) && eC < F) {
      const uint32_t v = lz_gather(mem, wp + hl, DC, hl, nC, 0u);
      const uint32_t m = lz_bytes_mask(hl, nC);
      res = (res & ~m) | (v & m);
      pend &= ~4u;
      changed = true;
    }
    store(changed);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(LZ_THREADS, 5) lz_resolve_kernel(const ResolveArgs a)
{
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;             // lanes below this one
  const int next_lane = static_cast<int>((lane + 1u) & 31u);
  const uint32_t bit_sh = 4u * (lane & 7u);               // my 4 bits inside my bitmap word
  for (;;) {
    unsigned long long si = 0;
    if (lane == 0) si = atomicAdd(a.stream_counter, 1ull);
    si = __shfl_sync(FULL, si, 0);
    if (si >= a.n) break;
    si = a.todo_list ? a.todo_list[si] : a.idx_base + si;
    const uint64_t off = a.dst_off[si] + a.dst_delta;
    const uint64_t wr = a.written[si];
    if (wr == 0) continue;
    // view positions: byte v sits at base[v]; the stream occupies [q, end)
    LzGlobalView mem{a.dst_base + (off & ~127ull)};
    const uint32_t q = static_cast<uint32_t>(off & 127u);
    const uint32_t end = q + static_cast<uint32_t>(wr);
    // this lane's word and bitmap word of the current chunk, and of the two after it
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(mem.base) + lane;
    const uint32_t* pm = a.match_bits + ((off & ~127ull) >> 5) + (lane >> 3);
    // a word is read iff one of its bytes belongs to the stream (never a word wholly outside)
    auto load_w = [&](uint32_t P) -> uint32_t {
      const uint32_t wp = P + 4u * lane;
      return (wp + 4u > q && wp < end) ? pw[P >> 2] : 0u;
    };
    auto load_m = [&](uint32_t P) -> uint32_t { return P < end ? pm[P >> 5] : 0u; };
    uint32_t cw = load_w(0), ncw = load_w(128), mw = load_m(0), nmw = load_m(128);
    // the most recent match seen so far: [c_o, c_end) at distance c_d (none yet: empty range)
    uint32_t c_o = 0, c_end = 0, c_d = 1;
    for (uint32_t P = 0; P < end; P += 128) {
      const uint32_t nncw = load_w(P + 256), nnmw = load_m(P + 256);
      const uint32_t wp = P + 4u * lane;
      // chunks whose bytes are not all ours: bytewise stores at the ends
      const bool edge = !(P >= q && P + 128u <= end);  // (warp-uniform)
      uint32_t vm = 15u;
      if (edge) {
        const uint32_t lo = q > wp ? (q - wp < 4u ? q - wp : 4u) : 0u;
        const uint32_t hi = end > wp ? (end - wp < 4u ? end - wp : 4u) : 0u;
        vm = hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
      }
      const uint32_t hb4 = (mw >> bit_sh) & vm;  // match heads among my bytes
      // the word after mine (lane 31: first word of the next chunk)
      const uint32_t t = __shfl_sync(FULL, cw, next_lane);
      const uint32_t u = __shfl_sync(FULL, ncw, 0);
      const uint32_t nx = lane == 31u ? u : t;
      // my last head, packed: position in chunk (7 bits) | descriptor (23 bits)
      const uint32_t hl = 31u - static_cast<uint32_t>(__clz(static_cast<int>(hb4 | 1u)));
      const uint32_t d_last = lz_funnel(cw, nx, 8u * hl) & 0xffffffu;
      const uint32_t own_pack = (4u * lane + hl) | (d_last << 7);
      const uint32_t hm = __ballot_sync(FULL, hb4 != 0);
      if (hm == 0u && !edge && c_end >= P + 128u) {
        // ---- the chunk lies wholly inside one match (c_o < P: it has no head here): every
        //      source is below the match start, hence below P and final ------------------------
        const uint32_t k0 = wp - c_o;
        const uint32_t m0 = k0 >= c_d ? lz_mod_tab(k0, c_d) : k0;
        mem.stw(wp, lz_gather(mem, c_o, c_d, 0u, 4u, m0));
        __syncwarp();
      } else {
        // the match my first byte may lie in: the last head of the nearest lower lane that has
        // one, else the match carried in from earlier chunks
        const uint32_t below = hm & lt_mask;
        const uint32_t sl = 31u - static_cast<uint32_t>(__clz(static_cast<int>(below | 1u)));
        const uint32_t in_pack = __shfl_sync(FULL, own_pack, static_cast<int>(sl));
        uint32_t t_o = c_o, t_end = c_end, t_d = c_d;
        if (below) {
          t_o = P + (in_pack & 127u);
          t_end = t_o + ((in_pack >> 7) & 255u) + 3u;
          t_d = (in_pack >> 15) + 1u;
        }
        if (hm) {  // warp-uniform: the last head of this chunk is carried into the next ones
          const uint32_t top = 31u - static_cast<uint32_t>(__clz(static_cast<int>(hm)));
          const uint32_t pk = __shfl_sync(FULL, own_pack, static_cast<int>(top));
          c_o = P + (pk & 127u);
          c_end = c_o + ((pk >> 7) & 255u) + 3u;
          c_d = (pk >> 15) + 1u;
        }
        if (edge) lz_chunk_segments<true>(mem, wp, cw, nx, hb4, hl, d_last, vm, t_o, t_end, t_d);
        else lz_chunk_segments<false>(mem, wp, cw, nx, hb4, hl, d_last, vm, t_o, t_end, t_d);
      }
      cw = ncw;
      ncw = nncw;
      mw = nmw;
      nmw = nnmw;
    }
  }
}

}  // namespace sfb
