// Batched raw-DEFLATE COMPRESSION for sm_100a, ONE WARP PER STREAM (the step on the
// other side of the hot path, which the reference's README names as the parallelisable direction
// it has not built, README.md:5-7).  First generation: LZ77 by hashing + ONE fixed-Huffman block
// (RFC 1951 §3.2.6; the reference's fixed tables are src/decompress.cpp:25-40), or stored blocks
// when that would be larger than the input.  What it writes is what the decoder of this repository,
// zlib and the reference decompress() all accept; the tests round-trip through all three.
//
// Per step the warp looks at 32 consecutive positions, one per lane:
//   1. hash of the 4 bytes at the position -> candidate (the last earlier position with that hash:
//      a 4 096-entry table of 16-bit positions per warp in shared memory); the highest lane of every
//      hash value then enters its own position (deterministic: __match_any_sync);
//   2. each lane measures its match against the candidate, 4 bytes per comparison (unaligned words
//      put together from aligned ones; never a byte past the stream);
//   3. greedy parse, front to back, by the whole warp: the token at the cursor is the lane's match if
//      it is at least 4 bytes long, else a literal; the cursor moves by its length (possibly past
//      the 32 positions: the next step starts where the cursor stands, positions inside a match are
//      not entered in the table);
//   4. every token lane builds its bits — fixed literal/length code MSB first, extra bits, 5-bit
//      distance code, extra bits: at most 31 — an exclusive scan of the bit counts places them, and
//      the lanes OR them into a 64-word staging window in shared memory that is written out to dst
//      in whole words (bytes where dst is not word-aligned or the capacity ends inside a word).
#pragma once

#include <cstdint>
#ifndef SFB_CPU_EMU
#include <cuda_runtime.h>
#endif

#include "deflate_lane.cuh"

namespace sfb {

struct CompressArgs {
  const uint8_t* src_base;
  const uint64_t* src_off;
  const uint64_t* src_len;
  uint8_t* dst_base;
  const uint64_t* dst_off;
  const uint64_t* dst_cap;
  uint8_t* status;       // ST_SUCCESS, ST_DST_TOO_SMALL (neither form fits dst_cap), ST_ERROR (src_len >= 4 GiB - 256)
  uint64_t* written;
  uint64_t n;
  unsigned long long* counter;  // zeroed before launch
};

constexpr int CMP_WARPS = 8;
constexpr int CMP_HASH_BITS = 12;
constexpr int CMP_STAGE_WORDS = 64;
constexpr int CMP_WARP_BYTES = (1 << CMP_HASH_BITS) * 2 + CMP_STAGE_WORDS * 4;
constexpr int CMP_SMEM_BYTES = CMP_WARPS * CMP_WARP_BYTES;
constexpr uint32_t CMP_MIN_MATCH = 4;
constexpr uint32_t CMP_MAX_MATCH = 258;
constexpr uint32_t CMP_MAX_DIST = 32768;

// bytes [p, p + 4) of the stream at `s` as a little-endian word, from aligned words that each hold
// at least one of them (p + 4 <= stream length is the caller's business)
__device__ __forceinline__ uint32_t cmp_load32(const uint8_t* s, uint32_t p)
{
  const uintptr_t a = reinterpret_cast<uintptr_t>(s + p);
  const uint32_t r = static_cast<uint32_t>(a & 3u);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a - r);
  const uint32_t lo = w[0];
  const uint32_t hi = r ? w[1] : 0u;
  return funnel_r(lo, hi, 8u * r);
}

__device__ __forceinline__ uint32_t cmp_ctz(uint32_t v)
{
#ifdef SFB_CPU_EMU
  return static_cast<uint32_t>(__builtin_ctz(v));
#else
  return static_cast<uint32_t>(__ffs(static_cast<int>(v)) - 1);
#endif
}

// fixed-Huffman bits of one token, LSB = first bit of the stream; *nb = how many (<= 31)
__device__ __forceinline__ uint32_t cmp_literal_bits(uint32_t lit, uint32_t* nb)
{
  // 0-143: 8 bits 00110000.. ; 144-255: 9 bits 110010000..  (codes go out MSB first)
  const bool small = lit < 144u;
  const uint32_t code = small ? 0x30u + lit : 0x190u + (lit - 144u);
  const uint32_t n = small ? 8u : 9u;
  *nb = n;
  return __brev(code) >> (32u - n);
}
__device__ __forceinline__ uint32_t cmp_match_bits(uint32_t len, uint32_t dist, uint32_t* nb)
{
  // length symbol 257 + k, extra bits (RFC 1951 §3.2.5)
  const uint32_t l = len - 3u;
  uint32_t k, leb, lev;
  if (len == 258u) {
    k = 28u;
    leb = 0u;
    lev = 0u;
  } else if (l < 8u) {
    k = l;
    leb = 0u;
    lev = 0u;
  } else {
    const uint32_t n = 31u - static_cast<uint32_t>(__clz(static_cast<int>(l)));
    leb = n - 2u;
    k = 4u * leb + 4u + ((l >> leb) & 3u);
    lev = l & ((1u << leb) - 1u);
  }
  // fixed code of 257 + k: 256-279 are 7 bits 0000000.., 280-287 8 bits 11000000..
  const uint32_t sym = 257u + k;
  const bool seven = sym < 280u;
  const uint32_t lcode = seven ? sym - 256u : 0xC0u + (sym - 280u);
  const uint32_t ln = seven ? 7u : 8u;
  uint32_t bits = __brev(lcode) >> (32u - ln);
  uint32_t at = ln;
  bits |= lev << at;
  at += leb;
  // distance symbol, 5-bit fixed code, extra bits
  const uint32_t x = dist - 1u;
  uint32_t ds, deb, dev;
  if (x < 4u) {
    ds = x;
    deb = 0u;
    dev = 0u;
  } else {
    const uint32_t n = 31u - static_cast<uint32_t>(__clz(static_cast<int>(x)));
    deb = n - 1u;
    ds = 2u * n + ((x >> deb) & 1u);
    dev = x & ((1u << deb) - 1u);
  }
  bits |= (__brev(ds) >> 27) << at;
  at += 5u;
  bits |= dev << at;
  at += deb;
  *nb = at;
  return bits;
}

// The warp's bit writer: a window of CMP_STAGE_WORDS words over the output, starting at word `wbase`.
struct CmpWriter {
  uint32_t* stage;   // shared
  uint8_t* d;        // dst of the stream
  uint32_t cap;      // its capacity in bytes
  uint32_t wbase;    // output words already written out
  uint32_t cur;      // bits in the window
  uint32_t lane;

  __device__ __forceinline__ void open(uint32_t* stage_, uint8_t* d_, uint32_t cap_, uint32_t lane_)
  {
    stage = stage_;
    d = d_;
    cap = cap_;
    wbase = 0;
    cur = 0;
    lane = lane_;
    stage[lane] = 0;
    stage[lane + 32] = 0;
    __syncwarp();
  }
  // word j of the window -> dst (only bytes inside the capacity)
  __device__ __forceinline__ void store_word(uint32_t j, uint32_t nbytes)
  {
    const uint32_t at = 4u * (wbase + j);
    const uint32_t v = stage[j];
    uint8_t* p = d + at;
    if (nbytes == 4u && at + 4u <= cap && (reinterpret_cast<uintptr_t>(p) & 3u) == 0u) {
      *reinterpret_cast<uint32_t*>(p) = v;
    } else {
      for (uint32_t b = 0; b < nbytes; ++b)
        if (at + b < cap) p[b] = static_cast<uint8_t>(v >> (8u * b));
    }
  }
  // every lane brings `nb` bits (0: none), in lane order
  __device__ __forceinline__ void put(uint32_t bits, uint32_t nb)
  {
    constexpr unsigned FULL = 0xffffffffu;
    uint32_t off = nb;  // inclusive scan
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, off, static_cast<unsigned>(dlt));
      if (lane >= static_cast<uint32_t>(dlt)) off += t;
    }
    const uint32_t total = __shfl_sync(FULL, off, 31);
    off -= nb;
    if (nb) {
      const uint32_t at = cur + off, w = at >> 5, sh = at & 31u;
      atomicOr(&stage[w], bits << sh);
      if (sh + nb > 32u) atomicOr(&stage[w + 1], bits >> (32u - sh));
    }
    cur += total;
    __syncwarp();
    if (cur >= 32u * 32u) flush_words();
  }
  // write the complete words out and move the open one to the front
  __device__ __forceinline__ void flush_words()
  {
    const uint32_t nw = cur >> 5;   // < CMP_STAGE_WORDS
    for (uint32_t j = lane; j < nw; j += 32u) store_word(j, 4u);
    __syncwarp();
    const uint32_t open_word = stage[nw];
    __syncwarp();
    stage[lane] = lane == 0u ? open_word : 0u;
    stage[lane + 32] = 0;
    wbase += nw;
    cur &= 31u;
    __syncwarp();
  }
  // the end of the stream: everything out; -> total bytes
  __device__ __forceinline__ uint64_t close()
  {
    const uint32_t nw = cur >> 5, rest = (cur & 31u) + 7u >> 3;
    for (uint32_t j = lane; j < nw; j += 32u) store_word(j, 4u);
    if (lane == 0u && rest) store_word(nw, rest);
    __syncwarp();
    return 4ull * (static_cast<uint64_t>(wbase) + nw) + rest;
  }
};

// stored blocks (src/decompress.cpp:416-436 reads them): 5 header bytes per <= 65 535 payload bytes
__device__ __forceinline__ uint64_t cmp_stored_size(uint64_t n) { return n + 5ull * (n ? (n + 65534ull) / 65535ull : 1ull); }

__device__ __forceinline__ void cmp_write_stored(const uint8_t* s, uint32_t n, uint8_t* d, uint32_t lane)
{
  uint32_t done = 0;
  uint64_t at = 0;
  do {
    const uint32_t blk = n - done < 65535u ? n - done : 65535u;
    const bool fin = done + blk == n;
    if (lane == 0u) {
      d[at] = fin ? 1u : 0u;
      d[at + 1] = static_cast<uint8_t>(blk);
      d[at + 2] = static_cast<uint8_t>(blk >> 8);
      d[at + 3] = static_cast<uint8_t>(~blk);
      d[at + 4] = static_cast<uint8_t>((~blk) >> 8);
    }
    for (uint32_t i = lane; i < blk; i += 32u) d[at + 5 + i] = s[done + i];
    at += 5ull + blk;
    done += blk;
  } while (done < n);
}

__global__ void __launch_bounds__(CMP_WARPS * 32) deflate_compress_kernel(const CompressArgs a)
{
#ifdef SFB_CPU_EMU
  uint8_t* const smem = reinterpret_cast<uint8_t*>(SFB_EMU_SMEM);
#else
  extern __shared__ __align__(16) uint8_t smem[];
#endif
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  uint16_t* const tab = reinterpret_cast<uint16_t*>(smem + warp * CMP_WARP_BYTES);
  uint32_t* const stage = reinterpret_cast<uint32_t*>(smem + warp * CMP_WARP_BYTES + (1 << CMP_HASH_BITS) * 2);
  for (;;) {
    unsigned long long si = 0;
    if (lane == 0) si = atomicAdd(a.counter, 1ull);
    si = __shfl_sync(FULL, si, 0);
    if (si >= a.n) break;
    const uint64_t slen = a.src_len[si], cap64 = a.dst_cap[si];
    if (slen >= 0xffffff00ull) {
      if (lane == 0) {
        a.status[si] = ST_ERROR;
        a.written[si] = 0;
      }
      continue;
    }
    const uint8_t* const s = a.src_base + a.src_off[si];
    uint8_t* const d = a.dst_base + a.dst_off[si];
    const uint32_t n = static_cast<uint32_t>(slen);
    const uint32_t cap = cap64 > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(cap64);
    for (uint32_t j = lane; j < (1u << CMP_HASH_BITS) / 2u; j += 32u) reinterpret_cast<uint32_t*>(tab)[j] = 0u;
    CmpWriter w;
    w.open(stage, d, cap, lane);
    w.put(3u, lane == 0u ? 3u : 0u);   // BFINAL = 1, BTYPE = 01
    uint32_t c = 0;                    // the cursor: everything before it is encoded
    while (c < n) {
      const uint32_t pos = c + lane;
      const bool can = pos + CMP_MIN_MATCH <= n;   // four bytes to hash
      uint32_t v = 0, h = 0xffffffffu - lane;      // (lanes without a hash never share one)
      uint32_t cand16 = 0;
      if (can) {
        v = cmp_load32(s, pos);
        h = (v * 0x9E3779B1u) >> (32 - CMP_HASH_BITS);
        cand16 = tab[h];
      }
      __syncwarp();
      {
        const uint32_t same = __match_any_sync(FULL, h);
        if (can && (same >> lane) <= 1u) tab[h] = static_cast<uint16_t>(pos);   // the highest lane with this hash
      }
      // candidate position: the table holds the low 16 bits of a position before this step
      uint32_t mlen = 0, dist = 0;
      if (can) {
        int64_t cand = static_cast<int64_t>((pos & ~0xffffu) | cand16);
        if (cand >= static_cast<int64_t>(c)) cand -= 65536;
        if (cand >= 0 && pos - static_cast<uint32_t>(cand) <= CMP_MAX_DIST) {
          const uint32_t cp = static_cast<uint32_t>(cand);
          const uint32_t maxlen = n - pos < CMP_MAX_MATCH ? n - pos : CMP_MAX_MATCH;
          uint32_t m = 0;
          bool open_end = true;
          while (m + 4u <= maxlen) {
            const uint32_t x = cmp_load32(s, pos + m) ^ cmp_load32(s, cp + m);
            if (x) {
              m += cmp_ctz(x) >> 3;
              open_end = false;
              break;
            }
            m += 4u;
          }
          if (open_end)
            while (m < maxlen && s[pos + m] == s[cp + m]) ++m;
          if (m >= CMP_MIN_MATCH) {
            mlen = m;
            dist = pos - cp;
          }
        }
      }
      // greedy parse of the 32 positions, by the whole warp
      bool start = false;
      uint32_t t = 0;
      while (t < 32u && c + t < n) {
        const uint32_t L = __shfl_sync(FULL, mlen, static_cast<int>(t));
        if (lane == t) start = true;
        t += L ? L : 1u;
      }
      uint32_t bits = 0, nb = 0;
      if (start) bits = mlen ? cmp_match_bits(mlen, dist, &nb) : cmp_literal_bits(s[pos], &nb);
      w.put(bits, nb);
      c += t;
    }
    w.put(0u, lane == 0u ? 7u : 0u);   // end of block: seven zero bits
    const uint64_t comp = w.close();
    const uint64_t stored = cmp_stored_size(n);
    uint64_t out = comp;
    int st = ST_SUCCESS;
    if (comp > stored || comp > cap) {
      if (stored <= cap) {
        __syncwarp();
        cmp_write_stored(s, n, d, lane);
        out = stored;
      } else {
        st = ST_DST_TOO_SMALL;
        out = 0;
      }
    }
    if (lane == 0) {
      a.status[si] = static_cast<uint8_t>(st);
      a.written[si] = out;
    }
    __syncwarp();
  }
}

}  // namespace sfb
