// Batched raw-DEFLATE decoder for sm_100a: ONE LANE PER STREAM.
//
// Replaces, for the batched path, the reference's
//   bit reader + block-header parser      huffman/src/bit_span.hpp:18-183, src/decompress.cpp:370-385
//   dynamic-Huffman code-length/table build  src/decompress.cpp:253-367, huffman/src/table.hpp:177-216
//   symbol decode + LZ77 copy             src/decompress.cpp:122-242,388-398, huffman/src/decode.hpp:83-102
// with identical per-stream results (bytes and DecompressStatus).
//
// Why a lane and not a warp per stream: Huffman decode of one stream is a serial dependency
// chain.  A warp per stream spends one issue slot per instruction on ONE token; a lane per
// stream retires up to 32 tokens per issued instruction.  Each lane owns
//   * a register bit window (two 32-bit words + one prefetched word, funnel-shift peeks),
//   * a two-level decode LUT in shared memory, interleaved so that element j of lane l sits at
//     u16 index j*32+l (lane pairs share a bank: at most 2-way conflicts on random lookups),
//   * an 8-byte write-combining register for its output window in HBM.
// The token loop is a small state machine (decode | match-copy | stored-copy) so that lanes
// stay converged: every iteration each lane either decodes one token or moves <= 8 bytes.
// The code lengths of the current block live in a per-lane slice of a global scratch buffer
// (written once per block, read sequentially by the table builder and the slow path), which
// keeps the shared-memory footprint at the LUT alone: 896 B per lane -> 8 warps per SM.
//
// Exactness: the fast path only handles tokens that decode cleanly with input to spare.
// Anything unusual (code not in the LUT, LUT pool overflow marker, over-subscribed code set,
// symbols 286/287/30/31, running out of input anywhere inside the token) is handed to
// slow_token(), which restates the reference's bit-serial algorithm on absolute bit positions
// and yields exactly its status.
#pragma once

#include <cstdint>
#ifndef SFB_CPU_EMU
#include <cuda_runtime.h>
#endif

#ifndef SFB_STAT
#define SFB_STAT(name) ((void)0)
#endif

namespace sfb {

enum : uint8_t {
  ST_SUCCESS = 0,
  ST_ERROR = 1,
  ST_INVALID_BLOCK_HEADER = 2,
  ST_LEN_MISMATCH = 3,
  ST_DST_TOO_SMALL = 4,
  ST_SRC_TOO_SMALL = 5,
  ST_INVALID_LIT_OR_LEN = 6,
  ST_INVALID_DISTANCE = 7,
};

// Per-lane slice of the global scratch (u32 words; word j of lane l at g[j*32 + l]):
constexpr int LENS_WORDS = 40;      // [0,40)    320 code lengths as nibbles
constexpr int SCR_FIRST = 40;       // [40,72)   first code per length: lit/len [40,56), distance [56,72)
constexpr int SCR_COUNT = 72;       // [72,104)  count | offset << 16 per length (same split)
constexpr int SCR_SORTED = 104;     // [104,264) symbols in canonical order as u16: lit/len 0..287, distance 288..319
constexpr int SCRATCH_WORDS = 264;

struct BatchArgs {
  const uint8_t* src_base;
  const uint64_t* src_off;
  const uint64_t* src_len;
  uint8_t* dst_base;
  const uint64_t* dst_off;
  const uint64_t* dst_cap;
  uint8_t* status;
  uint64_t* written;  // may be null
  uint64_t n;
  unsigned long long* group_counter;  // dynamic work distribution (zeroed before launch)
  uint32_t* lens_scratch;             // gridDim.x * WARPS * SCRATCH_WORDS * 32 words
};

// ---------------------------------------------------------------------------------------------
// Geometry of one lane's shared-memory slice (u16 units).
template <int ROOT_LIT_, int ROOT_DIST_, int POOL_, int WARPS_>
struct Cfg {
  static constexpr int ROOT_LIT = ROOT_LIT_;
  static constexpr int ROOT_DIST = ROOT_DIST_;
  static constexpr int POOL = POOL_;
  static constexpr int WARPS = WARPS_;
  static constexpr int LIT_OFF = 0;
  static constexpr int DIST_OFF = 1 << ROOT_LIT;
  static constexpr int POOL_OFF = DIST_OFF + (1 << ROOT_DIST);
  static constexpr int LANE_U16 = POOL_OFF + POOL;  // decode LUT entries per lane
  static constexpr int WARP_U16 = LANE_U16 * 32;
  static constexpr int INFO_WORDS = 32;             // shared distance info table
  static constexpr int SMEM_BYTES = WARPS * WARP_U16 * 2 + INFO_WORDS * 4;
  static_assert(POOL <= 510, "sub-table offsets are 9 bits");
  static_assert(POOL >= 64, "the 128-entry CL LUT (one byte each) is overlaid on the pool");
};

// LUT entry (u16).  bits 0-3 = code length L (1..15); L == 0 marks a special entry.
//   L != 0, literal/length table:
//     bit 15 = 1   length code: bits 4-11 = base - 3, bits 12-14 = number of extra bits
//     bit 15 = 0   bits 12-14 kind: 0 literal (bits 4-11 = byte), 1 end of block,
//                  3 symbol 286/287 (rejected by the reference -> slow_token())
//   L != 0, distance table: bits 4-8 = distance symbol (0..31)
//   L == 0: 0x0000 no code here; E_SLOW "not representable, use slow_token()";
//           otherwise a sub-table pointer: bits 4-12 offset from the pool start,
//           bits 13-15 sub-table index bits (1..7)
constexpr uint32_t E_SLOW = 0xFFF0u;   // (sub-table of 2^7 entries at pool offset 511: never allocated)
constexpr uint32_t E_KIND_EOB = 0x1000u;
constexpr uint32_t E_KIND_BAD = 0x3000u;

// RFC 1951 §3.2.5 (reference: src/decompress.cpp:52-84).  info = base | extra << 16
__constant__ uint32_t c_len_info[32] = {
    3 | (0u << 16),   4 | (0u << 16),   5 | (0u << 16),   6 | (0u << 16),   7 | (0u << 16),
    8 | (0u << 16),   9 | (0u << 16),   10 | (0u << 16),  11 | (1u << 16),  13 | (1u << 16),
    15 | (1u << 16),  17 | (1u << 16),  19 | (2u << 16),  23 | (2u << 16),  27 | (2u << 16),
    31 | (2u << 16),  35 | (3u << 16),  43 | (3u << 16),  51 | (3u << 16),  59 | (3u << 16),
    67 | (4u << 16),  83 | (4u << 16),  99 | (4u << 16),  115 | (4u << 16), 131 | (5u << 16),
    163 | (5u << 16), 195 | (5u << 16), 227 | (5u << 16), 258 | (0u << 16), 0, 0, 0};
__constant__ uint32_t c_dist_info[32] = {
    1 | (0u << 16),     2 | (0u << 16),     3 | (0u << 16),      4 | (0u << 16),
    5 | (1u << 16),     7 | (1u << 16),     9 | (2u << 16),      13 | (2u << 16),
    17 | (3u << 16),    25 | (3u << 16),    33 | (4u << 16),     49 | (4u << 16),
    65 | (5u << 16),    97 | (5u << 16),    129 | (6u << 16),    193 | (6u << 16),
    257 | (7u << 16),   385 | (7u << 16),   513 | (8u << 16),    769 | (8u << 16),
    1025 | (9u << 16),  1537 | (9u << 16),  2049 | (10u << 16),  3073 | (10u << 16),
    4097 | (11u << 16), 6145 | (11u << 16), 8193 | (12u << 16),  12289 | (12u << 16),
    16385 | (13u << 16), 24577 | (13u << 16), 0, 0};
// src/decompress.cpp:250-251
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__device__ __forceinline__ uint64_t shl64(uint64_t x, unsigned s) { return s >= 64 ? 0 : x << s; }
__device__ __forceinline__ uint64_t low_bytes_mask(unsigned n)  // n in 0..8
{
  return n >= 8 ? ~0ull : ((1ull << (8 * n)) - 1);
}
__device__ __forceinline__ uint32_t bitrev(uint32_t v, int n) { return __brev(v) >> (32 - n); }
// bits [s, s+32) of the 64-bit value hi:lo, 0 <= s < 32
__device__ __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s)
{
#ifdef SFB_CPU_EMU
  return static_cast<uint32_t>(((static_cast<uint64_t>(hi) << 32) | lo) >> (s & 31u));
#else
  return __funnelshift_r(lo, hi, s);
#endif
}

// ---------------------------------------------------------------------------------------------
// Bit reader (replaces huffman::bit_span, huffman/src/bit_span.hpp).  Bits are consumed LSB
// first.  w0:w1:w2 is a 96-bit window over the stream, w3 and w4 are the two already-fetched
// following words, `bo` the offset of the next unread bit inside the window.  With bo < 32 at
// the start of a token the window holds >= 65 unread bits, more than the longest token
// (15 + 5 + 15 + 13 = 48), so a token is decoded with two funnel-shift peeks and the window
// slides once per token (norm2(): by 0, 1 or 2 words, branch-free; the words loaded there are
// first consumed a whole token later, which hides their latency).  Words at or past the end of
// the stream are fetched as zeros and raise `tail`; while `tail` is clear no consumer can have
// run past the end, so the hot loop only checks overrun() in tail mode.

// word at byte offset `at` when it is not entirely inside the stream (rare): zero-filled
__device__ __noinline__ uint32_t fetch_tail_word(const uint8_t* base, uint32_t at, uint32_t iend)
{
  if (at < iend) {
    const unsigned k = iend - at;  // 1..3 valid bytes
    return *reinterpret_cast<const uint32_t*>(base + at) & ((1u << (8 * k)) - 1u);
  }
  return 0;
}

struct BitReader {
  uint32_t w0, w1, w2, w3, w4;
  uint32_t bo;
  uint32_t ip;          // byte offset (from base) of the next word to fetch (w4 sits at ip-4)
  uint32_t tail;        // some fetched word was not entirely inside the stream
  uint32_t iend;        // byte offset (from base) one past the last stream byte
  uint32_t lead0;       // stream start - base (0..3)
  const uint8_t* base;  // stream start rounded down to 4 bytes

  __device__ __forceinline__ const uint8_t* begin() const { return base + lead0; }

  // (a non-inlined *member* would force this struct out of registers into local memory,
  //  hence the free function for the rare case)
  __device__ __forceinline__ uint32_t fetch(uint32_t at)
  {
    if (at + 4 <= iend) return *reinterpret_cast<const uint32_t*>(base + at);
    tail = 1;
    return fetch_tail_word(base, at, iend);
  }

  __device__ __forceinline__ void open(const uint8_t* begin_, uint32_t len)
  {
    lead0 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(begin_) & 3u);
    base = begin_ - lead0;
    iend = lead0 + len;
    init_at(lead0, 0);
  }

  // start reading at byte offset `at` from base (lead0 <= at <= iend), then skip 0..7 bits
  __device__ __forceinline__ void init_at(uint32_t at, unsigned skip_bits)
  {
    const uint32_t a0 = at & ~3u;
    tail = 0;
    w0 = fetch(a0);
    w1 = fetch(a0 + 4);
    w2 = fetch(a0 + 8);
    w3 = fetch(a0 + 12);
    w4 = fetch(a0 + 16);
    ip = a0 + 20;
    bo = 8 * (at & 3u) + skip_bits;
    norm();
  }

  // slide by one word
  __device__ __forceinline__ void slide1()
  {
    w0 = w1;
    w1 = w2;
    w2 = w3;
    w3 = w4;
    w4 = fetch(ip);
    ip += 4;
    bo -= 32;
  }
  // make bo < 32 again, any bo (header parsing)
  __device__ __forceinline__ void norm()
  {
    while (bo >= 32) slide1();
  }
  // make bo < 32 again for bo < 96, branch-free in the common case (token loop)
  __device__ __forceinline__ void norm2()
  {
    const bool p1 = bo >= 32, p2 = bo >= 64;
    uint32_t n0 = 0, n1 = 0;
    if (p1) {
      if (ip + 8 <= iend) {
        n0 = *reinterpret_cast<const uint32_t*>(base + ip);
        if (p2) n1 = *reinterpret_cast<const uint32_t*>(base + ip + 4);
      } else {
        n0 = fetch(ip);
        if (p2) n1 = fetch(ip + 4);
      }
    }
    w0 = p2 ? w2 : (p1 ? w1 : w0);
    w1 = p2 ? w3 : (p1 ? w2 : w1);
    w2 = p2 ? w4 : (p1 ? w3 : w2);
    w3 = p2 ? n0 : (p1 ? w4 : w3);
    w4 = p2 ? n1 : (p1 ? n0 : w4);
    ip += p2 ? 8u : (p1 ? 4u : 0u);
    bo &= 31u;
  }
  // 32 bits starting `off` bits into the window (off < 64)
  __device__ __forceinline__ uint32_t peek_at(uint32_t off) const
  {
    const bool hi = off >= 32;
    return funnel_r(hi ? w1 : w0, hi ? w2 : w1, off);
  }
  // next 32 bits (requires bo < 32)
  __device__ __forceinline__ uint32_t peek() const { return funnel_r(w0, w1, bo); }
  __device__ __forceinline__ void skip(uint32_t n) { bo += n; }

  // absolute bit position of the next unread bit, relative to the stream start
  // (the window may start before the stream: ip - 20 < lead0 right after init_at)
  __device__ __forceinline__ uint64_t bitpos() const
  {
    return static_cast<uint64_t>(8ll * (static_cast<int64_t>(ip) - 20 - static_cast<int64_t>(lead0)) +
                                 static_cast<int64_t>(bo));
  }
  __device__ __forceinline__ uint64_t total_bits() const
  {
    return 8ull * static_cast<uint64_t>(iend - lead0);
  }
  // real (inside the stream) bits from the next unread bit on; negative after an overrun
  __device__ __forceinline__ int64_t real_left() const
  {
    return static_cast<int64_t>(total_bits()) - static_cast<int64_t>(bitpos());
  }
  __device__ __forceinline__ bool overrun() const { return real_left() < 0; }
  __device__ __forceinline__ void seek_bit(uint64_t bit)
  {
    init_at(lead0 + static_cast<uint32_t>(bit >> 3), static_cast<unsigned>(bit & 7));
  }
};

// Bit-serial access by absolute position (slow paths only).
__device__ __forceinline__ uint32_t bit_at(const uint8_t* begin, uint64_t i)
{
  return (begin[i >> 3] >> (i & 7)) & 1u;
}

// ---------------------------------------------------------------------------------------------
// Code lengths of the current block: 320 nibbles per lane in global scratch, lane-interleaved
// (word j of lane l at g[j*32 + l], so lanes working in lock-step coalesce).  Lit/len lengths
// occupy nibbles [0, n_lit), distance lengths follow at [n_lit, n_lit + n_dist).
struct LensWriter {
  uint32_t* g;
  uint32_t acc;
  int k;
  __device__ __forceinline__ explicit LensWriter(uint32_t* g_) : g(g_), acc(0), k(0) {}
  __device__ __forceinline__ void put(uint32_t v)
  {
    acc |= v << (4 * (k & 7));
    ++k;
    if ((k & 7) == 0) {
      g[((k >> 3) - 1) * 32] = acc;
      acc = 0;
    }
  }
  __device__ __forceinline__ void finish()
  {
    if (k & 7) g[(k >> 3) * 32] = acc;
  }
};

// Sequential reader of `count` nibbles starting at nibble `start`.
struct LensReader {
  const uint32_t* g;
  uint32_t w;
  int s, end;
  __device__ __forceinline__ LensReader(const uint32_t* g_, int start, int count)
      : g(g_), w(0), s(start), end(start + count)
  {
    if (count > 0) w = g[(start >> 3) * 32] >> (4 * (start & 7));
  }
  __device__ __forceinline__ uint32_t next()
  {
    const uint32_t v = w & 15u;
    ++s;
    w >>= 4;
    if ((s & 7) == 0 && s < end) w = g[(s >> 3) * 32];
    return v;
  }
};

struct LaneMem {
  uint16_t* lut;       // element j at lut[j*32]
  uint32_t* lens;      // this lane's slice of the global code-length scratch
};

// ---------------------------------------------------------------------------------------------
// Exact canonical decode of one symbol, bit-serial, from the per-length first-code / count
// arrays and the canonical symbol order that build_lut() leaves in the lane's scratch.
// Restates huffman::decode_one + table::find (huffman/src/decode.hpp:83-102,
// huffman/src/table.hpp:426-452): after each bit a code of the current bitsize matches iff
// (value - first[len]) <u count[len]; the shortest match wins; "not found" when the input is
// exhausted or the bitsize passes the longest code.  Valid for ANY length set the reference
// accepts (incomplete, over-subscribed).  `which` = 0 lit/len, 1 distance.
// Returns the code length (0 = not found).
__device__ __noinline__ int canon_decode(const uint32_t* scratch, int which, const uint8_t* begin,
                                         uint64_t pos, uint64_t end, int* symbol)
{
  const uint32_t* first = scratch + (SCR_FIRST + 16 * which) * 32;
  const uint32_t* cnt_off = scratch + (SCR_COUNT + 16 * which) * 32;
  const uint32_t maxlen = first[0];  // slot 0 holds the longest code length of the table
  uint32_t code = 0;
  for (uint32_t L = 1; L <= maxlen; ++L) {
    if (pos + (L - 1) >= end) return 0;  // ran out of input
    code = (code << 1) | bit_at(begin, pos + (L - 1));
    const uint32_t co = cnt_off[L * 32];
    const uint32_t k = code - first[L * 32];
    if (k < (co & 0xffffu)) {
      const uint32_t i = (which ? 288u : 0u) + (co >> 16) + k;
      const uint32_t w = scratch[(SCR_SORTED + (i >> 1)) * 32];
      *symbol = static_cast<int>((i & 1u) ? (w >> 16) : (w & 0xffffu));
      return static_cast<int>(L);
    }
  }
  return 0;
}

struct SlowToken {
  int status;     // ST_SUCCESS if a token was decoded
  int kind;       // 0 literal, 1 end of block, 2 match
  int value;      // literal byte or match length
  int dist;
  uint64_t next;  // absolute bit position after the token
};

// One token, exactly as decompress_block_huffman's loop body does it
// (src/decompress.cpp:206-240 with decode_lit_or_len :122-144 and
// decompress_length_distance :157-177), bit-serially.  Output-side checks
// (distance > written, room) stay with the caller.
__device__ __noinline__ SlowToken slow_token(const uint32_t* lens, const uint8_t* begin,
                                             uint64_t pos, uint64_t end)
{
  SlowToken t;
  t.status = ST_SUCCESS;
  t.kind = 0;
  t.value = 0;
  t.dist = 0;
  t.next = pos;
  int sym = 0;
  int used = canon_decode(lens, 0, begin, pos, end, &sym);
  if (!used) {
    t.status = ST_INVALID_LIT_OR_LEN;
    return t;
  }
  pos += static_cast<uint64_t>(used);
  if (sym < 256) {
    t.value = sym;
    t.next = pos;
    return t;
  }
  if (sym == 256) {
    t.kind = 1;
    t.next = pos;
    return t;
  }
  if (sym > 285) {
    t.status = ST_INVALID_LIT_OR_LEN;
    return t;
  }
  t.kind = 2;
  uint32_t info = c_len_info[sym - 257];
  int extra = static_cast<int>(info >> 16);
  if (end - pos < static_cast<uint64_t>(extra)) {  // reference: unchecked pop_bits (class U)
    t.status = ST_SRC_TOO_SMALL;
    return t;
  }
  int v = 0;
  for (int i = 0; i < extra; ++i) v |= static_cast<int>(bit_at(begin, pos + i)) << i;
  pos += static_cast<uint64_t>(extra);
  t.value = static_cast<int>(info & 0xffffu) + v;
  used = canon_decode(lens, 1, begin, pos, end, &sym);
  if (!used) {
    t.status = ST_INVALID_DISTANCE;
    return t;
  }
  pos += static_cast<uint64_t>(used);
  if (sym >= 30) {
    t.status = ST_INVALID_LIT_OR_LEN;  // sic: src/decompress.cpp:171-173
    return t;
  }
  info = c_dist_info[sym];
  extra = static_cast<int>(info >> 16);
  if (end - pos < static_cast<uint64_t>(extra)) {
    t.status = ST_SRC_TOO_SMALL;
    return t;
  }
  v = 0;
  for (int i = 0; i < extra; ++i) v |= static_cast<int>(bit_at(begin, pos + i)) << i;
  pos += static_cast<uint64_t>(extra);
  t.dist = static_cast<int>(info & 0xffffu) + v;
  t.next = pos;
  return t;
}

// ---------------------------------------------------------------------------------------------
// Build one two-level LUT from code lengths [s0, s0+n).  Canonical code assignment as in
// huffman::table::canonicalize (huffman/src/table.hpp:177-216).  Length sets the reference
// accepts but a prefix LUT cannot represent exactly (over-subscribed: some code value reaches
// 2^len, "shortest code wins") get E_SLOW in every root slot.
// LITLEN selects the entry payload (see the entry format above).
template <bool LITLEN>
__device__ __forceinline__ uint16_t make_entry(uint32_t s, uint32_t L)
{
  if (!LITLEN) return static_cast<uint16_t>((s << 4) | L);
  if (s < 256) return static_cast<uint16_t>((s << 4) | L);
  if (s == 256) return static_cast<uint16_t>(E_KIND_EOB | L);
  if (s > 285) return static_cast<uint16_t>(E_KIND_BAD | L);
  const uint32_t info = c_len_info[s - 257];
  return static_cast<uint16_t>(0x8000u | ((info >> 16) << 12) | (((info & 0xffffu) - 3u) << 4) | L);
}

template <int ROOT, bool LITLEN, int POOL_OFF>
__device__ __forceinline__ void build_lut(const LaneMem& m, int s0, int n, int root_off, int pool_end,
                          int& pool_at)
{
  uint16_t* const lut = m.lut;
  uint32_t* const g_first = m.lens + (SCR_FIRST + (LITLEN ? 0 : 16)) * 32;
  uint32_t* const g_count = m.lens + (SCR_COUNT + (LITLEN ? 0 : 16)) * 32;
  uint16_t* const g_sorted = reinterpret_cast<uint16_t*>(m.lens + SCR_SORTED * 32);
  constexpr uint32_t SORT0 = LITLEN ? 0u : 288u;
  uint16_t count[16], offset[16];
  uint32_t next[16];
#pragma unroll
  for (int L = 0; L < 16; ++L) count[L] = 0;
  {
    LensReader r(m.lens, s0, n);
#pragma unroll 1
    for (int s = 0; s < n; ++s) count[r.next()]++;
  }
  count[0] = 0;
  bool over = false;
  {
    uint32_t code = 0, off = 0, maxlen = 0;
#pragma unroll
    for (int L = 1; L < 16; ++L) {
      code = (code + count[L - 1]) << 1;
      next[L] = code;
      offset[L] = static_cast<uint16_t>(off);
      g_first[L * 32] = code;
      g_count[L * 32] = count[L] | (off << 16);
      off += count[L];
      if (count[L]) maxlen = L;
      if (code + count[L] > (1u << L)) over = true;
    }
    g_first[0] = maxlen;
  }
#pragma unroll 1
  for (int j = 0; j < (1 << ROOT); ++j)
    lut[(root_off + j) * 32] = over ? static_cast<uint16_t>(E_SLOW) : static_cast<uint16_t>(0);
  // pass A: canonical symbol order for canon_decode(); direct LUT entries; long codes leave
  // (0xF000 | max length) in their root slot
  uint32_t pmin = 1u << ROOT, pmax = 0;
  {
    LensReader r(m.lens, s0, n);
#pragma unroll 1
    for (int s = 0; s < n; ++s) {
      const uint32_t L = r.next();
      if (!L) continue;
      const uint32_t code = next[L]++;
      {
        // rank within the length = symbols of this length seen so far
        const uint32_t i = SORT0 + offset[L]++;
        g_sorted[(i >> 1) * 64 + (i & 1u)] = static_cast<uint16_t>(s);
      }
      if (over) continue;
      if (L <= static_cast<uint32_t>(ROOT)) {
        const uint16_t e = make_entry<LITLEN>(static_cast<uint32_t>(s), L);
        for (uint32_t j = bitrev(code, static_cast<int>(L)); j < (1u << ROOT); j += 1u << L)
          lut[(root_off + j) * 32] = e;
      } else {
        const uint32_t pfx = code >> (L - ROOT);
        uint16_t& slot = lut[(root_off + bitrev(pfx, ROOT)) * 32];
        const uint32_t prev = slot & 15u;
        slot = static_cast<uint16_t>(0xF000u | (L > prev ? L : prev));
        pmin = pfx < pmin ? pfx : pmin;
        pmax = pfx > pmax ? pfx : pmax;
      }
    }
  }
  if (pmax < pmin) return;  // no long codes
  // pass B: carve sub-tables out of the pool
#pragma unroll 1
  for (uint32_t pfx = pmin; pfx <= pmax; ++pfx) {
    uint16_t& slot = lut[(root_off + bitrev(pfx, ROOT)) * 32];
    const uint32_t e = slot;
    if ((e & 0xF000u) != 0xF000u) continue;
    const int sb = static_cast<int>(e & 15u) - ROOT;
    const int size = 1 << sb;
    if (sb <= 7 && pool_at + size <= pool_end) {
      slot = static_cast<uint16_t>((static_cast<uint32_t>(sb) << 13) |
                                   (static_cast<uint32_t>(pool_at - POOL_OFF) << 4));
      for (int j = 0; j < size; ++j) lut[(pool_at + j) * 32] = 0;
      pool_at += size;
    } else {
      slot = static_cast<uint16_t>(E_SLOW);  // pool exhausted: these codes go through slow_token()
    }
  }
  // pass C: fill sub-tables
  {
    uint32_t code = 0;
#pragma unroll
    for (int L = 1; L < 16; ++L) {
      code = (code + count[L - 1]) << 1;
      next[L] = code;
    }
  }
  {
    LensReader r(m.lens, s0, n);
#pragma unroll 1
    for (int s = 0; s < n; ++s) {
      const uint32_t L = r.next();
      if (L <= static_cast<uint32_t>(ROOT)) continue;  // next[] of short lengths is not needed
      const uint32_t code = next[L]++;
      const uint32_t rest = L - ROOT;
      const uint32_t e = lut[(root_off + bitrev(code >> rest, ROOT)) * 32];
      if (e == E_SLOW) continue;
      const uint32_t sb = e >> 13;
      const uint32_t off = POOL_OFF + ((e >> 4) & 0x1ffu);
      const uint16_t v = make_entry<LITLEN>(static_cast<uint32_t>(s), L);
      for (uint32_t j = bitrev(code & ((1u << rest) - 1u), static_cast<int>(rest)); j < (1u << sb);
           j += 1u << rest)
        lut[(off + j) * 32] = v;
    }
  }
}

// LUT lookup.  Returns a direct entry (L != 0), 0 (no code) or E_SLOW (ask slow_token()).
template <int ROOT, int POOL_OFF, uint32_t POOL>
__device__ __forceinline__ uint32_t lut_lookup(const uint16_t* lut, int root_off, uint32_t bits)
{
  uint32_t e = lut[(root_off + static_cast<int>(bits & ((1u << ROOT) - 1u))) * 32];
  if ((e & 15u) == 0 && e != 0 && e != E_SLOW) {  // sub-table pointer (codes longer than ROOT)
    // (lanes that are not decoding run this on stale table contents and discard the result:
    //  the clamp keeps even a garbage pointer inside this lane's slice)
    const uint32_t sb = e >> 13;
    const uint32_t idx = ((e >> 4) & 0x1ffu) + ((bits >> ROOT) & ((1u << sb) - 1u));
    e = lut[(POOL_OFF + (idx < POOL ? idx : POOL - 1u)) * 32];
  }
  return e;
}

// ---------------------------------------------------------------------------------------------
// Per-lane output window: 8-byte write combining over the stream's dst region.
// Invariant: bytes [vpos & ~7, vpos) of the output live in the low bytes of `obuf` (its
// higher bytes are unspecified); everything below vpos & ~7 is in memory.
struct OutWin {
  uint8_t* al;      // 8-byte aligned address of virtual position 0
  uint32_t lead;    // dst start within the first word (0..7): virtual position of byte 0
  uint32_t vpos;    // virtual position of the next byte to produce
  uint32_t vend;    // virtual position one past the capacity
  uint64_t obuf;

  __device__ __forceinline__ uint32_t written() const { return vpos - lead; }
  __device__ __forceinline__ uint32_t room() const { return vend - vpos; }

  // word of the output at 8-aligned virtual position wv (general case, may be the open word)
  __device__ __forceinline__ uint64_t word_at(uint32_t wv) const
  {
    const uint32_t cur = vpos & ~7u;
    if (wv < cur) return *reinterpret_cast<const uint64_t*>(al + wv);
    return wv == cur ? obuf : 0;
  }

  // append n (1..8) bytes from the low bytes of `chunk` (bytes above n are ignored)
  __device__ __forceinline__ void append(uint64_t chunk, uint32_t n)
  {
    const uint32_t k = vpos & 7u;
    const uint32_t sh = 8 * k;
    const uint64_t merged = (obuf & ((1ull << sh) - 1ull)) | (chunk << sh);
    if (k + n >= 8) {
      const uint32_t wv = vpos & ~7u;
      if (wv >= lead) {
        *reinterpret_cast<uint64_t*>(al + wv) = merged;
      } else {  // first word of an unaligned dst: bytes before `lead` are not ours
        for (uint32_t b = lead; b < 8; ++b) al[b] = static_cast<uint8_t>(merged >> (8 * b));
      }
      obuf = chunk >> ((64 - sh) & 63);  // k == 0: no byte is carried over, any value will do
    } else {
      obuf = merged;
    }
    vpos += n;
  }

  // write the unflushed tail (called once, when the stream ends for any reason)
  __device__ __forceinline__ void flush_tail()
  {
    const uint32_t wv = vpos & ~7u;
    uint32_t b = wv < lead ? lead : wv;
    for (; b < vpos; ++b) al[b] = static_cast<uint8_t>(obuf >> (8 * (b - wv)));
  }
};

// lane states
enum : int { S_DECODE = 0, S_MATCH = 1, S_STORED = 2, S_HEADER = 3, S_DONE = 4 };

// ---------------------------------------------------------------------------------------------
// Block header, lane-local.  Restates read_header (src/decompress.cpp:370-385), the stored
// branch (:416-436) and decode_dynamic_huffman_tables / decode_dynamic_huffman_table
// (:253-367).  Every field is bound-checked here (this is not the hot loop); where the
// reference would read past the input (assert-only pop_bits / pop_16) the documented choices
// of this repository apply (SrcTooSmall; InvalidLitOrLen for malformed repeats).
// Returns the next lane state; *status is set when the state is S_DONE.
template <class C>
__device__ __forceinline__ int parse_block_header(BitReader& br, const LaneMem& m, OutWin& ow, uint32_t& final_block,
                                  int& n_lit, int& n_dist, const uint8_t*& copy_src,
                                  uint32_t& copy_left, int* status)
{
  br.norm();
  if (br.real_left() < 3) {
    *status = ST_INVALID_BLOCK_HEADER;
    return S_DONE;
  }
  const uint32_t hdr = br.peek() & 7u;
  const uint32_t type = hdr >> 1;
  if (type == 3) {
    *status = ST_INVALID_BLOCK_HEADER;
    return S_DONE;
  }
  final_block = hdr & 1u;
  br.skip(3);

  if (type == 0) {
    // stored block: skip to the byte boundary (the window words are byte-aligned with the
    // stream, so the phase of `bo` is the phase of the stream)
    br.skip((8u - (br.bo & 7u)) & 7u);
    br.norm();
    if (br.real_left() < 32) {
      *status = ST_SRC_TOO_SMALL;  // reference: pop_16 past the end (class U)
      return S_DONE;
    }
    const uint32_t w = br.peek();
    br.skip(32);
    const uint32_t len = w & 0xffffu, nlen = w >> 16;
    if (len != ((~nlen) & 0xffffu)) {
      *status = ST_LEN_MISMATCH;
      return S_DONE;
    }
    const uint64_t pos = br.bitpos();
    if (br.total_bits() - pos < 8ull * len) {
      *status = ST_SRC_TOO_SMALL;
      return S_DONE;
    }
    if (ow.room() < len) {
      *status = ST_DST_TOO_SMALL;
      return S_DONE;
    }
    copy_src = br.begin() + (pos >> 3);
    copy_left = len;
    if (len == 0) {
      br.norm();
      if (final_block) {
        *status = ST_SUCCESS;
        return S_DONE;
      }
      return S_HEADER;
    }
    return S_STORED;
  }

  if (type == 1) {
    // fixed codes: src/decompress.cpp:25-40
    n_lit = 288;
    n_dist = 32;
#pragma unroll 1
    for (int j = 0; j < LENS_WORDS; ++j) {
      const int s = j * 8;
      uint32_t v;
      if (s < 144) v = 0x88888888u;
      else if (s < 256) v = 0x99999999u;
      else if (s < 280) v = 0x77777777u;
      else if (s < 288) v = 0x88888888u;
      else v = 0x55555555u;
      m.lens[j * 32] = v;
    }
  } else {
    // dynamic codes
    br.norm();
    if (br.real_left() < 14) {
      *status = ST_SRC_TOO_SMALL;
      return S_DONE;
    }
    uint32_t w = br.peek();
    n_lit = 257 + static_cast<int>(w & 31u);
    n_dist = 1 + static_cast<int>((w >> 5) & 31u);
    const int n_cl = 4 + static_cast<int>((w >> 10) & 15u);
    br.skip(14);
    uint64_t cl_lens = 0;  // 19 x 3 bits, indexed by CL symbol
#pragma unroll 1
    for (int i = 0; i < n_cl; ++i) {
      br.norm();
      if (br.real_left() < 3) {
        *status = ST_SRC_TOO_SMALL;
        return S_DONE;
      }
      cl_lens |= static_cast<uint64_t>(br.peek() & 7u) << (3 * c_cl_order[i]);
      br.skip(3);
    }
    // code-length code: 7-bit LUT of bytes (sym << 3 | len), overlaid on the pool region
    // (two bytes per u16 slot: byte j of this lane lives in slot j >> 1).
    // Filled from the longest length down so that, for over-subscribed sets, the shortest
    // matching code wins exactly as the reference's bit-serial search does.
    uint8_t* cl_lut = reinterpret_cast<uint8_t*>(m.lut + C::POOL_OFF * 32);
    {
      uint32_t cnt[8];
#pragma unroll
      for (int L = 0; L < 8; ++L) cnt[L] = 0;
#pragma unroll 1
      for (int s = 0; s < 19; ++s) cnt[(cl_lens >> (3 * s)) & 7u]++;
      cnt[0] = 0;
      uint32_t first[8];
      first[0] = 0;
      {
        uint32_t code = 0;
#pragma unroll
        for (int L = 1; L < 8; ++L) {
          code = (code + cnt[L - 1]) << 1;
          first[L] = code;
        }
      }
#pragma unroll 1
      for (int j = 0; j < 128; ++j) cl_lut[(j >> 1) * 64 + (j & 1)] = 0;
#pragma unroll 1
      for (int L = 7; L >= 1; --L) {
        uint32_t code = first[L];
#pragma unroll 1
        for (int s = 0; s < 19; ++s) {
          if (static_cast<int>((cl_lens >> (3 * s)) & 7u) != L) continue;
          if (code < (1u << L)) {
            for (uint32_t j = bitrev(code, L); j < 128u; j += 1u << L)
              cl_lut[(j >> 1) * 64 + (j & 1)] = static_cast<uint8_t>((s << 3) | L);
          }
          ++code;
        }
      }
    }
    // the two independent runs of code lengths (src/decompress.cpp:353-360)
    const int total = n_lit + n_dist;
    LensWriter lw(m.lens);
    int run_begin = 0, run_end = n_lit;
    uint32_t prev = 0;
#pragma unroll 1
    for (int i = 0; i < total;) {
      if (i == run_end) {
        run_begin = n_lit;
        run_end = total;
      }
      br.norm();
      const uint32_t bits = br.peek();
      const uint32_t j = bits & 127u;
      const uint32_t e = cl_lut[(j >> 1) * 64 + (j & 1)];
      const uint32_t L = e & 7u;
      const int64_t left = br.real_left();
      if (L == 0 || static_cast<int64_t>(L) > left) {
        *status = ST_INVALID_LIT_OR_LEN;  // src/decompress.cpp:265-267
        return S_DONE;
      }
      br.skip(L);
      const uint32_t sym = e >> 3;
      if (sym < 16) {
        lw.put(sym);
        prev = sym;
        ++i;
        continue;
      }
      const uint32_t xbits = sym == 16 ? 2u : sym == 17 ? 3u : 7u;
      if (left - static_cast<int64_t>(L) < static_cast<int64_t>(xbits)) {
        *status = ST_SRC_TOO_SMALL;  // reference: unchecked pop_bits (class U)
        return S_DONE;
      }
      // L + xbits <= 14 bits: still inside the peeked word
      const int repeat = static_cast<int>((bits >> L) & ((1u << xbits) - 1u)) + (sym == 18 ? 11 : 3);
      br.skip(xbits);
      uint32_t v = 0;
      if (sym == 16) {
        if (i == run_begin) {
          *status = ST_INVALID_LIT_OR_LEN;  // reference reads code_bitsizes[-1] (class U)
          return S_DONE;
        }
        v = prev;
      }
      if (i + repeat > run_end) {
        *status = ST_INVALID_LIT_OR_LEN;  // reference writes past code_bitsizes (class U)
        return S_DONE;
      }
#pragma unroll 1
      for (int r = 0; r < repeat; ++r) lw.put(v);
      prev = v;
      i += repeat;
    }
    lw.finish();
  }
  int pool_at = C::POOL_OFF;
  build_lut<C::ROOT_LIT, true, C::POOL_OFF>(m, 0, n_lit, C::LIT_OFF, C::POOL_OFF + C::POOL, pool_at);
  build_lut<C::ROOT_DIST, false, C::POOL_OFF>(m, n_lit, n_dist, C::DIST_OFF, C::POOL_OFF + C::POOL,
                                              pool_at);
  br.norm();
  return S_DECODE;
}

// ---------------------------------------------------------------------------------------------
// The batch kernel.  One lane per stream, 32 streams per warp, warps pull groups of 32
// consecutive streams from a global counter.
//
// Control structure (kept free of `break`s so that the warp reconverges after every stage):
//   rounds:  lanes that need a block header parse it (lock-step when several do), then
//   tokens:  while any lane is inside a block, every iteration runs three converged stages
//              1. decode   lanes in S_DECODE decode one token (literal | end of block | match)
//              2. source   lanes in S_MATCH / S_STORED fetch <= 8 source bytes
//              3. append   lanes with bytes merge them into their write-combining word
//            lanes waiting for the next header (or finished) idle until the round ends.
template <class C>
__global__ void __launch_bounds__(C::WARPS * 32)
inflate_lanes_kernel(const BatchArgs a)
{
#ifdef SFB_CPU_EMU
  uint16_t* const smem = SFB_EMU_SMEM;  // tests/cpu_emu: logic-only build, never shipped
#else
  extern __shared__ __align__(16) uint16_t smem[];
#endif
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = static_cast<int>(threadIdx.x & 31u);
  const int warp = static_cast<int>(threadIdx.x >> 5);
  uint32_t* const s_dist_info = reinterpret_cast<uint32_t*>(smem + C::WARPS * C::WARP_U16);
  for (unsigned t = threadIdx.x; t < 32; t += blockDim.x) s_dist_info[t] = c_dist_info[t];
  __syncthreads();

  LaneMem m;
  m.lut = smem + warp * C::WARP_U16 + lane;
  m.lens = a.lens_scratch +
           (static_cast<size_t>(blockIdx.x) * C::WARPS + static_cast<size_t>(warp)) * (SCRATCH_WORDS * 32) +
           lane;
  const uint16_t* const lut = m.lut;

  const uint64_t n_groups = (a.n + 31) / 32;
  for (;;) {
    unsigned long long g = 0;
    if (lane == 0) g = atomicAdd(a.group_counter, 1ull);
    g = __shfl_sync(FULL, g, 0);
    if (g >= n_groups) break;
    const uint64_t idx = g * 32 + static_cast<uint64_t>(lane);

    int state = S_DONE;
    int status = ST_SUCCESS;
    BitReader br;
    OutWin ow;
    // lanes without a stream still run the (predicated-off) token stages: give them a window
    // that never slides and never loads
    br.w0 = br.w1 = br.w2 = br.w3 = br.w4 = 0;
    br.bo = br.ip = br.tail = br.iend = br.lead0 = 0;
    br.base = nullptr;
    ow.al = nullptr;
    ow.lead = ow.vpos = ow.vend = 0;
    ow.obuf = 0;
    uint32_t final_block = 0;
    int n_lit = 0, n_dist = 0;
    uint32_t mlen = 0, mdist = 0;
    const uint8_t* copy_src = nullptr;
    uint32_t copy_left = 0;
    bool live = idx < a.n;
    if (live) {
      const uint64_t slen = a.src_len[idx];
      const uint64_t cap = a.dst_cap[idx];
      if (slen >= 0xffffff00ull || cap >= 0xffffff00ull) {
        a.status[idx] = ST_ERROR;  // outside the batch precondition
        if (a.written) a.written[idx] = 0;
        live = false;
      } else {
        br.open(a.src_base + a.src_off[idx], static_cast<uint32_t>(slen));
        uint8_t* d = a.dst_base + a.dst_off[idx];
        ow.lead = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(d) & 7u);
        ow.al = d - ow.lead;
        ow.vpos = ow.lead;
        ow.vend = ow.lead + static_cast<uint32_t>(cap);
        ow.obuf = 0;
        state = S_HEADER;
      }
    }

    // pending output chunk (software pipeline): the bytes decided in iteration j are merged
    // into the output window in iteration j+1, after the next token has been decoded, so the
    // loads of a match source overlap a whole decode stage.
    //   chunk = (pw0 >> psh) | (pw1 << (64 - psh)), pn bytes (0 = nothing pending)
    uint64_t pw0 = 0, pw1 = 0;
    uint32_t psh = 0, pn = 0;

    while (__any_sync(FULL, state != S_DONE)) {
      if (state == S_HEADER) {
        state = parse_block_header<C>(br, m, ow, final_block, n_lit, n_dist, copy_src, copy_left,
                                      &status);
      }
      // ---- token / copy iterations (all 32 lanes stay in this loop together) ----------------
      while (__any_sync(FULL, (state <= S_STORED) | (pn != 0))) {
        // ---- stage 1: decode one token (straight-line; results are only used by lanes in
        //      S_DECODE, the others compute on garbage and discard) ---------------------------
        const bool dec = state == S_DECODE;
        SFB_STAT(tokens);
        const uint32_t ip0 = br.ip;   // token start (slow path only)
        const uint32_t bo0 = br.bo;   // < 32
        const uint32_t bits = br.peek();
        const uint32_t e = lut_lookup<C::ROOT_LIT, C::POOL_OFF, C::POOL>(lut, C::LIT_OFF, bits);
        const uint32_t L = e & 15u;
        const bool is_len = (e & 0x8000u) != 0;           // (L != 0 for every direct entry)
        const uint32_t xb = is_len ? ((e >> 12) & 7u) : 0u;
        uint32_t value = ((e >> 4) & 0xffu) + (is_len ? 3u + ((bits >> L) & ((1u << xb) - 1u)) : 0u);
        const uint32_t used1 = L + xb;
        const uint32_t dbits = br.peek_at(bo0 + used1);   // bo0 + used1 <= 31 + 20
        const uint32_t de = lut_lookup<C::ROOT_DIST, C::POOL_OFF, C::POOL>(lut, C::DIST_OFF, dbits);
        const uint32_t dL = de & 15u;
        const uint32_t dsym = (de >> 4) & 31u;
        const uint32_t dinfo = s_dist_info[dsym];
        const uint32_t dxb = dinfo >> 16;
        uint32_t dist = (dinfo & 0xffffu) + ((dbits >> dL) & ((1u << dxb) - 1u));
        // kinds: literal (bits 12-15 clear), end of block, length; everything else is "slow"
        bool eob = (e & 0xF000u) == E_KIND_EOB;
        bool slow = (L == 0) | ((e & 0xB000u) == E_KIND_BAD) |       // no code / 286,287
                    (is_len & ((dL == 0) | (dsym >= 30u)));          // no distance code / 30,31
        if (dec) br.skip(used1 + (is_len ? dL + dxb : 0u));
        br.norm2();
        if (br.tail) slow |= br.overrun();
        bool is_match = is_len;
        if (dec & slow) {
          // anything the fast path cannot vouch for: redo this token exactly
          SFB_STAT(slow_tokens);
          const uint64_t tok_pos = static_cast<uint64_t>(
              8ll * (static_cast<int64_t>(ip0) - 20 - static_cast<int64_t>(br.lead0)) + bo0);
          const SlowToken t = slow_token(m.lens, br.begin(), tok_pos, br.total_bits());
          if (t.status != ST_SUCCESS) {
            status = t.status;
            state = S_DONE;
          } else {
            is_match = t.kind == 2;
            eob = t.kind == 1;
            value = static_cast<uint32_t>(t.value);
            dist = static_cast<uint32_t>(t.dist);
            br.seek_bit(t.next);
          }
        }
        // ---- stage 2: merge the chunk decided in the previous iteration ----------------------
        if (pn) {
          const uint64_t s8 = (pw0 >> psh) | ((pw1 << 1) << (63 - psh));
          ow.append(s8, pn);
          pn = 0;
        }
        // ---- stage 3: act on the token (the output position is up to date now) ---------------
        if (state == S_DECODE) {
          if (eob) {
            status = ST_SUCCESS;
            state = final_block ? S_DONE : S_HEADER;
          } else if (is_match) {
            if (dist > ow.written()) {            // src/decompress.cpp:178-180
              status = ST_INVALID_DISTANCE;
              state = S_DONE;
            } else if (ow.room() < value) {       // :181-183 (no partial copy)
              status = ST_DST_TOO_SMALL;
              state = S_DONE;
            } else {
              mlen = value;
              mdist = dist;
              state = S_MATCH;
            }
          } else if (ow.room() < 1) {             // decompress_literal, :150-152
            status = ST_DST_TOO_SMALL;
            state = S_DONE;
          } else {
            pw0 = value;
            pw1 = 0;
            psh = 0;
            pn = 1;
          }
        }
        // ---- stage 4: fetch up to 8 source bytes for lanes that are copying ------------------
        if (state == S_MATCH) {
          // copy_from_before (src/decompress.cpp:388-398), <= 8 bytes per iteration
          const uint32_t vs = ow.vpos - mdist;
          const uint32_t wv = vs & ~7u;
          const uint32_t sh = 8 * (vs & 7u);
          if (mdist >= 15) {
            // both source words are already in memory (strictly below the open word);
            // they are shifted together when the chunk is merged, one iteration from now
            pw0 = *reinterpret_cast<const uint64_t*>(ow.al + wv);
            pw1 = *reinterpret_cast<const uint64_t*>(ow.al + wv + 8);
            psh = sh;
          } else {
            uint64_t s8 = ow.word_at(wv) >> sh;
            if (sh) s8 |= ow.word_at(wv + 8) << (64 - sh);
            if (mdist < 8) {  // overlapping: replicate the mdist-byte period
              s8 &= low_bytes_mask(mdist);
              s8 |= shl64(s8, 8 * mdist);
              s8 |= shl64(s8, 16 * mdist);
              s8 |= shl64(s8, 32 * mdist);
            }
            pw0 = s8;
            pw1 = 0;
            psh = 0;
          }
          pn = mlen < 8 ? mlen : 8;
          mlen -= pn;
          if (mlen == 0) state = S_DECODE;
        } else if (state == S_STORED) {
          // stored payload: up to 8 bytes from the input (src/decompress.cpp:434)
          const unsigned k = static_cast<unsigned>(reinterpret_cast<uintptr_t>(copy_src) & 7u);
          const uint8_t* wa = copy_src - k;
          uint64_t s8 = *reinterpret_cast<const uint64_t*>(wa) >> (8 * k);
          pn = copy_left < 8 ? copy_left : 8;
          if (k && pn > 8 - k) s8 |= *reinterpret_cast<const uint64_t*>(wa + 8) << (64 - 8 * k);
          pw0 = s8;
          pw1 = 0;
          psh = 0;
          copy_src += pn;
          copy_left -= pn;
          if (copy_left == 0) {
            br.init_at(static_cast<uint32_t>(copy_src - br.base), 0);
            status = ST_SUCCESS;
            state = final_block ? S_DONE : S_HEADER;
          }
        }
      }
      if (state == S_DONE && live) {
        ow.flush_tail();
        a.status[idx] = static_cast<uint8_t>(status);
        if (a.written) a.written[idx] = ow.written();
        live = false;
      }
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
