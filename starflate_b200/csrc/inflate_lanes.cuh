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
//   * a register bit buffer (64-bit window + one prefetched 32-bit word),
//   * a two-level decode LUT in shared memory, interleaved so that element j of lane l sits at
//     u16 index j*32+l (lane pairs share a bank: at most 2-way conflicts on random lookups),
//   * an 8-byte write-combining register for its output window in HBM.
// The token loop is a small state machine (decode | match-copy | stored-copy) so that lanes
// stay converged: every iteration each lane either decodes one token or moves <= 8 bytes.
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
  static constexpr int LUT_U16 = POOL_OFF + POOL;   // decode LUT entries per lane
  static constexpr int LENS_U16 = 80;               // 320 code lengths as nibbles
  static constexpr int WORK_U16 = 32;               // count[16], next[16] during table build
  static constexpr int LANE_U16 = LUT_U16 + LENS_U16 + WORK_U16;
  static constexpr int WARP_U16 = LANE_U16 * 32;
  static constexpr int INFO_WORDS = 32;             // shared distance info table
  static constexpr int SMEM_BYTES = WARPS * WARP_U16 * 2 + INFO_WORDS * 4;
  static_assert(POOL <= 510, "sub-table offsets are 9 bits");
  static_assert(POOL >= 128, "the 128-entry CL LUT (one byte per u16 slot) is overlaid on the pool");
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
constexpr uint32_t E_SLOW = 0xFFF0u;
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
__device__ __forceinline__ uint64_t shr64(uint64_t x, unsigned s) { return s >= 64 ? 0 : x >> s; }
__device__ __forceinline__ uint64_t low_bytes_mask(unsigned n)  // n in 0..8
{
  return n >= 8 ? ~0ull : ((1ull << (8 * n)) - 1);
}
__device__ __forceinline__ uint32_t bitrev(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// ---------------------------------------------------------------------------------------------
// Bit reader (replaces huffman::bit_span, huffman/src/bit_span.hpp).  Bits are consumed LSB
// first.  `buf` holds `cnt` bits, `nextw` is the already-fetched following word.  Words past
// the end of the stream are fetched as zeros and counted in `phantom`, so the window always
// looks full; real_left() tells how many fetched-but-unconsumed bits are real.
struct BitReader {
  uint64_t buf;
  int cnt;
  int phantom;
  uint32_t nextw;
  uint32_t ip;          // byte offset (from base) of nextw's word; multiple of 4
  uint32_t iend;        // byte offset (from base) one past the last stream byte
  uint32_t lead0;       // begin - base (0..3)
  const uint8_t* base;  // stream start rounded down to 4 bytes

  __device__ __forceinline__ const uint8_t* begin() const { return base + lead0; }

  // general fetch (any position); the hot loop inlines the in-range case in refill()
  __device__ __forceinline__ uint32_t fetch(uint32_t at, int& ph) const
  {
    if (at + 4 <= iend) {
      ph = 0;
      return *reinterpret_cast<const uint32_t*>(base + at);
    }
    if (at < iend) {
      const unsigned k = iend - at;  // 1..3 valid bytes
      ph = 32 - 8 * static_cast<int>(k);
      return *reinterpret_cast<const uint32_t*>(base + at) & ((1u << (8 * k)) - 1u);
    }
    ph = 32;
    return 0;
  }

  __device__ void open(const uint8_t* begin_, uint32_t len)
  {
    lead0 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(begin_) & 3u);
    base = begin_ - lead0;
    iend = lead0 + len;
    init_at(lead0, 0);
  }

  // start reading at byte offset `at` from base (lead0 <= at <= iend), then skip 0..7 bits
  __device__ void init_at(uint32_t at, unsigned skip_bits)
  {
    const uint32_t lead = at & 3u;
    const uint32_t a0 = at - lead;
    int ph0 = 32, ph1 = 32;
    uint32_t w0 = 0;
    if (at < iend) w0 = fetch(a0, ph0);
    buf = static_cast<uint64_t>(w0 >> (8 * lead));
    cnt = 32 - 8 * static_cast<int>(lead);
    if (ph0 > cnt) ph0 = cnt;  // stream shorter than the rest of this word
    ip = a0 + 4;
    nextw = fetch(ip, ph1);
    phantom = ph0 + ph1;
    refill();
    buf >>= skip_bits;
    cnt -= static_cast<int>(skip_bits);
  }

  __device__ __forceinline__ void refill()
  {
    if (cnt <= 32) {
      buf |= static_cast<uint64_t>(nextw) << cnt;
      cnt += 32;
      ip += 4;
      if (ip + 4 <= iend) {
        nextw = *reinterpret_cast<const uint32_t*>(base + ip);
      } else {
        int ph;
        nextw = fetch(ip, ph);
        phantom += ph;
      }
    }
  }
  __device__ __forceinline__ void drop(uint32_t n)
  {
    buf >>= n;
    cnt -= static_cast<int>(n);
  }
  __device__ __forceinline__ uint32_t peek32() const { return static_cast<uint32_t>(buf); }
  // real bits among the fetched, unconsumed ones (negative after an overrun)
  __device__ __forceinline__ int real_left() const { return cnt + 32 - phantom; }
  // absolute bit position of the next unread bit, relative to the stream start
  __device__ __forceinline__ uint64_t bitpos() const
  {
    return 8ull * static_cast<uint64_t>(ip - lead0) - static_cast<uint64_t>(cnt);
  }
  __device__ __forceinline__ uint64_t total_bits() const
  {
    return 8ull * static_cast<uint64_t>(iend - lead0);
  }
  __device__ void seek_bit(uint64_t bit)
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
// Lane-interleaved shared memory views.
struct LaneMem {
  uint16_t* lut;   // element j at lut[j*32]
  uint16_t* lens;  // nibble s at (lens[(s>>2)*32] >> 4*(s&3)) & 15
  uint16_t* work;  // count[L] at work[L*32], next[L] at work[(16+L)*32]

  __device__ __forceinline__ uint32_t len_of(int s) const
  {
    return (lens[(s >> 2) * 32] >> ((s & 3) * 4)) & 15u;
  }
  __device__ __forceinline__ void set_len(int s, uint32_t v) const
  {
    uint16_t& w = lens[(s >> 2) * 32];
    const int sh = (s & 3) * 4;
    w = static_cast<uint16_t>((w & ~(15u << sh)) | (v << sh));
  }
};

// ---------------------------------------------------------------------------------------------
// Exact canonical decode of one symbol, bit-serial, straight from the code lengths.
// Restates huffman::decode_one + table::find (huffman/src/decode.hpp:83-102,
// huffman/src/table.hpp:426-452) over the table that huffman::table(symbol_bitsize, ...) +
// canonicalize() would build (table.hpp:177-216,360-376) for lengths lens[s0 .. s0+n).
// Returns the code length (0 = not found: unassigned code, or input exhausted first).
__device__ __noinline__ int canon_decode(const LaneMem& m, int s0, int n, const uint8_t* begin,
                                         uint64_t pos, uint64_t end, int* symbol)
{
  uint16_t cnt[16];
#pragma unroll
  for (int L = 0; L < 16; ++L) cnt[L] = 0;
  int maxlen = 0;
  for (int s = 0; s < n; ++s) {
    const int L = static_cast<int>(m.len_of(s0 + s));
    if (L) {
      cnt[L]++;
      if (L > maxlen) maxlen = L;
    }
  }
  uint32_t code = 0, first = 0;
  for (int L = 1; L <= maxlen; ++L) {
    if (pos + static_cast<uint64_t>(L - 1) >= end) return 0;  // ran out of input
    code = (code << 1) | bit_at(begin, pos + static_cast<uint64_t>(L - 1));
    first = (first + (L > 1 ? cnt[L - 1] : 0u)) << 1;
    const uint32_t k = code - first;
    if (cnt[L] && k < cnt[L]) {
      // k-th symbol (ascending symbol order) among those of length L
      uint32_t seen = 0;
      for (int s = 0; s < n; ++s) {
        if (static_cast<int>(m.len_of(s0 + s)) == L) {
          if (seen == k) {
            *symbol = s;
            return L;
          }
          ++seen;
        }
      }
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
__device__ __noinline__ SlowToken slow_token(const LaneMem& m, int n_lit, int n_dist,
                                             const uint8_t* begin, uint64_t pos, uint64_t end)
{
  SlowToken t;
  t.status = ST_SUCCESS;
  t.kind = 0;
  t.value = 0;
  t.dist = 0;
  int sym = 0;
  int used = canon_decode(m, 0, n_lit, begin, pos, end, &sym);
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
  used = canon_decode(m, n_lit, n_dist, begin, pos, end, &sym);
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
// Build one two-level LUT from code lengths lens[s0 .. s0+n).  Canonical code assignment as in
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
__device__ void build_lut(const LaneMem& m, int s0, int n, int root_off, int pool_end,
                          int& pool_at)
{
  uint16_t* const lut = m.lut;
  uint16_t* const count = m.work;
  uint16_t* const next = m.work + 16 * 32;
#pragma unroll 1
  for (int L = 0; L < 16; ++L) count[L * 32] = 0;
#pragma unroll 1
  for (int s = 0; s < n; ++s) {
    const uint32_t L = m.len_of(s0 + s);
    count[L * 32]++;
  }
  count[0] = 0;
  bool over = false;
  {
    uint32_t code = 0;
#pragma unroll 1
    for (int L = 1; L < 16; ++L) {
      code = (code + count[(L - 1) * 32]) << 1;
      next[L * 32] = static_cast<uint16_t>(code);
      if (code + count[L * 32] > (1u << L)) over = true;
    }
  }
  if (over) {
#pragma unroll 1
    for (int j = 0; j < (1 << ROOT); ++j) lut[(root_off + j) * 32] = static_cast<uint16_t>(E_SLOW);
    return;
  }
#pragma unroll 1
  for (int j = 0; j < (1 << ROOT); ++j) lut[(root_off + j) * 32] = 0;
  // pass A: direct entries; long codes leave (0xF000 | max length) in their root slot
  uint32_t pmin = 1u << ROOT, pmax = 0;
#pragma unroll 1
  for (int s = 0; s < n; ++s) {
    const uint32_t L = m.len_of(s0 + s);
    if (!L) continue;
    const uint32_t code = next[L * 32]++;
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
#pragma unroll 1
    for (int L = 1; L < 16; ++L) {
      code = (code + count[(L - 1) * 32]) << 1;
      next[L * 32] = static_cast<uint16_t>(code);
    }
  }
#pragma unroll 1
  for (int s = 0; s < n; ++s) {
    const uint32_t L = m.len_of(s0 + s);
    if (L <= static_cast<uint32_t>(ROOT)) continue;
    const uint32_t code = next[L * 32]++;
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

// LUT lookup.  Returns a direct entry (L != 0), 0 (no code) or E_SLOW (ask slow_token()).
template <int ROOT, int POOL_OFF>
__device__ __forceinline__ uint32_t lut_lookup(const uint16_t* lut, int root_off, uint32_t bits)
{
  uint32_t e = lut[(root_off + static_cast<int>(bits & ((1u << ROOT) - 1u))) * 32];
  if ((e & 15u) == 0 && e != 0 && e != E_SLOW) {  // sub-table pointer (codes longer than ROOT)
    const uint32_t sb = e >> 13;
    const uint32_t off = POOL_OFF + ((e >> 4) & 0x1ffu);
    e = lut[(off + ((bits >> ROOT) & ((1u << sb) - 1u))) * 32];
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
  __device__ void flush_tail()
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
__device__ int parse_block_header(BitReader& br, const LaneMem& m, OutWin& ow, bool& final_block,
                                  int& n_lit, int& n_dist, const uint8_t*& copy_src,
                                  uint32_t& copy_left, int* status)
{
  br.refill();
  if (br.real_left() < 3) {
    *status = ST_INVALID_BLOCK_HEADER;
    return S_DONE;
  }
  const uint32_t hdr = br.peek32() & 7u;
  const uint32_t type = hdr >> 1;
  if (type == 3) {
    *status = ST_INVALID_BLOCK_HEADER;
    return S_DONE;
  }
  final_block = (hdr & 1u) != 0;
  br.drop(3);

  if (type == 0) {
    // stored block
    br.drop(br.cnt & 7);  // phantom words are whole bytes, so cnt mod 8 is the stream's phase
    br.refill();
    if (br.real_left() < 32) {
      *status = ST_SRC_TOO_SMALL;  // reference: pop_16 past the end (class U)
      return S_DONE;
    }
    const uint32_t w = br.peek32();
    br.drop(32);
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
    for (int j = 0; j < 80; ++j) {
      const int s = j * 4;
      uint32_t v;
      if (s < 144) v = 0x8888u;
      else if (s < 256) v = 0x9999u;
      else if (s < 280) v = 0x7777u;
      else if (s < 288) v = 0x8888u;
      else v = 0x5555u;
      m.lens[j * 32] = static_cast<uint16_t>(v);
    }
  } else {
    // dynamic codes
    br.refill();
    if (br.real_left() < 14) {
      *status = ST_SRC_TOO_SMALL;
      return S_DONE;
    }
    uint32_t w = br.peek32();
    n_lit = 257 + static_cast<int>(w & 31u);
    n_dist = 1 + static_cast<int>((w >> 5) & 31u);
    const int n_cl = 4 + static_cast<int>((w >> 10) & 15u);
    br.drop(14);
    uint64_t cl_lens = 0;  // 19 x 3 bits, indexed by CL symbol
#pragma unroll 1
    for (int i = 0; i < n_cl; ++i) {
      br.refill();
      if (br.real_left() < 3) {
        *status = ST_SRC_TOO_SMALL;
        return S_DONE;
      }
      cl_lens |= static_cast<uint64_t>(br.peek32() & 7u) << (3 * c_cl_order[i]);
      br.drop(3);
    }
    // code-length code: 7-bit LUT of bytes (sym << 3 | len), overlaid on the pool region.
    // Filled from the longest length down so that, for over-subscribed sets, the shortest
    // matching code wins exactly as the reference's bit-serial search does.
    uint8_t* cl_lut = reinterpret_cast<uint8_t*>(m.lut + C::POOL_OFF * 32);  // byte j at [j*64]
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
      for (int j = 0; j < 128; ++j) cl_lut[j * 64] = 0;
#pragma unroll 1
      for (int L = 7; L >= 1; --L) {
        uint32_t code = first[L];
#pragma unroll 1
        for (int s = 0; s < 19; ++s) {
          if (static_cast<int>((cl_lens >> (3 * s)) & 7u) != L) continue;
          if (code < (1u << L)) {
            for (uint32_t j = bitrev(code, L); j < 128u; j += 1u << L)
              cl_lut[j * 64] = static_cast<uint8_t>((s << 3) | L);
          }
          ++code;
        }
      }
    }
    // the two independent runs of code lengths (src/decompress.cpp:353-360)
    const int total = n_lit + n_dist;
#pragma unroll 1
    for (int j = 0; j < 80; ++j) m.lens[j * 32] = 0;
    int run_begin = 0, run_end = n_lit;
#pragma unroll 1
    for (int i = 0; i < total;) {
      if (i == run_end) {
        run_begin = n_lit;
        run_end = total;
      }
      br.refill();
      const uint32_t e = cl_lut[(br.peek32() & 127u) * 64];
      const int L = static_cast<int>(e & 7u);
      if (L == 0 || L > br.real_left()) {
        *status = ST_INVALID_LIT_OR_LEN;  // src/decompress.cpp:265-267
        return S_DONE;
      }
      br.drop(L);
      const uint32_t sym = e >> 3;
      if (sym < 16) {
        if (sym) m.set_len(i, sym);
        ++i;
        continue;
      }
      const int xbits = sym == 16 ? 2 : sym == 17 ? 3 : 7;
      if (br.real_left() < xbits) {
        *status = ST_SRC_TOO_SMALL;  // reference: unchecked pop_bits (class U)
        return S_DONE;
      }
      int repeat = static_cast<int>(br.peek32() & ((1u << xbits) - 1u)) + (sym == 18 ? 11 : 3);
      br.drop(xbits);
      uint32_t v = 0;
      if (sym == 16) {
        if (i == run_begin) {
          *status = ST_INVALID_LIT_OR_LEN;  // reference reads code_bitsizes[-1] (class U)
          return S_DONE;
        }
        v = m.len_of(i - 1);
      }
      if (i + repeat > run_end) {
        *status = ST_INVALID_LIT_OR_LEN;  // reference writes past code_bitsizes (class U)
        return S_DONE;
      }
      if (v) {
        for (int j = 0; j < repeat; ++j) m.set_len(i + j, v);
      }
      i += repeat;
    }
  }
  int pool_at = C::POOL_OFF;
  build_lut<C::ROOT_LIT, true, C::POOL_OFF>(m, 0, n_lit, C::LIT_OFF, C::POOL_OFF + C::POOL, pool_at);
  build_lut<C::ROOT_DIST, false, C::POOL_OFF>(m, n_lit, n_dist, C::DIST_OFF, C::POOL_OFF + C::POOL,
                                              pool_at);
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
  m.lens = m.lut + C::LUT_U16 * 32;
  m.work = m.lens + C::LENS_U16 * 32;
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
    bool final_block = false;
    int n_lit = 0, n_dist = 0;
    uint32_t mlen = 0, mdist = 0;
    const uint8_t* copy_src = nullptr;
    uint32_t copy_left = 0;
    bool live = idx < a.n;
    if (live) {
      const uint64_t slen = a.src_len[idx];
      const uint64_t cap = a.dst_cap[idx];
      if (slen >= 0xfffffff0ull || cap >= 0xfffffff0ull) {
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

    while (__any_sync(FULL, state != S_DONE)) {
      if (state == S_HEADER) {
        state = parse_block_header<C>(br, m, ow, final_block, n_lit, n_dist, copy_src, copy_left,
                                      &status);
      }
      // ---- token / copy iterations (all 32 lanes stay in this loop together) ----------------
      while (__any_sync(FULL, state <= S_STORED)) {
        uint64_t chunk = 0;
        uint32_t n = 0;
        // ---- stage 1: decode one token ------------------------------------------------------
        if (state == S_DECODE) {
          br.refill();
          SFB_STAT(tokens);
          const uint32_t ip0 = br.ip;   // token start = 8*(ip0 - lead0) - cnt0 (slow path only)
          const int cnt0 = br.cnt;
          bool slow = false, is_match = false, eob = false;
          uint32_t value = 0, dist = 0;
          uint32_t e = lut_lookup<C::ROOT_LIT, C::POOL_OFF>(lut, C::LIT_OFF, br.peek32());
          uint32_t L = e & 15u;
          slow = (L == 0);
          br.drop(L);
          if (e & 0xF000u) {            // not a literal: end of block, length code, or oddity
            if (e & 0x8000u) {
              if (L) {
                is_match = true;
                const uint32_t xb = (e >> 12) & 7u;
                value = 3u + ((e >> 4) & 0xffu) + (br.peek32() & ((1u << xb) - 1u));
                br.drop(xb);
                br.refill();
                const uint32_t de = lut_lookup<C::ROOT_DIST, C::POOL_OFF>(lut, C::DIST_OFF, br.peek32());
                const uint32_t dL = de & 15u;
                const uint32_t dsym = (de >> 4) & 31u;
                slow = (dL == 0) | (dsym >= 30u);
                br.drop(dL);
                const uint32_t dinfo = s_dist_info[dsym];
                const uint32_t dxb = dinfo >> 16;
                dist = (dinfo & 0xffffu) + (br.peek32() & ((1u << dxb) - 1u));
                br.drop(dxb);
              }
            } else if ((e & 0xF000u) == E_KIND_EOB) {
              eob = true;
            } else {
              slow = true;              // symbols 286 / 287
            }
          } else {
            value = (e >> 4) & 0xffu;
          }
          if (slow | (br.real_left() < 0)) {
            // anything the fast path cannot vouch for: redo this token exactly
            SFB_STAT(slow_tokens);
            const uint64_t tok_pos = 8ull * static_cast<uint64_t>(ip0 - br.lead0) - static_cast<uint64_t>(cnt0);
            const SlowToken t = slow_token(m, n_lit, n_dist, br.begin(), tok_pos, br.total_bits());
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
          if (state == S_DECODE) {
            if (eob) {
              status = ST_SUCCESS;
              state = final_block ? S_DONE : S_HEADER;
            } else if (is_match) {
              if (dist > ow.written()) {          // src/decompress.cpp:178-180
                status = ST_INVALID_DISTANCE;
                state = S_DONE;
              } else if (ow.room() < value) {     // :181-183 (no partial copy)
                status = ST_DST_TOO_SMALL;
                state = S_DONE;
              } else {
                mlen = value;
                mdist = dist;
                state = S_MATCH;
              }
            } else if (ow.room() < 1) {           // decompress_literal, :150-152
              status = ST_DST_TOO_SMALL;
              state = S_DONE;
            } else {
              chunk = value;
              n = 1;
            }
          }
        }
        __syncwarp();
        // ---- stage 2: up to 8 source bytes for lanes that are copying ------------------------
        if (state == S_MATCH) {
          // copy_from_before (src/decompress.cpp:388-398), <= 8 bytes per iteration
          const uint32_t vs = ow.vpos - mdist;
          const uint32_t wv = vs & ~7u;
          const uint32_t sh = 8 * (vs & 7u);
          uint64_t s8;
          if (mdist >= 15) {
            // both source words are already in memory (strictly below the open word)
            const uint64_t w0 = *reinterpret_cast<const uint64_t*>(ow.al + wv);
            const uint64_t w1 = *reinterpret_cast<const uint64_t*>(ow.al + wv + 8);
            s8 = (w0 >> sh) | ((w1 << 1) << (63 - sh));
          } else {
            s8 = ow.word_at(wv) >> sh;
            if (sh) s8 |= ow.word_at(wv + 8) << (64 - sh);
            if (mdist < 8) {  // overlapping: replicate the mdist-byte period
              s8 &= low_bytes_mask(mdist);
              s8 |= shl64(s8, 8 * mdist);
              s8 |= shl64(s8, 16 * mdist);
              s8 |= shl64(s8, 32 * mdist);
            }
          }
          n = mlen < 8 ? mlen : 8;
          chunk = s8;
          mlen -= n;
          if (mlen == 0) state = S_DECODE;
        } else if (state == S_STORED) {
          // stored payload: up to 8 bytes from the input (src/decompress.cpp:434)
          const unsigned k = static_cast<unsigned>(reinterpret_cast<uintptr_t>(copy_src) & 7u);
          const uint8_t* wa = copy_src - k;
          uint64_t s8 = *reinterpret_cast<const uint64_t*>(wa) >> (8 * k);
          n = copy_left < 8 ? copy_left : 8;
          if (k && n > 8 - k) s8 |= *reinterpret_cast<const uint64_t*>(wa + 8) << (64 - 8 * k);
          chunk = s8;
          copy_src += n;
          copy_left -= n;
          if (copy_left == 0) {
            br.init_at(static_cast<uint32_t>(copy_src - br.base), 0);
            status = ST_SUCCESS;
            state = final_block ? S_DONE : S_HEADER;
          }
        }
        __syncwarp();
        // ---- stage 3: merge into the write-combining word ------------------------------------
        if (n) ow.append(chunk, n);
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
