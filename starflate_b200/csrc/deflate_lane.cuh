// Lane-local DEFLATE machinery for sm_100a, shared by the batch kernels: bit reader, block-header
// parser, dynamic-Huffman code-length decode, canonical table / two-level LUT build, and the
// exact bit-serial slow path.  ONE LANE PER STREAM: each CUDA thread runs this code on its own
// stream (see huff_lanes.cuh for why).
//
// Replaces the reference's
//   bit reader + block-header parser         huffman/src/bit_span.hpp:18-183, src/decompress.cpp:370-385
//   dynamic-Huffman code-length/table build  src/decompress.cpp:253-367, huffman/src/table.hpp:177-216
//   symbol decode                            src/decompress.cpp:122-242, huffman/src/decode.hpp:83-102
// with identical per-stream results (tokens and DecompressStatus).
//
// Exactness: the LUT fast path only handles tokens that decode cleanly with input to spare.
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

// Hand-over from pass 1 to a pass 2 that runs BESIDE it (lz_window.cuh: lz_window_queue_kernel): a
// lane publishes every SEG bytes of output it has left behind — item (stream, k) says "the first
// (k + 1) SEG + LAG bytes of this stream's pass-1 output are in memory" — and one final item per
// stream once status and written are.  Items are 64-bit words, 0 = not there yet.
struct QueueArgs {
  unsigned long long* items;       // null: no queue (pass 2 runs after pass 1)
  unsigned long long cap;          // slots in `items` (more than there will ever be items)
  unsigned long long* tail;        // producers
  unsigned long long* head;        // consumers
  unsigned long long* final_tail;  // the number of items there will ever be, once done[1] is set
  unsigned int* done;              // [0] pass-1 CTAs that have exited, [1] all of them have
  unsigned long long* dbg;         // [0] != 0: a consumer gave up waiting (1 = for an item, 2 = for its turn); [1..3] what for
  unsigned int* turn;              // per stream: the segment whose turn it is
  uint32_t* state;                 // per stream: 5 words a segment hands to the next (LzwState)
  uint32_t seg_shift;              // SEG = 1 << seg_shift bytes of output
  uint32_t lag;                    // LAG
};
constexpr uint32_t Q_SEG_SHIFT = 14;
constexpr uint32_t Q_LAG = 4096;   // >= a window (1024) + the look-ahead of a dense window (1024 + 256 + 8) + a match (258) + slack

__device__ __forceinline__ void q_push(const QueueArgs& q, uint64_t idx, uint32_t k, bool fin)
{
#ifndef SFB_CPU_EMU
  __threadfence();   // what this lane stored so far is visible before the item is
  const unsigned long long at = atomicAdd(q.tail, 1ull);
  *reinterpret_cast<volatile unsigned long long*>(q.items + at) =
      ((idx + 1ull) << 24) | (static_cast<unsigned long long>(k) << 1) | (fin ? 1ull : 0ull);
#else
  (void)q; (void)idx; (void)k; (void)fin;
#endif
}

struct BatchArgs {
  const uint8_t* src_base;
  const uint64_t* src_off;
  const uint64_t* src_len;
  uint8_t* dst_base;       // 128-byte aligned
  uint64_t dst_delta;      // added to every dst_off (the caller's dst_base - this dst_base)
  const uint64_t* dst_off;
  const uint64_t* dst_cap;
  uint8_t* status;
  uint64_t* written;  // never null here (the resolve pass reads it)
  uint64_t n;                         // streams of this launch: indices idx_base .. idx_base + n - 1,
  uint64_t idx_base;                  // or todo_list[0 .. n) when a list is given
  unsigned long long* group_counter;  // dynamic work distribution (zeroed before launch)
  uint32_t* lens_scratch;             // gridDim.x * WARPS * SCRATCH_WORDS * 32 words
  uint32_t* match_bits;               // match-head bitmap, bit k <-> dst_base[k] (zeroed before launch)
  // Two LUT geometries: the launch with the small one hands streams whose first
  // block's codes do not fit its tables to a launch with the large one.
  uint32_t* defer_list;                     // out (small geometry): indices of the streams handed on; else null
  unsigned long long* defer_count;          // out: how many
  const uint32_t* todo_list;                // in: the streams to do, in this order (null: idx_base + 0 .. n)
  const unsigned long long* todo_count;     // in: how many of them (null: n)
  uint32_t no_pair;                         // A/B switch (SFB200_NO_PAIR=1): one token per iteration, as before
  const uint8_t* handled;                   // in (may be null): 1 = stored_streams_kernel finished this stream already
  // Chunked input (sfb200_inflate_stream_*; all three null otherwise): stream i starts at bit
  // start_bit[i] of its src — a block header — with start_out[i] bytes of earlier output already
  // in place at the front of its dst region (they count as written: distances may reach into
  // them); blk_end[2i] / blk_end[2i+1] receive the bit position and the output position of the last
  // block boundary the decode passed (what a later call can continue from).
  const uint64_t* start_bit;
  const uint64_t* start_out;
  uint64_t* blk_end;
  QueueArgs q;                              // q.items null: no hand-over queue
};

// ---------------------------------------------------------------------------------------------
// Geometry of one warp's shared-memory slice: the decode LUTs of its 32 lanes (u16 entries,
// element j of lane l at u16 index j*32+l) followed by the lanes' input rings (RING_BLOCKS
// 16-byte blocks per lane, block s of lane l at byte (s*32+l)*16 of the ring region).
constexpr int RING_BLOCKS = 4;
constexpr int RING_WARP_BYTES = RING_BLOCKS * 32 * 16;
template <int ROOT_LIT_, int ROOT_DIST_, int POOL_, int WARPS_, int CTAS_ = 2>
struct Cfg {
  static constexpr int CTAS = CTAS_;                // resident CTAs per SM the kernel is built for
  static constexpr int ROOT_LIT = ROOT_LIT_;
  static constexpr int ROOT_DIST = ROOT_DIST_;
  static constexpr int POOL = POOL_;
  static constexpr int WARPS = WARPS_;
  static constexpr int LIT_OFF = 0;
  static constexpr int DIST_OFF = 1 << ROOT_LIT;
  static constexpr int POOL_OFF = DIST_OFF + (1 << ROOT_DIST);
  static constexpr int LANE_U16 = POOL_OFF + POOL;  // decode LUT entries per lane
  static constexpr int WARP_U16 = LANE_U16 * 32;
  static constexpr int WARP_BYTES = WARP_U16 * 2 + RING_WARP_BYTES;
  static constexpr int INFO_WORDS = 32;             // shared distance info table
  static constexpr int SMEM_BYTES = WARPS * WARP_BYTES + INFO_WORDS * 4;
  static constexpr uint32_t DEFER_LOST = 64;        // 2^-9 of the code space, in units of 2^-15
  static_assert(POOL <= 256, "sub-table offsets are 8 bits");
  static_assert(POOL >= 64, "the 128-entry CL LUT (one byte each) is overlaid on the pool");
};

// LUT entry (u16).  bits 0-3 = code length L (1..15); L == 0 marks a special entry.
//   L != 0, literal/length table: bits 4-11 = literal byte, or (length base - 3) with
//           bit 15 set and bits 12-14 = number of extra bits (literals keep bits 12-15 clear)
//   L != 0, distance table: bits 4-8 = distance symbol (0..29), bits 9-12 = its number of extra
//           bits (so that the bit count of a token does not wait for the base/extra table)
//   L == 0: 0x0000 no code here;
//           bit 15 set: sub-table pointer, bits 4-11 offset from the pool start,
//                       bits 12-14 sub-table index bits (1..7);
//           E_SLOW: "ask slow_token()" — end of block, symbols 286/287/30/31 (which the
//                   reference rejects after decoding them), codes the pool had no room for,
//                   and every slot of an over-subscribed code set.
constexpr uint32_t E_SLOW = 0x0010u;
constexpr int SUB_BITS_MAX = 5;
constexpr uint32_t E_PTR = 0x8000u;

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
// Shared-memory addresses of the input ring: 32-bit shared-window addresses on the device (so
// that ring accesses are plain LDS / cp.async with no generic-address arithmetic), ordinary
// pointers in the CPU build of tests/cpu_emu.
#ifdef SFB_CPU_EMU
using saddr_t = uint8_t*;
__device__ __forceinline__ saddr_t to_saddr(void* p) { return static_cast<uint8_t*>(p); }
__device__ __forceinline__ uint32_t lds32(saddr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
__device__ __forceinline__ uint32_t lds16(saddr_t a) { return *reinterpret_cast<const uint16_t*>(a); }
#else
using saddr_t = uint32_t;
__device__ __forceinline__ saddr_t to_saddr(void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lds32(saddr_t a)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds16(saddr_t a)
{
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
#endif

// cp.async staging (global -> shared, 16 bytes, L2 only).  `src_size` < 16 zero-fills the rest
// of the block, which is exactly the "bits past the end read as zero" rule of the bit reader,
// and means no byte past a stream's end is ever read.
__device__ __forceinline__ void cp_async16(saddr_t smem_dst, const void* gsrc, uint32_t src_size)
{
#ifdef SFB_CPU_EMU
  uint8_t tmp[16] = {0};
  if (src_size) std::memcpy(tmp, gsrc, src_size);
  std::memcpy(smem_dst, tmp, 16);
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc),
               "r"(src_size)
               : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit()
{
#ifndef SFB_CPU_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
#ifndef SFB_CPU_EMU
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------
// Bit reader (replaces huffman::bit_span, huffman/src/bit_span.hpp).  Bits are consumed LSB
// first.  The compressed bytes are staged by cp.async into a per-lane ring of RING_BLOCKS
// 16-byte blocks in shared memory; w0:w1:w2 is a 96-bit register window over the stream and
// `bo` the offset of the next unread bit inside it.  With bo < 32 at the start of a token the
// window holds >= 65 unread bits, more than the longest token (15 + 5 + 15 + 13 = 48), so a
// token is decoded with two funnel-shift peeks and the window slides once per token, by 0, 1
// or 2 words: the two candidate words are read from the ring at the START of the token
// (next2(): their addresses do not depend on the token) and selected at its end (norm2()).
//
// Words are numbered from `base`, the stream start rounded down to 16 bytes; word x lives in
// ring block (x >> 2) & 3.  Block b may be issued once block b-4 has been consumed, i.e. while
// b <= (rp >> 2) + 3.  Two disciplines keep readers behind completed copies:
//   * synchronous (header parsing, seeks): every eligible block is issued and waited for;
//   * token loop: once per two tokens at most one block is issued and committed, and
//     cp.async.wait_group 2 follows.  Two tokens move rp by at most one block, so a block is
//     first read at least two such steps after it was issued, when its group has been waited
//     for; typically it was issued ~25 tokens earlier and the wait is free.
struct BitReader {
  uint32_t w0, w1, w2;
  uint32_t bo;
  uint32_t rp;          // index of the next word to take from the ring (w2 is word rp-1)
  uint32_t pfb;         // index of the next 16-byte block to stage
  uint32_t iend;        // byte offset (from base) one past the last stream byte
  uint32_t ebits;       // 8 * iend + 96 (mod 2^32): real bits left at a token start are
                        // ebits - 32 * rp - bo whenever the window touches the end of the stream
  uint32_t lead0;       // stream start - base (0..15)
  const uint8_t* base;  // stream start rounded down to 16 bytes
  saddr_t ring;         // this lane's block 0 in shared memory (block s at ring + s*512)

  __device__ __forceinline__ const uint8_t* begin() const { return base + lead0; }

  // ring offset of word x, and of the word after the one at offset o
  static __device__ __forceinline__ uint32_t ring_off(uint32_t x)
  {
    return ((x & 12u) << 7) | ((x & 3u) << 2);
  }
  static __device__ __forceinline__ uint32_t ring_next(uint32_t o)
  {
    return ((o | 0x1F0u) + 4u) & 0x60Cu;  // the carry out of bits 2-3 ripples into bits 9-10
  }
  __device__ __forceinline__ uint32_t ring_word(uint32_t x) const { return lds32(ring + ring_off(x)); }

  __device__ __forceinline__ void issue_block()
  {
    const uint32_t at = pfb << 4;
    const uint32_t left = at < iend ? iend - at : 0u;
    // (a block wholly past the end is zero-filled without touching memory; keep its address valid)
    cp_async16(ring + ((pfb & 3u) << 9), base + (left ? at : 0u), left < 16u ? left : 16u);
    ++pfb;
  }
  __device__ __forceinline__ void fill_sync()
  {
    while (pfb <= (rp >> 2) + 3u) issue_block();
    cp_async_commit();
    cp_async_wait<0>();
  }
  // token-loop staging step (see above): call once per two tokens
  __device__ __forceinline__ void stage_step()
  {
    if (pfb <= (rp >> 2) + 3u) issue_block();
    cp_async_commit();
    cp_async_wait<2>();
  }

  __device__ __forceinline__ void open(const uint8_t* begin_, uint32_t len, saddr_t ring_)
  {
    ring = ring_;
    lead0 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(begin_) & 15u);
    base = begin_ - lead0;
    iend = lead0 + len;
    ebits = 8u * iend + 96u;
    init_at(lead0, 0);
  }
  // a reader that owns no stream: never slides, never stages
  __device__ __forceinline__ void park(saddr_t ring_)
  {
    ring = ring_;
    w0 = w1 = w2 = 0;
    bo = 0;
    rp = 3;
    pfb = 0xfffffff0u;
    iend = lead0 = 0;
    ebits = 96u;
    base = nullptr;
  }

  // start reading at byte offset `at` from base (lead0 <= at <= iend), then skip 0..7 bits
  __device__ __forceinline__ void init_at(uint32_t at, unsigned skip_bits)
  {
    const uint32_t x0 = at >> 2;
    // copies still in flight from before the jump target the same ring slots: let them land
    // first, or one of them could overwrite a block staged below
    cp_async_commit();
    cp_async_wait<0>();
    rp = x0;
    pfb = x0 >> 2;
    fill_sync();
    w0 = ring_word(x0);
    w1 = ring_word(x0 + 1);
    w2 = ring_word(x0 + 2);
    rp = x0 + 3;
    fill_sync();
    bo = 8 * (at & 3u) + skip_bits;
    norm();
  }

  // slide by one word (synchronous discipline)
  __device__ __forceinline__ void slide1()
  {
    w0 = w1;
    w1 = w2;
    w2 = ring_word(rp);
    ++rp;
    bo -= 32;
    if ((rp & 3u) == 0) fill_sync();
  }
  // make bo < 32 again, any bo (header parsing)
  __device__ __forceinline__ void norm()
  {
    while (bo >= 32) slide1();
  }
  // the two words that may enter the window at the end of this token (token loop)
  __device__ __forceinline__ void next2(uint32_t& n0, uint32_t& n1) const
  {
    const uint32_t o0 = ring_off(rp);
    n0 = lds32(ring + o0);
    n1 = lds32(ring + ring_next(o0));
  }
  // make bo < 32 again for bo < 96, branch-free (token loop; staging is the caller's job)
  __device__ __forceinline__ void norm2(uint32_t n0, uint32_t n1)
  {
    const bool p1 = bo >= 32, p2 = bo >= 64;
    w0 = p2 ? w2 : (p1 ? w1 : w0);
    w1 = p2 ? n0 : (p1 ? w2 : w1);
    w2 = p2 ? n1 : (p1 ? n0 : w2);
    rp += bo >> 5;
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
  // (w0 is word rp-3; the window may start before the stream right after init_at)
  __device__ __forceinline__ int64_t bitpos() const
  {
    return 32ll * (static_cast<int64_t>(rp) - 3) - 8ll * static_cast<int64_t>(lead0) +
           static_cast<int64_t>(bo);
  }
  __device__ __forceinline__ uint64_t total_bits() const
  {
    return 8ull * static_cast<uint64_t>(iend - lead0);
  }
  // real (inside the stream) bits from the next unread bit on; negative after an overrun
  __device__ __forceinline__ int64_t real_left() const
  {
    return static_cast<int64_t>(total_bits()) - bitpos();
  }
  // some word already in the window reaches past the end of the stream
  __device__ __forceinline__ bool tail() const { return 4u * rp > iend; }
  // continue at bit `bit` of the stream (synchronous staging)
  __device__ __forceinline__ void seek_bit(uint64_t bit)
  {
    init_at(lead0 + static_cast<uint32_t>(bit >> 3), static_cast<unsigned>(bit & 7));
  }
};

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
// `bits` holds the next stream bits (LSB first; zero past the end of the stream), `avail`
// how many of them are real.  Returns the code length (0 = not found).
// The 30 table words are loaded up front (independent loads: one L2 round trip, not fifteen).
__device__ __noinline__ int canon_decode(const uint32_t* scratch, int which, uint32_t bits,
                                         int64_t avail, int* symbol)
{
  const uint32_t* first = scratch + (SCR_FIRST + 16 * which) * 32;
  const uint32_t* cnt_off = scratch + (SCR_COUNT + 16 * which) * 32;
  uint32_t f[16], c[16];
#pragma unroll
  for (int L = 0; L < 16; ++L) {
    f[L] = first[L * 32];
    c[L] = L ? cnt_off[L * 32] : 0u;
  }
  const uint32_t maxlen = f[0];  // slot 0 holds the longest code length of the table
  uint32_t code = 0;
  int found = 0;
  uint32_t index = 0;
#pragma unroll
  for (int L = 1; L < 16; ++L) {
    code = (code << 1) | ((bits >> (L - 1)) & 1u);
    const uint32_t k = code - f[L];
    const bool ok = !found && static_cast<uint32_t>(L) <= maxlen && static_cast<int64_t>(L) <= avail &&
                    k < (c[L] & 0xffffu);
    if (ok) {
      found = L;
      index = (which ? 288u : 0u) + (c[L] >> 16) + k;
    }
  }
  if (found) {
    const uint32_t w = scratch[(SCR_SORTED + (index >> 1)) * 32];
    *symbol = static_cast<int>((index & 1u) ? (w >> 16) : (w & 0xffffu));
  }
  return found;
}

struct SlowToken {
  int status;     // ST_SUCCESS if a token was decoded
  int kind;       // 0 literal, 1 end of block, 2 match
  int value;      // literal byte or match length
  int dist;
  uint32_t used;  // bits the token occupies
};

// One token, exactly as decompress_block_huffman's loop body does it
// (src/decompress.cpp:206-240 with decode_lit_or_len :122-144 and
// decompress_length_distance :157-177), bit-serially.  `snap` = the 64 stream bits from the
// token start (zero past the end), `avail` = how many real bits remain from there (<= 0 when
// the token starts at or past the end).  A token is at most 48 bits long.  Output-side checks
// (distance > written, room) stay with the caller.
__device__ __noinline__ SlowToken slow_token(const uint32_t* lens, uint64_t snap, int64_t avail)
{
  SlowToken t;
  t.status = ST_SUCCESS;
  t.kind = 0;
  t.value = 0;
  t.dist = 0;
  t.used = 0;
  if (avail < 0) avail = 0;
  int sym = 0;
  uint32_t pos = 0;
  int used = canon_decode(lens, 0, static_cast<uint32_t>(snap), avail, &sym);
  if (!used) {
    t.status = ST_INVALID_LIT_OR_LEN;
    return t;
  }
  pos += static_cast<uint32_t>(used);
  if (sym < 256) {
    t.value = sym;
    t.used = pos;
    return t;
  }
  if (sym == 256) {
    t.kind = 1;
    t.used = pos;
    return t;
  }
  if (sym > 285) {
    t.status = ST_INVALID_LIT_OR_LEN;
    return t;
  }
  t.kind = 2;
  uint32_t info = c_len_info[sym - 257];
  uint32_t extra = info >> 16;
  if (avail - static_cast<int64_t>(pos) < static_cast<int64_t>(extra)) {  // reference: unchecked pop_bits (class U)
    t.status = ST_SRC_TOO_SMALL;
    return t;
  }
  t.value = static_cast<int>((info & 0xffffu) + (static_cast<uint32_t>(snap >> pos) & ((1u << extra) - 1u)));
  pos += extra;
  used = canon_decode(lens, 1, static_cast<uint32_t>(snap >> pos), avail - static_cast<int64_t>(pos), &sym);
  if (!used) {
    t.status = ST_INVALID_DISTANCE;
    return t;
  }
  pos += static_cast<uint32_t>(used);
  if (sym >= 30) {
    t.status = ST_INVALID_LIT_OR_LEN;  // sic: src/decompress.cpp:171-173
    return t;
  }
  info = c_dist_info[sym];
  extra = info >> 16;
  if (avail - static_cast<int64_t>(pos) < static_cast<int64_t>(extra)) {
    t.status = ST_SRC_TOO_SMALL;
    return t;
  }
  t.dist = static_cast<int>((info & 0xffffu) + (static_cast<uint32_t>(snap >> pos) & ((1u << extra) - 1u)));
  pos += extra;
  t.used = pos;
  return t;
}

// ---------------------------------------------------------------------------------------------
// Build one two-level LUT from code lengths [s0, s0+n).  Canonical code assignment as in
// huffman::table::canonicalize (huffman/src/table.hpp:177-216).  Length sets the reference
// accepts but a prefix LUT cannot represent exactly (over-subscribed: some code value reaches
// 2^len, "shortest code wins") get E_SLOW in every root slot.
// LITLEN selects the entry payload (see the entry format above).
// ---------------------------------------------------------------------------------------------
// Codes longer than the LUT root that found no slot in the sub-table pool (alphabets with more
// symbols than the tables have entries: real HTML has 240) are still decoded in registers: per
// length, first code and count — the canonical form — kept by the lane, one compare per length,
// and the symbol fetched from the sorted list in the lane's scratch.  Only valid for code sets
// that are not over-subscribed (prefix-free), where "no code of <= ROOT bits matches" follows
// from the root slot being a pointer / pool marker; anything it does not find goes to slow_token().
template <int ROOT>
struct LongTab {
  uint32_t fc[15 - ROOT];  // lengths ROOT+1 .. 15: first code | count << 16
  uint32_t off0;           // symbols with shorter codes (rank of the first one of length ROOT+1)
  bool usable;             // false for over-subscribed sets
};

// -> code length (0: none) and rank of the symbol in canonical order
template <int ROOT>
__device__ __forceinline__ uint32_t long_decode(const LongTab<ROOT>& t, uint32_t bits, uint32_t& rank)
{
  const uint32_t rev = __brev(bits);  // codes are packed MSB first
  uint32_t found = 0, off = t.off0;
#pragma unroll
  for (int i = 0; i < 15 - ROOT; ++i) {
    const uint32_t L = ROOT + 1 + i;
    const uint32_t k = (rev >> (32u - L)) - (t.fc[i] & 0xffffu);
    const uint32_t cnt = t.fc[i] >> 16;
    if (found == 0 && k < cnt) {
      found = L;
      rank = off + k;
    }
    off += cnt;
  }
  return found;
}

template <bool LITLEN>
__device__ __forceinline__ uint16_t make_entry(uint32_t s, uint32_t L)
{
  if (!LITLEN) return static_cast<uint16_t>(s < 30 ? (((c_dist_info[s] >> 16) << 9) | (s << 4) | L) : E_SLOW);
  if (s < 256) return static_cast<uint16_t>((s << 4) | L);
  if (s == 256 || s > 285) return static_cast<uint16_t>(E_SLOW);
  const uint32_t info = c_len_info[s - 257];
  return static_cast<uint16_t>(0x8000u | ((info >> 16) << 12) | (((info & 0xffffu) - 3u) << 4) | L);
}

template <int ROOT, bool LITLEN, int POOL_OFF>
__device__ __forceinline__ void build_lut(const LaneMem& m, int s0, int n, int root_off, int pool_end,
                          int& pool_at, uint32_t& lost, LongTab<ROOT>& lt)
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
      if (L > ROOT) lt.fc[L - ROOT - 1] = (code & 0xffffu) | (static_cast<uint32_t>(count[L]) << 16);
      if (L == ROOT + 1) lt.off0 = offset[L];
    }
    g_first[0] = maxlen;
    lt.usable = !over;
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
  // pass B: carve sub-tables out of the pool (none when there are no long codes: pmax < pmin).
  // A sub-table indexes at most SUB_BITS_MAX bits — the deepest prefix of a text-like code
  // would otherwise take 2^(15-ROOT) slots and starve the rest — and shrinks further to fit
  // what is left of the pool; codes it is too short for keep "no code" slots and are decoded,
  // exactly, by slow_token() (they are the rarest symbols of the block).
#pragma unroll 1
  for (uint32_t pfx = pmin; pfx <= pmax; ++pfx) {
    uint16_t& slot = lut[(root_off + bitrev(pfx, ROOT)) * 32];
    const uint32_t e = slot;
    if ((e & 0xF000u) != 0xF000u) continue;
    int sb = static_cast<int>(e & 15u) - ROOT;
    if (sb > SUB_BITS_MAX) sb = SUB_BITS_MAX;
    while (sb > 0 && pool_at + (1 << sb) > pool_end) --sb;
    const int size = 1 << sb;
    if (sb > 0) {
      slot = static_cast<uint16_t>(E_PTR | (static_cast<uint32_t>(sb) << 12) |
                                   (static_cast<uint32_t>(pool_at - POOL_OFF) << 4));
      for (int j = 0; j < size; ++j) lut[(pool_at + j) * 32] = 0;
      pool_at += size;
    } else {
      slot = static_cast<uint16_t>(E_SLOW);  // pool exhausted: these codes go through slow_token()
    }
  }
  // pass C: fill sub-tables
  if (pmax >= pmin) {
    uint32_t code = 0;
#pragma unroll
    for (int L = 1; L < 16; ++L) {
      code = (code + count[L - 1]) << 1;
      next[L] = code;
    }
    LensReader r(m.lens, s0, n);
#pragma unroll 1
    for (int s = 0; s < n; ++s) {
      const uint32_t L = r.next();
      if (L <= static_cast<uint32_t>(ROOT)) continue;  // next[] of short lengths is not needed
      const uint32_t code = next[L]++;
      const uint32_t rest = L - ROOT;
      const uint32_t e = lut[(root_off + bitrev(code >> rest, ROOT)) * 32];
      // codes the tables cannot hold go through slow_token(); `lost` adds up their share of
      // the code space (2^-L in units of 2^-15) as an estimate of how often that will happen
      const bool is_eob = LITLEN && s == 256;  // (the end-of-block symbol takes the slow path anyway)
      if (e == E_SLOW) {
        if (!is_eob) lost += 1u << (15u - L);
        continue;
      }
      const uint32_t sb = (e >> 12) & 7u;
      if (rest > sb) {  // longer than its (capped) sub-table reaches
        if (!is_eob) lost += 1u << (15u - L);
        continue;
      }
      const uint32_t off = POOL_OFF + ((e >> 4) & 0xffu);
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
  if ((e & (E_PTR | 15u)) == E_PTR) {  // sub-table pointer (codes longer than ROOT)
    // (lanes that are not decoding run this on stale table contents and discard the result:
    //  the clamp keeps even a garbage pointer inside this lane's slice)
    const uint32_t sb = (e >> 12) & 7u;
    const uint32_t idx = ((e >> 4) & 0xffu) + ((bits >> ROOT) & ((1u << sb) - 1u));
    e = lut[(POOL_OFF + (idx < POOL ? idx : POOL - 1u)) * 32];
  }
  return e;
}

// The same lookup for the token loop, on shared-window addresses: `lutb` is the address of this
// lane's element 0, element j sits 64*j bytes further.  Branch-free: the second-level load is
// always made (some lane of the warp needs it on practically every token anyway) and selected.
template <int ROOT, int ROOT_OFF, int POOL_OFF, uint32_t POOL>
__device__ __forceinline__ uint32_t lut_lookup_s(saddr_t lutb, uint32_t bits)
{
  const uint32_t e = lds16(lutb + (ROOT_OFF * 64) + ((bits << 6) & (((1u << ROOT) - 1u) << 6)));
  const uint32_t sb = (e >> 12) & 7u;
  uint32_t idx = ((e >> 4) & 0xffu) + ((bits >> ROOT) & ~(0xffffffffu << sb));
  idx = idx < POOL ? idx : POOL - 1u;  // (a direct entry is not a pointer: stay inside the slice)
  const uint32_t e2 = lds16(lutb + (POOL_OFF * 64) + (idx << 6));
  return (e & (E_PTR | 15u)) == E_PTR ? e2 : e;
}

// lane states
enum : int { S_DECODE = 0, S_STORED = 1, S_HEADER = 2, S_DONE = 3 };

// ---------------------------------------------------------------------------------------------
// Block header, lane-local.  Restates read_header (src/decompress.cpp:370-385), the stored
// branch (:416-436) and decode_dynamic_huffman_tables / decode_dynamic_huffman_table
// (:253-367).  Every field is bound-checked here (this is not the hot loop); where the
// reference would read past the input (assert-only pop_bits / pop_16) the documented choices
// of this repository apply (SrcTooSmall; InvalidLitOrLen for malformed repeats).
// Returns the next lane state; *status is set when the state is S_DONE.
// Control flow: ONE exit and no jumps out of loops.  With early returns inside the loops the
// only reconvergence point the compiler can use is the end of this (inlined) function, and the
// lanes of a warp, which all parse their headers at the same time, would run the whole parse
// and both table builds one lane after the other (ncu: 1.2 active threads per instruction).
template <class C>
__device__ __forceinline__ int parse_block_header(BitReader& br, const LaneMem& m, uint32_t room,
                                                  uint32_t& final_block, int& n_lit, int& n_dist,
                                                  const uint8_t*& copy_src, uint32_t& copy_left,
                                                  int* status, uint32_t* lost_out,
                                                  LongTab<C::ROOT_LIT>& lt_lit,
                                                  LongTab<C::ROOT_DIST>& lt_dist, bool& tables_fixed)
{
  *lost_out = 0;
  int err = -1;           // DecompressStatus once the stream is known to end here
  int next = S_DECODE;
  uint32_t type = 3;
  br.norm();
  if (br.real_left() < 3) {
    err = ST_INVALID_BLOCK_HEADER;
  } else {
    const uint32_t hdr = br.peek() & 7u;
    type = hdr >> 1;
    if (type == 3) {
      err = ST_INVALID_BLOCK_HEADER;
    } else {
      final_block = hdr & 1u;
      br.skip(3);
    }
  }

  if (err < 0 && type == 0) {
    // stored block: skip to the byte boundary (the window words are byte-aligned with the
    // stream, so the phase of `bo` is the phase of the stream)
    br.skip((8u - (br.bo & 7u)) & 7u);
    br.norm();
    if (br.real_left() < 32) {
      err = ST_SRC_TOO_SMALL;  // reference: pop_16 past the end (class U)
    } else {
      const uint32_t w = br.peek();
      br.skip(32);
      const uint32_t len = w & 0xffffu, nlen = w >> 16;
      const uint64_t pos = static_cast<uint64_t>(br.bitpos());
      if (len != ((~nlen) & 0xffffu)) {
        err = ST_LEN_MISMATCH;
      } else if (br.total_bits() - pos < 8ull * len) {
        err = ST_SRC_TOO_SMALL;
      } else if (room < len) {
        err = ST_DST_TOO_SMALL;
      } else {
        copy_src = br.begin() + (pos >> 3);
        copy_left = len;
        next = S_STORED;
        if (len == 0) {
          br.norm();
          next = S_HEADER;
          if (final_block) err = ST_SUCCESS;
        }
      }
    }
  } else if (err < 0) {
    // `tables_fixed`: this lane's LUTs, scratch arrays and LongTabs already hold the fixed codes
    // (the previous block it decoded — of this or of an earlier stream — was a fixed one):
    // nothing to build.  Batches sorted by type give whole warps of such streams.
    const bool reuse = type == 1 && tables_fixed;
    if (type == 1) {
      // fixed codes: src/decompress.cpp:25-40
      n_lit = 288;
      n_dist = 32;
      if (!reuse) {
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
      }
    } else {
      // dynamic codes (from here on the lane's tables and scratch are being overwritten)
      tables_fixed = false;
      br.norm();
      int n_cl = 0;
      if (br.real_left() < 14) {
        err = ST_SRC_TOO_SMALL;
      } else {
        const uint32_t w = br.peek();
        n_lit = 257 + static_cast<int>(w & 31u);
        n_dist = 1 + static_cast<int>((w >> 5) & 31u);
        n_cl = 4 + static_cast<int>((w >> 10) & 15u);
        br.skip(14);
      }
      uint64_t cl_lens = 0;  // 19 x 3 bits, indexed by CL symbol
#pragma unroll 1
      for (int i = 0; i < n_cl && err < 0; ++i) {
        br.norm();
        if (br.real_left() < 3) {
          err = ST_SRC_TOO_SMALL;
        } else {
          cl_lens |= static_cast<uint64_t>(br.peek() & 7u) << (3 * c_cl_order[i]);
          br.skip(3);
        }
      }
      if (err < 0) {
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
              if (static_cast<int>((cl_lens >> (3 * s)) & 7u) == L) {
                if (code < (1u << L)) {
                  for (uint32_t j = bitrev(code, L); j < 128u; j += 1u << L)
                    cl_lut[(j >> 1) * 64 + (j & 1)] = static_cast<uint8_t>((s << 3) | L);
                }
                ++code;
              }
            }
          }
        }
        // The HLIT + HDIST code lengths are ONE sequence (RFC 1951 §3.2.7: a repeat may run from
        // the literal/length lengths into the distance lengths, and libdeflate, zopfli and 7-zip
        // write such headers).  The reference decodes two independent runs
        // (src/decompress.cpp:353-360) and has undefined behaviour exactly where the two readings
        // differ — a repeat that crosses the boundary writes past its array, a 16 as the first
        // distance length reads before it — so every input on which it is defined decodes the
        // same here.
        const int total = n_lit + n_dist;
        LensWriter lw(m.lens);
        uint32_t prev = 0;
#pragma unroll 1
        for (int i = 0; i < total && err < 0;) {
          br.norm();
          const uint32_t bits = br.peek();
          const uint32_t j = bits & 127u;
          const uint32_t e = cl_lut[(j >> 1) * 64 + (j & 1)];
          const uint32_t L = e & 7u;
          const int64_t left = br.real_left();
          const uint32_t sym = e >> 3;
          const uint32_t xbits = sym < 16 ? 0u : sym == 16 ? 2u : sym == 17 ? 3u : 7u;
          // L + xbits <= 14 bits: still inside the peeked word
          int repeat = 1;
          uint32_t v = sym;
          if (sym >= 16) {
            repeat = static_cast<int>((bits >> L) & ((1u << xbits) - 1u)) + (sym == 18 ? 11 : 3);
            v = sym == 16 ? prev : 0u;
          }
          if (L == 0 || static_cast<int64_t>(L) > left) {
            err = ST_INVALID_LIT_OR_LEN;  // src/decompress.cpp:265-267
          } else if (left - static_cast<int64_t>(L) < static_cast<int64_t>(xbits)) {
            err = ST_SRC_TOO_SMALL;       // reference: unchecked pop_bits (class U)
          } else if (sym == 16 && i == 0) {
            err = ST_INVALID_LIT_OR_LEN;  // no previous length to repeat (reference: reads code_bitsizes[-1], class U)
          } else if (i + repeat > total) {
            err = ST_INVALID_LIT_OR_LEN;  // more lengths than HLIT + HDIST (reference: writes past code_bitsizes, class U)
          } else {
            br.skip(L + xbits);
#pragma unroll 1
            for (int r = 0; r < repeat; ++r) lw.put(v);
            prev = v;
            i += repeat;
          }
        }
        if (err < 0) lw.finish();
      }
    }
    if (err < 0 && reuse) br.norm();
    if (err < 0 && !reuse) {
      tables_fixed = type == 1;
      int pool_at = C::POOL_OFF;
      uint32_t lost = 0;
      build_lut<C::ROOT_LIT, true, C::POOL_OFF>(m, 0, n_lit, C::LIT_OFF, C::POOL_OFF + C::POOL, pool_at,
                                                lost, lt_lit);
      build_lut<C::ROOT_DIST, false, C::POOL_OFF>(m, n_lit, n_dist, C::DIST_OFF, C::POOL_OFF + C::POOL,
                                                  pool_at, lost, lt_dist);
      *lost_out = lost;
      br.norm();
    }
  }
  if (err >= 0) {
    *status = err;
    next = S_DONE;
  }
  return next;
}

}  // namespace sfb
