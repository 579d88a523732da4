/* ===========================================================================
 * TEST INFRASTRUCTURE ONLY — NOT PART OF THE PRODUCT.
 *
 * Plain-C restatement of the reference's raw-DEFLATE decompressor, used as the
 * parity oracle for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product
 * library (starflate_b200/) never does and has no CPU fallback.
 *
 * Every function cites the reference code it restates (paths relative to
 * /root/reference).  The algorithm is kept deliberately close to the
 * reference's: bit-at-a-time canonical decode, first error wins, bytes already
 * produced stay in dst, dst past the write cursor is never touched.
 *
 * Parity status: PINNED.  oracle/pin_oracle.py checks this file against the
 * unmodified reference built by oracle/Makefile (NDEBUG build for status and
 * bytes, ASan+UBSan build for the defined/undefined split) on every reference
 * test vector, every SURVEY.md §8(c) known answer, and truncation / mutation
 * sweeps; the resulting fixtures are committed under tests/golden/.
 *
 * Inputs on which the reference itself has undefined behaviour (it relies on
 * assert-only preconditions: pop_bits / pop_16 past the end of input,
 * code-length repeat before the first length or past the last) are reported
 * with ref_undefined = 1 and this repository's documented status choice:
 *   - running out of input inside a fixed-width field  -> SrcTooSmall
 *   - code-length repeat with nothing to repeat, or past
 *     the HLIT + HDIST lengths the header announced     -> InvalidLitOrLen
 *   - a repeat that runs from the lit/len lengths into
 *     the distance lengths (valid per RFC 1951)         -> decoded as the RFC says
 * =========================================================================== */
#include "inflate_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAX_BITS 15
#define MAX_SYMS 288

/* ---- bit reader: huffman::bit_span (huffman/src/bit_span.hpp:18-183) -------
 * Bit i of the stream is bit (i % 8), LSB first, of byte i / 8 (:46-53). */
typedef struct {
  const uint8_t* data;
  uint64_t pos; /* next unread bit */
  uint64_t end; /* one past the last bit */
} bits_t;

static inline unsigned bit_at(const uint8_t* data, uint64_t i)
{
  return (data[i >> 3] >> (i & 7)) & 1u;
}

static inline uint64_t bits_left(const bits_t* b) { return b->end - b->pos; }

/* pop_bits<T> (src/decompress.cpp:94-114): n <= 16 bits, LSB first.  The
 * reference does not bound-check; *ub is raised where it would read past the
 * end. */
static unsigned pop_bits(bits_t* b, unsigned n, int* ub)
{
  unsigned res = 0;
  if (bits_left(b) < n) {
    *ub = 1;
    return 0;
  }
  for (unsigned i = 0; i < n; i++) res |= bit_at(b->data, b->pos + i) << i;
  b->pos += n;
  return res;
}

/* ---- canonical table: huffman::table symbol_bitsize ctor ------------------
 * (huffman/src/table.hpp:360-376) -> canonicalize (:177-216).
 * The reference keeps a sorted vector of nodes plus per-node skip counts; the
 * equivalent compact form used here is, per bitsize L, the first code value,
 * the number of codes and the position of the first one in the sorted order. */
typedef struct {
  uint64_t first[MAX_BITS + 1]; /* code value of the first code of bitsize L */
  uint32_t count[MAX_BITS + 1];
  uint32_t start[MAX_BITS + 1]; /* index into order[] */
  uint16_t order[MAX_SYMS];     /* symbols sorted by (bitsize, symbol) */
  unsigned max_bits;            /* 0 = empty table */
} ctable_t;

static void ctable_build(ctable_t* t, const uint8_t* lens, size_t n)
{
  memset(t, 0, sizeof *t);
  for (size_t s = 0; s < n; s++)
    if (lens[s]) t->count[lens[s]]++;
  /* sort by (bitsize, symbol): table.hpp:182-187 */
  uint32_t at = 0;
  for (unsigned L = 1; L <= MAX_BITS; L++) {
    t->start[L] = at;
    at += t->count[L];
    if (t->count[L]) t->max_bits = L;
  }
  uint32_t fill[MAX_BITS + 1];
  memcpy(fill, t->start, sizeof fill);
  for (size_t s = 0; s < n; s++)
    if (lens[s]) t->order[fill[lens[s]]++] = (uint16_t)s;
  /* code assignment: table.hpp:189-210.  base_code counts every element seen
   * so far in the scale of the current bitsize; a new bitsize shifts it. */
  uint64_t base_code = 0;
  unsigned cur_bits = 0;
  for (unsigned L = 1; L <= MAX_BITS; L++) {
    if (!t->count[L]) continue;
    base_code <<= (L - cur_bits);
    cur_bits = L;
    t->first[L] = base_code;
    base_code += t->count[L];
  }
}

/* decode_one (huffman/src/decode.hpp:83-102) with table::find
 * (huffman/src/table.hpp:426-452): accumulate bits MSB-first; after each bit a
 * code of the current bitsize matches iff (value - first) <u count; stop with
 * "not found" when the input is exhausted or the bitsize reached the longest
 * code in the table.  Does not consume. */
static int ctable_decode_one(const ctable_t* t, const uint8_t* data,
                             uint64_t pos, uint64_t end, uint16_t* symbol)
{
  uint64_t value = 0;
  for (unsigned len = 1; pos + (len - 1) < end; len++) {
    if (len > t->max_bits) break; /* find() ran off the table: decode.hpp:96 */
    value = (value << 1) | bit_at(data, pos + (len - 1));
    if (t->count[len] && (uint64_t)(value - t->first[len]) < t->count[len]) {
      *symbol = t->order[t->start[len] + (uint32_t)(value - t->first[len])];
      return (int)len;
    }
  }
  return 0;
}

/* ---- RFC 1951 §3.2.5 tables (src/decompress.cpp:42-84) --------------------- */
static const uint8_t len_extra[28] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2,
                                      2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5};
static const uint16_t len_base[28] = {3,  4,  5,  6,  7,  8,  9,  10, 11, 13,
                                      15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                      67, 83, 99, 115, 131, 163, 195, 227};
static const uint8_t dist_extra[30] = {0, 0, 0,  0,  1,  1,  2,  2,  3,  3,
                                       4, 4, 5,  5,  6,  6,  7,  7,  8,  8,
                                       9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint16_t dist_base[30] = {
    1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
    193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
/* src/decompress.cpp:250-251 */
static const uint8_t cl_order[19] = {16, 17, 18, 0, 8,  7, 9,  6, 10, 5,
                                     11, 4,  12, 3, 13, 2, 14, 1, 15};

/* copy_from_before (src/decompress.cpp:388-398): chunked copy whose chunk is
 * min(left, dst - src) with src fixed == a forward byte-by-byte copy. */
void sfo_copy_from_before(uint8_t* buf, size_t dst_index, uint16_t distance,
                          uint16_t n)
{
  uint8_t* dst = buf + dst_index;
  const uint8_t* from = dst - distance;
  long left = n;
  while (left > 0) {
    long chunk = dst - from;
    if (chunk > left) chunk = left;
    memmove(dst, from, (size_t)chunk); /* ranges never overlap: chunk <= dst-from */
    dst += chunk;
    left -= chunk;
  }
}

/* read_header (src/decompress.cpp:370-385) */
static int read_header(bits_t* b, int* final, int* type)
{
  if (bits_left(b) < 3) return SFO_INVALID_BLOCK_HEADER;
  const int t = (int)(bit_at(b->data, b->pos + 1) | (bit_at(b->data, b->pos + 2) << 1));
  if (t == 3) return SFO_INVALID_BLOCK_HEADER;
  *final = (int)bit_at(b->data, b->pos);
  *type = t;
  b->pos += 3;
  return SFO_SUCCESS;
}

void sfo_read_header(const uint8_t* src, size_t bit_size, uint8_t bit_offset,
                     int* out)
{
  bits_t b = {src, bit_offset, (uint64_t)bit_offset + bit_size};
  int final = 0, type = 0;
  const int st = read_header(&b, &final, &type);
  out[0] = st == SFO_SUCCESS;
  out[1] = final;
  out[2] = type;
  out[3] = st;
  out[4] = (int)(b.pos - bit_offset);
}

/* decode_dynamic_huffman_table (src/decompress.cpp:253-312).  The reference decodes the
 * lit/len and the distance code lengths as two independent runs (:353-360); RFC 1951 §3.2.7
 * makes them ONE sequence of n_lit + n_dist lengths (a repeat may cross from the first into the
 * second, and real encoders — libdeflate, zopfli, 7-zip — write such headers).  The two readings
 * differ only where the reference has undefined behaviour: a repeat that crosses the boundary
 * writes past its first array, a 16 as the first distance length reads before its second.
 * Those inputs are decoded as the RFC says and flagged ref_undefined; everything the reference
 * defines comes out identically. */
static int read_code_lengths(bits_t* b, const ctable_t* cl, unsigned n_lit, unsigned n_dist,
                             uint8_t* lens, int* ub, int* ub_rfc_valid)
{
  const unsigned n_codes = n_lit + n_dist;
  memset(lens, 0, n_codes);
  for (unsigned i = 0; i < n_codes; i++) {
    uint16_t sym = 0;
    const int used = ctable_decode_one(cl, b->data, b->pos, b->end, &sym);
    if (!used) return SFO_INVALID_LIT_OR_LEN; /* :265-267 */
    b->pos += (unsigned)used;
    if (sym < 16) {
      lens[i] = (uint8_t)sym; /* :269-272 */
      continue;
    }
    unsigned repeat, value;
    if (sym == 16) { /* :273-280 */
      repeat = 3 + pop_bits(b, 2, ub);
      if (*ub) return SFO_SRC_TOO_SMALL;
      if (i == 0) { /* nothing to repeat; reads code_bitsizes[-1] in the reference */
        *ub = 1;
        return SFO_INVALID_LIT_OR_LEN;
      }
      if (i == n_lit) *ub_rfc_valid = 1; /* the reference reads before its distance array here */
      value = lens[i - 1];
    } else if (sym == 17) { /* :281-288 */
      repeat = 3 + pop_bits(b, 3, ub);
      if (*ub) return SFO_SRC_TOO_SMALL;
      value = 0;
    } else { /* 18, :289-296 */
      repeat = 11 + pop_bits(b, 7, ub);
      if (*ub) return SFO_SRC_TOO_SMALL;
      value = 0;
    }
    if (i + repeat > n_codes) { /* more lengths than the header announced */
      *ub = 1;
      return SFO_INVALID_LIT_OR_LEN;
    }
    if (i < n_lit && i + repeat > n_lit) *ub_rfc_valid = 1; /* crosses: the reference writes past its lit/len array */
    for (unsigned j = 0; j < repeat; j++) lens[i + j] = (uint8_t)value;
    i += repeat - 1;
  }
  return SFO_SUCCESS;
}

/* decode_dynamic_huffman_tables (src/decompress.cpp:314-367) */
static int read_dynamic_tables(bits_t* b, ctable_t* lit, ctable_t* dist, int* ub, int* ub_rfc_valid)
{
  const unsigned n_lit = 257 + pop_bits(b, 5, ub);
  if (*ub) return SFO_SRC_TOO_SMALL;
  const unsigned n_dist = 1 + pop_bits(b, 5, ub);
  if (*ub) return SFO_SRC_TOO_SMALL;
  const unsigned n_cl = 4 + pop_bits(b, 4, ub);
  if (*ub) return SFO_SRC_TOO_SMALL;
  uint8_t cl_lens[19] = {0};
  for (unsigned i = 0; i < n_cl; i++) {
    cl_lens[cl_order[i]] = (uint8_t)pop_bits(b, 3, ub);
    if (*ub) return SFO_SRC_TOO_SMALL;
  }
  ctable_t cl;
  ctable_build(&cl, cl_lens, 19);
  uint8_t lens[MAX_SYMS + 32];
  const int st = read_code_lengths(b, &cl, n_lit, n_dist, lens, ub, ub_rfc_valid);
  if (st != SFO_SUCCESS) return st;
  ctable_build(lit, lens, n_lit);
  ctable_build(dist, lens + n_lit, n_dist);
  return SFO_SUCCESS;
}

/* decompress_block_huffman (src/decompress.cpp:197-242) with decode_lit_or_len
 * (:122-144), decompress_literal (:146-155), decompress_length_distance
 * (:157-187). */
static int inflate_block(bits_t* b, uint8_t* dst, size_t cap, uint64_t* written,
                         const ctable_t* lit, const ctable_t* dist, int* ub)
{
  for (;;) {
    uint16_t sym = 0;
    int used = ctable_decode_one(lit, b->data, b->pos, b->end, &sym);
    if (!used) return SFO_INVALID_LIT_OR_LEN; /* :215-217 */
    b->pos += (unsigned)used;
    if (sym < 256) { /* literal */
      if (cap - *written < 1) return SFO_DST_TOO_SMALL;
      dst[(*written)++] = (uint8_t)sym;
      continue;
    }
    if (sym == 256) return SFO_SUCCESS;             /* end of block */
    if (sym > 285) return SFO_INVALID_LIT_OR_LEN;   /* :132-134 */
    unsigned len;
    if (sym == 285) {
      len = 258;
    } else {
      len = len_base[sym - 257] + pop_bits(b, len_extra[sym - 257], ub);
      if (*ub) return SFO_SRC_TOO_SMALL;
    }
    used = ctable_decode_one(dist, b->data, b->pos, b->end, &sym);
    if (!used) return SFO_INVALID_DISTANCE; /* :167-169 */
    b->pos += (unsigned)used;
    if (sym >= 30) return SFO_INVALID_LIT_OR_LEN; /* sic, :171-173 */
    const unsigned distance = dist_base[sym] + pop_bits(b, dist_extra[sym], ub);
    if (*ub) return SFO_SRC_TOO_SMALL;
    if (distance > *written) return SFO_INVALID_DISTANCE; /* :178-180 */
    if (cap - *written < len) return SFO_DST_TOO_SMALL;   /* :181-183 */
    sfo_copy_from_before(dst, (size_t)*written, (uint16_t)distance, (uint16_t)len);
    *written += len;
  }
}

static void fixed_tables(ctable_t* lit, ctable_t* dist)
{
  /* src/decompress.cpp:25-40 */
  uint8_t lens[MAX_SYMS];
  for (int s = 0; s < 288; s++) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
  ctable_build(lit, lens, 288);
  memset(lens, 5, 32);
  ctable_build(dist, lens, 32);
}

static void decompress_impl(const uint8_t* src, size_t src_len, uint8_t* dst,
                            size_t dst_cap, sfo_result* res, uint64_t* starts,
                            size_t max_starts, size_t* n_starts)
{
  bits_t b = {src, 0, (uint64_t)src_len * 8};
  uint64_t written = 0;
  int ub = 0;           /* the reference reads / writes out of bounds: decoding stops with the documented status */
  int ub_rfc_valid = 0; /* the reference is undefined, RFC 1951 is not: decoding goes on as the RFC says */
  int status = SFO_SUCCESS;
  ctable_t lit, dist;
  for (int was_final = 0; !was_final;) {
    int type = 0;
    if (starts && *n_starts < max_starts) starts[(*n_starts)++] = b.pos;
    status = read_header(&b, &was_final, &type);
    if (status != SFO_SUCCESS) break;
    if (type == 0) { /* stored: src/decompress.cpp:416-436 */
      b.pos = (b.pos + 7) & ~(uint64_t)7;
      if (bits_left(&b) < 32) { /* pop_16 x2 are assert-only in the reference */
        ub = 1;
        status = SFO_SRC_TOO_SMALL;
        break;
      }
      const uint8_t* p = src + (b.pos >> 3);
      const unsigned len = p[0] | ((unsigned)p[1] << 8);
      const unsigned nlen = p[2] | ((unsigned)p[3] << 8);
      b.pos += 32;
      if (len != ((~nlen) & 0xffffu)) {
        status = SFO_NO_COMPRESSION_LEN_MISMATCH;
        break;
      }
      if (bits_left(&b) < (uint64_t)len * 8) {
        status = SFO_SRC_TOO_SMALL;
        break;
      }
      if (dst_cap - written < len) {
        status = SFO_DST_TOO_SMALL;
        break;
      }
      memcpy(dst + written, src + (b.pos >> 3), len);
      b.pos += (uint64_t)len * 8;
      written += len;
    } else {
      if (type == 1) {
        fixed_tables(&lit, &dist);
      } else {
        status = read_dynamic_tables(&b, &lit, &dist, &ub, &ub_rfc_valid);
        if (status != SFO_SUCCESS) break;
      }
      status = inflate_block(&b, dst, dst_cap, &written, &lit, &dist, &ub);
      if (status != SFO_SUCCESS) break;
    }
  }
  res->status = (uint8_t)status;
  res->ref_undefined = (uint8_t)(ub | ub_rfc_valid);
  res->written = written;
  res->bits_consumed = b.pos;
}

void sfo_decompress(const uint8_t* src, size_t src_len, uint8_t* dst,
                    size_t dst_cap, sfo_result* res)
{
  decompress_impl(src, src_len, dst, dst_cap, res, NULL, 0, NULL);
}

size_t sfo_block_starts(const uint8_t* src, size_t src_len, uint8_t* dst,
                        size_t dst_cap, sfo_result* res, uint64_t* starts,
                        size_t max_starts)
{
  size_t n = 0;
  decompress_impl(src, src_len, dst, dst_cap, res, starts, max_starts, &n);
  return n;
}

size_t sfo_canonical_codes(const uint8_t* lens, size_t n, uint64_t* codes,
                           uint16_t* order)
{
  ctable_t t;
  ctable_build(&t, lens, n);
  size_t total = 0;
  for (unsigned L = 1; L <= MAX_BITS; L++) {
    for (uint32_t k = 0; k < t.count[L]; k++) {
      const uint16_t s = t.order[t.start[L] + k];
      codes[s] = t.first[L] + k;
      order[total++] = s;
    }
  }
  return total;
}

int sfo_decode_one(const uint8_t* lens, size_t n, const uint8_t* src,
                   uint64_t bit_pos, uint64_t bit_end, uint16_t* symbol)
{
  ctable_t t;
  ctable_build(&t, lens, n);
  return ctable_decode_one(&t, src, bit_pos, bit_end, symbol);
}

uint64_t sfo_fnv1a64(const uint8_t* p, size_t n)
{
  uint64_t h = 1469598103934665603ULL;
  for (size_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 1099511628211ULL;
  }
  return h;
}

/* ---- batched driver (CPU baseline) ------------------------- */
typedef struct {
  const uint8_t* src;
  const uint64_t *src_off, *src_len;
  uint8_t* dst;
  const uint64_t *dst_off, *dst_cap;
  uint8_t* status;
  uint64_t* written;
  uint8_t* ub;
  uint64_t lo, hi;
} job_t;

static void* job_main(void* arg)
{
  job_t* j = (job_t*)arg;
  for (uint64_t i = j->lo; i < j->hi; i++) {
    sfo_result r;
    sfo_decompress(j->src + j->src_off[i], (size_t)j->src_len[i],
                   j->dst + j->dst_off[i], (size_t)j->dst_cap[i], &r);
    j->status[i] = r.status;
    if (j->written) j->written[i] = r.written;
    if (j->ub) j->ub[i] = r.ref_undefined;
  }
  return NULL;
}

void sfo_decompress_batch(const uint8_t* src, const uint64_t* src_off,
                          const uint64_t* src_len, uint8_t* dst,
                          const uint64_t* dst_off, const uint64_t* dst_cap,
                          uint8_t* status, uint64_t* written, uint8_t* ub,
                          uint64_t n, int threads)
{
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n) threads = n ? (int)n : 1;
  pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)threads);
  const uint64_t per = (n + (uint64_t)threads - 1) / (uint64_t)threads;
  int started = 0;
  for (int t = 0; t < threads; t++) {
    const uint64_t lo = per * (uint64_t)t;
    const uint64_t hi = lo + per < n ? lo + per : n;
    if (lo >= hi) break;
    jobs[t] = (job_t){src, src_off, src_len, dst, dst_off, dst_cap, status, written, ub, lo, hi};
    pthread_create(&tid[t], NULL, job_main, &jobs[t]);
    started++;
  }
  for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
  free(jobs);
  free(tid);
}
