/* TEST INFRASTRUCTURE ONLY — see inflate_oracle.c. */
#ifndef SFB200_INFLATE_ORACLE_H
#define SFB200_INFLATE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Numeric values of starflate::DecompressStatus
 * (/root/reference/src/decompress.hpp:13-23). */
enum {
  SFO_SUCCESS = 0,
  SFO_ERROR = 1, /* never returned by the reference */
  SFO_INVALID_BLOCK_HEADER = 2,
  SFO_NO_COMPRESSION_LEN_MISMATCH = 3,
  SFO_DST_TOO_SMALL = 4,
  SFO_SRC_TOO_SMALL = 5,
  SFO_INVALID_LIT_OR_LEN = 6,
  SFO_INVALID_DISTANCE = 7
};

typedef struct sfo_result {
  uint8_t status;         /* DecompressStatus value */
  uint8_t ref_undefined;  /* 1 = the reference has undefined behaviour on this
                             input (SURVEY.md §8c class U); `status` is then
                             this repository's documented choice, not parity */
  uint64_t written;       /* bytes placed in dst (extension; reference has none) */
  uint64_t bits_consumed; /* bit position after the last consumed field */
} sfo_result;

/* Restatement of starflate::decompress
 * (/root/reference/src/decompress.cpp:402-461). */
void sfo_decompress(const uint8_t* src, size_t src_len, uint8_t* dst,
                    size_t dst_cap, sfo_result* res);

/* The same, also reporting the bit position of every block header it read (test aid for the
 * single-stream path, which decodes the blocks of one stream side by side). Returns their number. */
size_t sfo_block_starts(const uint8_t* src, size_t src_len, uint8_t* dst,
                        size_t dst_cap, sfo_result* res, uint64_t* starts,
                        size_t max_starts);

/* Batched driver: stream i = src[src_off[i] .. +src_len[i]) into
 * dst[dst_off[i] .. +dst_cap[i]); `threads` worker threads, static contiguous
 * partition. ub may be NULL. */
void sfo_decompress_batch(const uint8_t* src, const uint64_t* src_off,
                          const uint64_t* src_len, uint8_t* dst,
                          const uint64_t* dst_off, const uint64_t* dst_cap,
                          uint8_t* status, uint64_t* written, uint8_t* ub,
                          uint64_t n, int threads);

/* Restatement of detail::read_header (/root/reference/src/decompress.cpp:370-385).
 * out[0]=has_value out[1]=final out[2]=type out[3]=error status out[4]=bits consumed */
void sfo_read_header(const uint8_t* src, size_t bit_size, uint8_t bit_offset,
                     int* out);

/* Restatement of detail::copy_from_before (/root/reference/src/decompress.cpp:388-398). */
void sfo_copy_from_before(uint8_t* buf, size_t dst_index, uint16_t distance,
                          uint16_t n);

/* Canonical code assignment, restating huffman::table(symbol_bitsize, ...) +
 * canonicalize (/root/reference/huffman/src/table.hpp:177-216,360-376).
 * lens[i] = bitsize of symbol i (0 = absent). Writes code value and returns the
 * number of coded symbols; codes[i] is only meaningful where lens[i] != 0.
 * order[k] = symbol at table position k (sorted by (bitsize, symbol)). */
size_t sfo_canonical_codes(const uint8_t* lens, size_t n, uint64_t* codes,
                           uint16_t* order);

/* Restatement of huffman::decode_one (/root/reference/huffman/src/decode.hpp:83-102)
 * over a table given as per-symbol bitsizes. Reads bits LSB-first within bytes
 * starting at bit `bit_pos`, never past `bit_end`. Returns encoded_size
 * (0 = not found) and stores the symbol. */
int sfo_decode_one(const uint8_t* lens, size_t n, const uint8_t* src,
                   uint64_t bit_pos, uint64_t bit_end, uint16_t* symbol);

/* FNV-1a 64 over a byte range (fingerprints for fixtures). */
uint64_t sfo_fnv1a64(const uint8_t* p, size_t n);

#ifdef __cplusplus
}
#endif
#endif
