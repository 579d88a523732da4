"""Known-answer vectors for the raw-DEFLATE decompress path.

Sources:
  * the reference's own tests (/root/reference/src/test/decompress_test.cpp:62-181),
  * answers probed against the unmodified reference.
Each entry: (name, hex stream, dst capacity, expected status, expected dst prefix hex or None,
             class) where class is
  "D" defined in every reference build, "A" defined under NDEBUG only (assert build aborts),
  "U" undefined behaviour in the reference (expected status is this repo's documented choice).
oracle/pin_oracle.py re-derives status/class from the reference binaries and fails on any
disagreement with this table.
"""

SUCCESS, ERROR, INVALID_BLOCK_HEADER, LEN_MISMATCH, DST_TOO_SMALL, SRC_TOO_SMALL, \
    INVALID_LIT_OR_LEN, INVALID_DISTANCE = range(8)

ROSEBUD = bytes([0b000, 4, 0, 0xFB, 0xFF]) + b"rose" + bytes([0b001, 3, 0, 0xFC, 0xFF]) + b"bud"

KNOWN = [
    # --- reference tests -------------------------------------------------------------------
    ("ref_empty", "", 0, INVALID_BLOCK_HEADER, None, "D"),                 # decompress_test.cpp:91-95
    ("ref_rosebud_dst_small", ROSEBUD.hex(), 6, DST_TOO_SMALL, b"rose".hex(), "D"),   # :116-119
    ("ref_rosebud_src_small", ROSEBUD[:5].hex(), 7, SRC_TOO_SMALL, None, "D"),        # :121-123
    ("ref_rosebud_ok", ROSEBUD.hex(), 7, SUCCESS, b"rosebud".hex(), "D"),             # :125-127
    # --- SURVEY.md §8(c) ------------------------------------------------------------------
    ("btype3", "07", 4, INVALID_BLOCK_HEADER, None, "D"),
    ("stored_len_mismatch", "010200000041 42".replace(" ", ""), 4, LEN_MISMATCH, None, "D"),
    ("stored_empty", "010000ffff", 4, SUCCESS, "", "D"),
    ("stored_trunc_lennlen", "010200", 4, SRC_TOO_SMALL, None, "U"),
    ("fixed_A_cap1", "730400", 1, SUCCESS, "41", "D"),
    ("fixed_A_cap3", "730400", 3, SUCCESS, "41", "D"),
    ("fixed_A_cap0", "730400", 0, DST_TOO_SMALL, None, "D"),
    ("fixed_A_trailing", "730400ffff", 1, SUCCESS, "41", "D"),
    ("fixed_sym286", "1b03", 4, INVALID_LIT_OR_LEN, None, "D"),
    ("fixed_sym287", "1b07", 4, INVALID_LIT_OR_LEN, None, "D"),
    ("fixed_dist30", "73043e00", 8, INVALID_LIT_OR_LEN, "41", "D"),
    ("fixed_dist_too_far", "73044200", 8, INVALID_DISTANCE, "41", "D"),
    ("fixed_match_at_0", "030200", 8, INVALID_DISTANCE, None, "D"),
    ("fixed_len258_ok", "731c0500", 259, SUCCESS, "41" * 259, "D"),
    ("fixed_len258_small", "731c0500", 258, DST_TOO_SMALL, "41", "D"),
    ("fixed_nonfinal_eof_midsym", "7204", 1, INVALID_LIT_OR_LEN, "41", "D"),
    ("fixed_final_no_eob", "7304", 4, INVALID_LIT_OR_LEN, "41", "D"),
    ("fixed_nonfinal_then_pad", "720400", 1, SRC_TOO_SMALL, "41", "U"),
    ("fixed_eof_before_dist", "731c01", 300, INVALID_DISTANCE, "41", "A"),
    ("dyn_incomplete_cl_empty_dist", "05e0dbb66ddbb66ddb261b4cff948280", 2, SUCCESS, "4141", "D"),
    ("dyn_oversubscribed", "05e0dbb66ddbb66ddb261b84e99f501008", 4, DST_TOO_SMALL, "41424141", "A"),
    ("dyn_hlit31", "fde0dbb66ddbb66ddb261b4cff9482890204", 1, SUCCESS, "41", "D"),
    ("dyn_match_empty_dist", "0de0dbb66ddbb66ddb261b4aff94421004", 8, INVALID_DISTANCE, "41", "A"),
    ("dyn_single_dist_code_bit1", "0de0dbb66ddbb66ddb261b4aff94421026", 8, INVALID_DISTANCE, "41", "A"),
    ("dyn_single_dist_code_bit0", "0de0dbb66ddbb66ddb261b4aff944210c6", 4, SUCCESS, "41414141", "D"),
    ("dyn_no_eob_cap4", "05e0dbb66ddbb66ddb261b4cffa4020000", 4, DST_TOO_SMALL, "41414141", "D"),
    ("dyn_no_eob_cap64", "05e0dbb66ddbb66ddb261b4cffa4020000", 64, INVALID_LIT_OR_LEN, "41" * 16, "D"),
    ("dyn_all_zero_litlen", "05e0dbb66ddbb66ddba67f620300", 4, INVALID_LIT_OR_LEN, None, "D"),
    ("dyn_cl_unassigned", "05e0dbb66ddbb66ddb4e", 4, INVALID_LIT_OR_LEN, None, "D"),
    ("dyn_rep16_first", "05e0dbb66ddbb66ddb06920da67f4a4100", 4, INVALID_LIT_OR_LEN, None, "U"),
    ("dyn_rep16_first_dist", "05e0dbb66ddbb66ddb261b4cff948201", 4, INVALID_LIT_OR_LEN, None, "U"),
    ("dyn_rep18_overflow", "05e0dbb66ddbb66ddb261b4cff344a7f00", 4, INVALID_LIT_OR_LEN, None, "U"),
    ("dyn_eof_in_hclen", "05e0db", 4, SRC_TOO_SMALL, None, "U"),
    # RFC 1951 §3.2.7: the HLIT + HDIST code lengths are one sequence.  A repeat that runs from the
    # lit/len lengths into the distance lengths, or a 16 as the first distance length, is valid
    # (zlib decodes both to "aaaaa"; libdeflate, zopfli and 7-zip write such headers); the reference
    # decodes two independent runs and is undefined here (class U): decoded as the RFC says.
    ("dyn_rep16_crosses_into_dist", "05e307090000008030649dfd43f8c0", 8, SUCCESS, "6161616161", "U"),
    ("dyn_rep16_first_dist_valid", "05e307090000008030649dfd43780403", 8, SUCCESS, "6161616161", "U"),
]

# read_header vectors (decompress_test.cpp:62-89): (bytes, bit_size, has_value, final, type, error)
READ_HEADER = [
    (b"", 0, 0, 0, 0, INVALID_BLOCK_HEADER),
    (bytes([0b111]), 8, 0, 0, 0, INVALID_BLOCK_HEADER),
    (bytes([0b010]), 8, 1, 0, 1, 0),
    (bytes([0b001]), 8, 1, 1, 0, 0),
]

# copy_from_before (decompress_test.cpp:176-181)
COPY_FROM_BEFORE = (bytes([1, 2, 0, 0, 0, 0]), 2, 2, 3, bytes([1, 2, 1, 2, 1, 0]))

STARFLEET_MD5 = "75c49efb38e3688dfe00eb4316ccd69d"
STARFLEET_LEN = 149618
