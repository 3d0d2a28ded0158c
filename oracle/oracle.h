/*
 * oracle.h -- CPU oracle for the n-ary Huffman / nybble hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product path
 * (data_compression_b200/, include/dc_b200.h) never links or imports it.
 *
 * Parity status:
 *   - histogram, code lengths, canonical code values: PINNED against the
 *     unmodified reference functions (oracle/_ref, built from
 *     /root/reference by oracle/Makefile) and against committed golden
 *     vectors in tests/golden/ generated from those functions.
 *   - nibble order: PINNED against write_nybble / the decoder split of
 *     nybble_compression.c through the same harness.
 *   - Huffman payload bit layout and decoder: PARITY UNPINNED.  The reference
 *     has neither (n_ary_huffman.c:1661 assert(0), :2081-2089 assert(0));
 *     the layout is defined by this repository (DESIGN.md "payload layout")
 *     and this oracle is its CPU statement.
 *
 * All citations file:line are into /root/reference/.
 */
#ifndef DC_ORACLE_H
#define DC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORC_OK = 0,
    ORC_ERR_ARG = -1,
    ORC_ERR_CODE_TOO_LONG = -3,
    ORC_ERR_CAPACITY = -4,
    ORC_ERR_CORRUPT = -5,
    ORC_ERR_SYMBOL = -6
};

/* n_ary_huffman.c:461-493 -- counts bytes until NUL; zeroes h[0..max_symbol_value] first. */
void orc_histogram_cstr(const char *text, int max_symbol_value, uint64_t h[]);
/* length-explicit form (accepts 0x00); zeroes h[0..nslots-1] first. */
void orc_histogram_u8(const uint8_t *in, size_t n, uint64_t h[], int nslots);
void orc_histogram_u8_mt(const uint8_t *in, size_t n, uint64_t h[], int nslots, int threads);

/* n_ary_huffman.c:1161-1208 (setup_nodes :773, generate_huffman_tree :868,
 * summarize_tree_with_lengths :1033), as compiled with -DNDEBUG.  64-bit counts. */
int orc_huffman(int max_leaf_value, const uint64_t freqs[], int compressed_symbols, int lengths[]);

/* n_ary_huffman.c:1382-1612 incl. the "i < max_symbol_value" quirks (:1336,:1360,:1421). */
int orc_convert_lengths_to_encode_table(int max_symbol_value, const int lengths[], int compressed_symbols,
                                        int encode_length_table[], unsigned int encode_value_table[]);

/* log2(n) for n in {2,4,16}, else 0 (no bit packing defined). */
int orc_bits_per_digit(int compressed_symbols);

/* Repo-defined payload (n_ary_huffman.c:1621-1678 intent): codes MSB-first, concatenated in input
 * order starting at bit `bit_phase` (0..7) of out[0]; bytes touched are fully written, unused bits 0. */
int orc_pack(const uint8_t *in, size_t n, const int elen[], const unsigned int eval[], int bits_per_digit,
             unsigned bit_phase, uint8_t *out, size_t out_capacity, uint64_t *total_bits);
int orc_pack_mt(const uint8_t *in, size_t n, const int elen[], const unsigned int eval[], int bits_per_digit,
                unsigned bit_phase, uint8_t *out, size_t out_capacity, uint64_t *total_bits, int threads,
                uint64_t *block_bit_offsets, size_t block_symbols);

/* Sequential canonical decoder of that payload: decodes codes from bit `bit_start` until `nbits` bits are
 * consumed or n_out_capacity symbols are produced. */
int orc_unpack(const uint8_t *bits, uint64_t bit_start, uint64_t nbits, int max_symbol_value,
               const int lengths[], int compressed_symbols, uint8_t *out, size_t n_out_capacity,
               size_t *n_decoded);
/* block-parallel decode given the encoder-side block offsets (a generous CPU baseline). */
int orc_unpack_mt(const uint8_t *bits, uint64_t bit_start, int max_symbol_value, const int lengths[],
                  int compressed_symbols, uint8_t *out, size_t n, const uint64_t *block_bit_offsets,
                  size_t block_symbols, int threads);

/* n = 3 (the reference's default radix, :2529): the trit payload sketched at n_ary_huffman.c:745-748 -- 5 trits per byte,
 * byte = 1 + base-3 value of the group, most significant trit first, last group zero-padded.  PARITY UNPINNED like the
 * bit payload (the reference has no packer).  The decoder is a plain O(alphabet) search per trit: small inputs only. */
int orc_pack_trits(const uint8_t *in, size_t n, const int elen[], const unsigned int eval[], uint8_t *out,
                   size_t out_capacity, uint64_t *total_trits);
int orc_unpack_trits(const uint8_t *packed, uint64_t total_trits, int max_symbol_value, const int lengths[],
                     uint8_t *out, size_t n_out_capacity, size_t *n_decoded);

/* base64url text form of a binary payload: int2digit() n_ary_huffman.c:371-426 / digit2int() :428-455, 6 bits per character
 * as the unfinished packer intends (:1646-1671).  pack returns the number of characters; unpack ORC_OK / ORC_ERR_CORRUPT. */
size_t orc_base64url_pack(const uint8_t *bits, uint64_t nbits, uint8_t *chars);
int orc_base64url_unpack(const uint8_t *chars, uint64_t nbits, uint8_t *bits);

/* nybble_compression.c:1091-1114 (write_nybble, #else branches) and :767-773 (split): high nibble first. */
void orc_nybble_pack(const uint8_t *sym, size_t n_sym, uint8_t *packed);
void orc_nybble_unpack(const uint8_t *packed, size_t n_sym, uint8_t *sym);
void orc_nybble_pack_mt(const uint8_t *sym, size_t n_sym, uint8_t *packed, int threads);
void orc_nybble_unpack_mt(const uint8_t *packed, size_t n_sym, uint8_t *sym, int threads);

/* nybble_compression.c:887-1038 / :734-817 with modify=false (static " etaoins" table), print-free,
 * length-explicit.  Returns bytes written (excluding the terminating NUL it also writes). */
size_t orc_nybble_static_compress(const uint8_t *src, size_t n, uint8_t *dst);
size_t orc_nybble_static_decompress(const uint8_t *src, size_t n, uint8_t *dst);
/* the same with modify=true: 16 move-to-front contexts (update_context :665-687). */
size_t orc_nybble_adaptive_compress(const uint8_t *src, size_t n, uint8_t *dst);
size_t orc_nybble_adaptive_decompress(const uint8_t *src, size_t n, uint8_t *dst);

#ifdef __cplusplus
}
#endif
#endif
