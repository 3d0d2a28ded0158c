/*
 * refapi.h -- the reference's own function names on top of libdc_b200.so.
 *
 * carycode/data_compression has no plugin/FFI layer: a C caller links the non-static functions of
 * n_ary_huffman.c / nybble_compression.c directly.  This header declares those functions with the
 * reference's signatures (file:line cited per function); refapi.c implements each one as a call into the
 * dc_host_* entry points of include/dc_b200.h, i.e. H2D copy -> sm_100a kernels -> D2H copy.  There is no
 * CPU implementation behind them.
 *
 * Error behaviour mirrors the reference: where the reference would assert() and abort, these print one
 * line to stderr and abort() too (dc_refapi_set_abort(0) turns that into a return of the dc_status).
 */
#ifndef DC_REFAPI_H
#define DC_REFAPI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* n_ary_huffman.c:461-493 -- counts bytes of a NUL-terminated text into h[0..max_symbol_value] */
void histogram(const char *text, const int max_symbol_value, int h[]);

/* n_ary_huffman.c:1161-1208 -- n-ary Huffman code lengths (digits), as-written dummy rule */
void huffman(const int max_leaf_value, const int symbol_frequencies[], const int compressed_symbols, int lengths[]);

/* n_ary_huffman.c:1382-1612 -- canonical code values for the lengths */
void convert_lengths_to_encode_table(const int max_symbol_value, const int canonical_lengths[],
                                     const int compressed_symbols, int encode_length_table[],
                                     unsigned int encode_value_table[]);

/* n_ary_huffman.c:1621-1678 -- append the codes of original_text[0..original_length) to
 * compressed_text[start...]; returns the number of bytes written.  (A stub in the reference; the payload
 * layout is the one DESIGN.md defines.)  n in {2,4,16}: bit fields; n == 3: 5 trits per byte, and
 * represent_items_last_total_bits() counts 2 per trit. */
int represent_items_with_codes(const int max_symbol_value, int canonical_lengths[], const int compressed_symbols,
                               const int bufsize, const int original_length, char original_text[], int start,
                               char compressed_text[]);

/* the decoder the reference leaves as assert(0) (n_ary_huffman.c:2081-2089): inverse of
 * represent_items_with_codes; writes n_symbols bytes + a terminating NUL; returns n_symbols */
int decode_items_with_codes(const int max_symbol_value, const int canonical_lengths[], const int compressed_symbols,
                            const uint64_t total_bits, const char compressed_text[], const int n_symbols,
                            char decompressed_text[]);

/* bits emitted by the last represent_items_with_codes() call (the reference's signature has no slot for it) */
uint64_t represent_items_last_total_bits(void);

/* n_ary_huffman.c:1317-1327, :1330-1352, :1354-1379, verbatim signatures: integer power; the largest length and the smallest
 * non-zero length over canonical_lengths[0 .. max_symbol_value) -- the last slot is NOT looked at, as written (300 if every
 * length is zero).  Scalar helpers over 259 integers: host code (the device table keeps the same two numbers in
 * dc_huff_table.max_len / .min_len).  A negative length aborts, like the reference's assert. */
int power(const int base, const int exp);
int array_max(const int max_symbol_value, const int canonical_lengths[]);
int array_min(const int max_symbol_value, const int canonical_lengths[]);

/* n_ary_huffman.c:428-455, verbatim signature: the value 0..63 of a base64url digit ('+' and '/' are accepted for 62 and 63,
 * :441-445); aborts on anything else, like the reference's assert.  A table look-up: host code.  The stream forms -- the
 * binary payload as base64url text, 6 bits per character, as the unfinished packer intends (:1646-1671) -- are
 * dc_base64url_pack / dc_base64url_unpack in dc_b200.h. */
int digit2int(char input_digit);

/* nybble_compression.c:1091-1114, verbatim signature: store one nibble into *dest, offset 0 = the HIGH nibble.
 * One nibble per call is no work for a GPU; it is here so that a caller of the reference links unchanged (it goes
 * through the same pack kernel as the stream form below, which is the one to use). */
void write_nybble(const int nybble, char *dest, bool nybble_offset);

/* stream forms of write_nybble() and of the decoder's split at :767-769 */
void nybble_pack_stream(const unsigned char *symbols, size_t n_symbols, unsigned char *packed);
void nybble_unpack_stream(const unsigned char *packed, size_t n_symbols, unsigned char *symbols);

/* nybble_compression.c:887-1038 / :734-817, both modes on the GPU: the static table (modify == false) and the 16
 * adaptive move-to-front contexts (modify == true).  nybble_compress() / nybble_decompress() are the reference's
 * modify == true wrappers (:1134, :1117). */
void compress_bytestring(const char *source, char *dest, bool modify);
void decompress_bytestring(const char *source, char *dest, bool modify);
void nybble_compress(const char *source, char *dest);
void nybble_decompress(const char *source, char *dest);

/* the netstring block container of n_ary_huffman.c:1705-1814 / :2014-2094 (container.c): table block 'X', data block
 * 'Z', raw block.  compress() / decompress() are static in the reference; these are their linkable counterparts.
 * Return the number of bytes written, (size_t)-1 on a malformed container. */
size_t dc_container_compress(int compressed_symbols, const int canonical_lengths[259], char *text, size_t text_len, char *out,
                             size_t out_cap);
size_t dc_container_decompress(int compressed_symbols, const char *in, size_t in_len, char *out, size_t out_cap);
void dc_container_set_verbose(int on);

/* 1 (default): abort() on a device error, like the reference's assert(); 0: record it and return */
void dc_refapi_set_abort(int on);
/* dc_status of the last call made through this header */
int dc_refapi_last_status(void);

#ifdef __cplusplus
}
#endif
#endif /* DC_REFAPI_H */
