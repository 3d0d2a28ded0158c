/* refapi.c -- the reference's function names implemented on the dc_host_* C-ABI (see refapi.h). */
#include "refapi.h"

#include <stdio.h>
#include <stdlib.h>

#include "dc_b200.h"

static int g_abort = 1;
static int g_last = DC_OK;
static uint64_t g_last_bits = 0;

void dc_refapi_set_abort(int on) { g_abort = on; }
int dc_refapi_last_status(void) { return g_last; }
uint64_t represent_items_last_total_bits(void) { return g_last_bits; }

static int done(const char *what, int status) {
    g_last = status < 0 ? status : DC_OK;
    if (status < 0 && g_abort) {
        fprintf(stderr, "%s: %s\n", what, dc_status_string(status));
        abort();
    }
    return status;
}

void histogram(const char *text, const int max_symbol_value, int h[]) {
    done("histogram", dc_host_histogram(text, max_symbol_value, h));
}

void huffman(const int max_leaf_value, const int symbol_frequencies[], const int compressed_symbols, int lengths[]) {
    done("huffman", dc_host_huffman(max_leaf_value, symbol_frequencies, compressed_symbols, lengths));
}

void convert_lengths_to_encode_table(const int max_symbol_value, const int canonical_lengths[],
                                     const int compressed_symbols, int encode_length_table[],
                                     unsigned int encode_value_table[]) {
    done("convert_lengths_to_encode_table",
         dc_host_convert_lengths_to_encode_table(max_symbol_value, canonical_lengths, compressed_symbols,
                                                 encode_length_table, encode_value_table));
}

int represent_items_with_codes(const int max_symbol_value, int canonical_lengths[], const int compressed_symbols,
                               const int bufsize, const int original_length, char original_text[], int start,
                               char compressed_text[]) {
    g_last_bits = 0;
    const int rc = dc_host_represent_items_with_codes(max_symbol_value, canonical_lengths, compressed_symbols, bufsize,
                                                      original_length, original_text, start, compressed_text,
                                                      &g_last_bits);
    return done("represent_items_with_codes", rc);
}

int decode_items_with_codes(const int max_symbol_value, const int canonical_lengths[], const int compressed_symbols,
                            const uint64_t total_bits, const char compressed_text[], const int n_symbols,
                            char decompressed_text[]) {
    if (max_symbol_value != DC_MAX_SYMBOL_VALUE || n_symbols < 0) return done("decode_items_with_codes", DC_ERR_ARG);
    const int rc = dc_host_huff_decompress((const uint8_t *)compressed_text, total_bits, canonical_lengths,
                                           compressed_symbols, (uint8_t *)decompressed_text, (size_t)n_symbols);
    if (rc == DC_OK) decompressed_text[n_symbols] = '\0';
    return done("decode_items_with_codes", rc) < 0 ? rc : n_symbols;
}

void write_nybble(const int nybble, char *dest, bool nybble_offset) {
    if (nybble < 0 || nybble >= 0x10 || !dest) {  /* assert( nybble < 0x10 ) :1093 */
        done("write_nybble", DC_ERR_SYMBOL);
        return;
    }
    const unsigned char old = (unsigned char)*dest;
    unsigned char sym[2], packed = 0;
    sym[0] = nybble_offset ? (unsigned char)(old >> 4) : (unsigned char)nybble;
    sym[1] = nybble_offset ? (unsigned char)nybble : (unsigned char)(old & 0x0F);
    if (done("write_nybble", dc_host_nybble_pack(sym, 2, &packed)) == DC_OK) *dest = (char)packed;
}

void nybble_pack_stream(const unsigned char *symbols, size_t n_symbols, unsigned char *packed) {
    done("nybble_pack_stream", dc_host_nybble_pack(symbols, n_symbols, packed));
}

void nybble_unpack_stream(const unsigned char *packed, size_t n_symbols, unsigned char *symbols) {
    done("nybble_unpack_stream", dc_host_nybble_unpack(packed, n_symbols, symbols));
}

void compress_bytestring(const char *source, char *dest, bool modify) {
    const long long rc = dc_host_compress_bytestring(source, dest, modify ? 1 : 0);
    done("compress_bytestring", rc < 0 ? (int)rc : DC_OK);
}

void decompress_bytestring(const char *source, char *dest, bool modify) {
    const long long rc = dc_host_decompress_bytestring(source, dest, modify ? 1 : 0);
    done("decompress_bytestring", rc < 0 ? (int)rc : DC_OK);
}

void nybble_compress(const char *source, char *dest) { compress_bytestring(source, dest, true); }

void nybble_decompress(const char *source, char *dest) { decompress_bytestring(source, dest, true); }

int digit2int(char input_digit) {
    static const char digits[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";
    if (input_digit == '+') return 62;
    if (input_digit == '/') return 63;
    for (int i = 0; i < 64; i++)
        if (digits[i] == input_digit) return i;
    done("digit2int", DC_ERR_ARG);
    return -1;
}

int power(const int base, const int exp) {
    if (exp < 0) done("power", DC_ERR_ARG);
    int r = 1;
    for (int b = base, e = exp; e > 0; e >>= 1, b *= b)
        if (e & 1) r *= b;
    return r;
}

int array_max(const int max_symbol_value, const int canonical_lengths[]) {
    int m = 0;
    for (int i = 0; i < max_symbol_value; i++) {
        if (canonical_lengths[i] < 0) done("array_max", DC_ERR_ARG);
        if (canonical_lengths[i] > m) m = canonical_lengths[i];
    }
    return m;
}

int array_min(const int max_symbol_value, const int canonical_lengths[]) {
    int m = 300;
    for (int i = 0; i < max_symbol_value; i++) {
        if (canonical_lengths[i] < 0) done("array_min", DC_ERR_ARG);
        if (canonical_lengths[i] != 0 && canonical_lengths[i] < m) m = canonical_lengths[i];
    }
    return m;
}
