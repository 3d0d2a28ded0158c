/*
 * nybble_compression.c -- host shim with the reference binary's command line: `./nybble_compression`.
 *
 * The reference's main() (nybble_compression.c:1139-1219) takes no input and round-trips a fixed text,
 * printing `Successful test.`.  This shim does the same for the nibble-granular hot path that runs on the
 * GPU: the split of every byte into (high, low) nibbles (:767-769) and write_nybble()'s packing order
 * (:1091-1114), through refapi.h -> libdc_b200.so.
 *
 * and for the compressor: compress_bytestring(text, c, false) / decompress_bytestring (:1160-1166) with the static
 * table, then nybble_compress() / nybble_decompress() (:1176-1215) with the adaptive contexts, each with the
 * reference's `assert(strlen(compressed) <= 70)` (:1162, :1178).
 *
 *   ./nybble_compression            fixed text, as the reference
 *   ./nybble_compression - < file   the nibble round trip over stdin (and the compressor, if the file is 7-bit text)
 */
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "refapi.h"

static void round_trip(const unsigned char *bytes, size_t n) {
    const size_t nsym = 2 * n;
    unsigned char *sym = malloc(nsym + 1), *packed = malloc(n + 1);
    if (!sym || !packed) abort();
    nybble_unpack_stream(bytes, nsym, sym);
    for (size_t i = 0; i < n; i++) { /* spot check of the order on the host: high nibble first */
        if (sym[2 * i] != (bytes[i] >> 4) || sym[2 * i + 1] != (bytes[i] & 0x0F)) {
            printf("Error: nibble order.\n");
            abort();
        }
        if (i == 63) i = n > 128 ? n - 65 : i;
    }
    nybble_pack_stream(sym, nsym, packed);
    if (memcmp(packed, bytes, n) != 0) {
        printf("Error: repacked text doesn't match original text.\n");
        abort();
    }
    /* odd symbol count: the low nibble of the last byte stays 0 */
    if (nsym > 1) {
        memset(packed, 0xFF, n);
        nybble_pack_stream(sym, nsym - 1, packed);
        if (packed[n - 1] != (bytes[n - 1] & 0xF0) || (n > 1 && packed[n - 2] != bytes[n - 2])) {
            printf("Error: odd tail.\n");
            abort();
        }
    }
    printf("Successful test.\n");
    free(sym);
    free(packed);
}

static void text_round_trip(const char *text, size_t limit, bool modify) {
    const size_t n = strlen(text);
    char *c = malloc(n + 2), *d = malloc(2 * n + 2);
    if (!c || !d) abort();
    compress_bytestring(text, c, modify);
    printf("# compressed %zu -> %zu bytes (%s)\n", n, strlen(c), modify ? "adaptive contexts" : "static table");
    if (limit && strlen(c) > limit) {
        printf("Error: compressed text longer than %zu bytes.\n", limit);
        abort();
    }
    decompress_bytestring(c, d, modify);
    if (strlen(d) != n || memcmp(text, d, n) != 0) {
        printf("Error: decompressed text doesn't match original text.\n");
        abort();
    }
    printf("Successful test.\n");
    free(c);
    free(d);
}

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "-")) {
        size_t cap = 1 << 16, used = 0;
        unsigned char *buf = malloc(cap);
        if (!buf) abort();
        for (;;) {
            used += fread(buf + used, 1, cap - used, stdin);
            if (used < cap || cap >= ((size_t)1 << 30)) break;
            cap *= 2;
            buf = realloc(buf, cap);
            if (!buf) abort();
        }
        printf("# %zu bytes -> %zu nybbles -> %zu bytes\n", used, 2 * used, used);
        if (used) round_trip(buf, used);
        int is_text = used > 0;
        for (size_t i = 0; i < used && is_text; i++) is_text = buf[i] != 0 && buf[i] < 0x80;
        if (is_text) {
            buf = realloc(buf, used + 1);
            if (!buf) abort();
            buf[used] = 0;
            text_round_trip((const char *)buf, 0, false);
            text_round_trip((const char *)buf, 0, true);
        }
        free(buf);
        return 0;
    }
    const char *text = "Hello, world. This is a test. This is only a test. Banana banana banana banana. ";
    printf("# %zu bytes -> %zu nybbles -> %zu bytes\n", strlen(text), 2 * strlen(text), strlen(text));
    round_trip((const unsigned char *)text, strlen(text));
    text_round_trip(text, 70, false);
    text_round_trip(text, 70, true);
    return 0;
}
