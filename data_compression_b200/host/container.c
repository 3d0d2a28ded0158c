/*
 * container.c -- the netstring block container of n_ary_huffman.c: writer :1705-1814, reader
 * :2014-2094, block length parser :1817-1830.
 *
 *   block   = <decimal length> ':' '\n' <type> <data> ',' '\n'          (the length counts "\n" + type + data)
 *   type    '\n'  raw pass-through (:1801-1814)          '#'  metadata, skipped (:2076-2079)
 *           'X'   table: "258:" + one digit per length for symbols 0..258 (:1727-1744)
 *           'Z'   data : "<symbols> <bits>\n" + payload bytes (radix 3: bits = 2 * trits, 5 trits per byte)
 * The reference prints every length with "%d", which only parses while all lengths are <= 9, and its 'Z' writer and
 * both Huffman readers are assert(0) stubs; here a length above 9 is one hex digit (its own limit is 15, :1414), the
 * data block carries the two counts a decoder needs, and block lengths are not capped at 32768 (:1826).  The radix is
 * not stored (the reference has no field for it either: "FUTURE ... a better way of encoding ... compressed_symbols",
 * :1714-1716); the caller passes it.  Payload and tables come from the GPU through refapi.h.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "refapi.h"

#define MAX_SYMBOL_VALUE 258
#define NSLOTS (MAX_SYMBOL_VALUE + 1)

static int g_verbose = 0;
void dc_container_set_verbose(int on) { g_verbose = on; }
#define SAY(...) do { if (g_verbose) printf(__VA_ARGS__); } while (0)

static size_t fail(const char *what) {
    fprintf(stderr, "container: %s\n", what);
    return (size_t)-1;
}

/* radices with a payload form: bit fields (2, 4, 16) and 5 trits per byte (3, the reference's default) */
static int packable(int n) { return n == 2 || n == 3 || n == 4 || n == 16; }

/* ---- block writer: returns bytes written to out (needs out_cap >= text_len + text_len / 4 + 4096) */
size_t dc_container_compress(int n, const int lengths[NSLOTS], char *text, size_t text_len, char *out, size_t out_cap) {
    char *d = out;
    int header_ok = packable(n) && text_len > 0 && text_len <= (size_t)INT32_MAX - 64 && out_cap <= (size_t)INT32_MAX;
    for (int i = 0; i < NSLOTS && header_ok; i++) header_ok = lengths[i] < 16;
    if (header_ok) {
        SAY("# %d : compressed_symbols.\n# header ....\n", n);
        const int table_len = 2 + 3 + 1 + NSLOTS; /* "\nX" "258" ":" digits */
        d += sprintf(d, "%d:\nX%d:", table_len, MAX_SYMBOL_VALUE);
        for (int i = 0; i < NSLOTS; i++) *d++ = "0123456789ABCDEF"[lengths[i]];
        d += sprintf(d, ",\n");
        SAY("# data ....\n");
        /* payload first (at a scratch position behind the largest possible prefix), then the netstring around it */
        char *scratch = d + 64;
        int lens[NSLOTS];
        memcpy(lens, lengths, sizeof lens);
        const int bufsize = (int)(out_cap - 1);
        const int nbytes = represent_items_with_codes(MAX_SYMBOL_VALUE, lens, n, bufsize, (int)text_len, text,
                                                      (int)(scratch - out), out);
        const uint64_t bits = represent_items_last_total_bits();
        char meta[64];
        const int meta_len = sprintf(meta, "\nZ%zu %llu\n", text_len, (unsigned long long)bits);
        const size_t coded = (size_t)(d - out) + 24 + (size_t)meta_len + (size_t)nbytes;
        if (coded < text_len) {
            d += sprintf(d, "%zu:%s", (size_t)meta_len + (size_t)nbytes, meta);
            memmove(d, scratch, (size_t)nbytes);
            d += nbytes;
            d += sprintf(d, ",\n");
            SAY("# compressed: %zu -> %zu bytes (%llu bits).\n", text_len, (size_t)(d - out), (unsigned long long)bits);
            return (size_t)(d - out);
        }
        d = out; /* no saving: fall through to the raw block, as the reference does (:1801-1814) */
    }
    SAY("# pass-through raw data.\n");
    d += sprintf(d, "%zu:\n\n", text_len + 2);
    memcpy(d, text, text_len);
    d += text_len;
    d += sprintf(d, ",\n");
    return (size_t)(d - out);
}

/* ---- block reader: returns the decompressed length, (size_t)-1 on a malformed container */
size_t dc_container_decompress(int n, const char *in, size_t in_len, char *out, size_t out_cap) {
    const char *s = in, *end = in + in_len;
    int lengths[NSLOTS];
    int have_table = 0;
    size_t produced = 0;
    while (s < end) {
        char *colon = NULL;
        const unsigned long long len = strtoull(s, &colon, 10);
        if (!colon || *colon != ':' || colon + 1 + len + 2 > end) return fail("malformed netstring");
        const char *body = colon + 1, *after = body + len;
        if (after[0] != ',' || after[1] != '\n' || len < 2 || body[0] != '\n') return fail("malformed block");
        const char type = body[1];
        const char *data = body + 2;
        const size_t data_len = (size_t)len - 2;
        if (type == '\n') {
            SAY("# raw data:\n");
            if (produced + data_len + 1 > out_cap) return fail("output buffer too small");
            memcpy(out + produced, data, data_len);
            produced += data_len;
        } else if (type == '#') {
            SAY("# skipping metadata.\n");
        } else if (type == 'X') {
            int msv = 0, used = 0;
            if (sscanf(data, "%d:%n", &msv, &used) != 1 || msv != MAX_SYMBOL_VALUE || data_len != (size_t)used + NSLOTS)
                return fail("unsupported table block");
            for (int i = 0; i < NSLOTS; i++) {
                const char c = data[used + i];
                lengths[i] = c >= '0' && c <= '9' ? c - '0' : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1;
                if (lengths[i] < 0) return fail("bad length digit");
            }
            have_table = 1;
        } else if (type == 'Z') {
            if (!have_table) return fail("data block before its table");
            size_t nsym = 0;
            unsigned long long bits = 0;
            int used = 0;
            /* (no "\n" in the format: it would also swallow payload bytes that happen to be white space) */
            if (sscanf(data, "%zu %llu%n", &nsym, &bits, &used) != 2 || data[used] != '\n') return fail("bad data block");
            used += 1;
            if (produced + nsym + 1 > out_cap || nsym > (size_t)INT32_MAX) return fail("output buffer too small");
            /* payload bytes: bits / 8 rounded up; radix 3 counts 2 "bits" per trit and stores 5 trits per byte */
            const unsigned long long want = n == 3 ? (bits / 2 + 4) / 5 : (bits + 7) / 8;
            if (want != data_len - (size_t)used) return fail("data block length mismatch");
            decode_items_with_codes(MAX_SYMBOL_VALUE, lengths, n, bits, data + used, (int)nsym, out + produced);
            produced += nsym;
        } else {
            return fail("unknown block type");
        }
        s = after + 2;
    }
    out[produced] = '\0';
    return produced;
}

