/*
 * oracle_nybble.c -- CPU restatement of the nibble primitives (and the static-table
 * compressor) of /root/reference/nybble_compression.c.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"

#include <string.h>

/*
 * write_nybble() nybble_compression.c:1091-1114 with `little_endian` undefined (#else branches
 * :1100-1101, :1109-1110): offset 0 -> high nibble, offset 1 -> low nibble.  Stream form: symbol 2i goes
 * to the high nibble of byte i, symbol 2i+1 to the low nibble; an odd tail leaves the low nibble 0
 * (SURVEY H12).
 */
void orc_nybble_pack(const uint8_t *sym, size_t n_sym, uint8_t *packed) {
    size_t i = 0;
    for (; i + 2 <= n_sym; i += 2) packed[i >> 1] = (uint8_t)(((sym[i] & 0x0F) << 4) | (sym[i + 1] & 0x0F));
    if (i < n_sym) packed[i >> 1] = (uint8_t)((sym[i] & 0x0F) << 4);
}

/* decoder split nybble_compression.c:767-769: hi = (b >> 4) & 0xF first, lo = b & 0xF second. */
void orc_nybble_unpack(const uint8_t *packed, size_t n_sym, uint8_t *sym) {
    for (size_t i = 0; i < n_sym; i++) {
        const uint8_t b = packed[i >> 1];
        sym[i] = (i & 1) ? (uint8_t)(b & 0x0F) : (uint8_t)((b >> 4) & 0x0F);
    }
}

void orc_nybble_pack_mt(const uint8_t *sym, size_t n_sym, uint8_t *packed, int threads) {
    const size_t nbytes = (n_sym + 1) / 2, chunk = 1u << 16;
    const long long nchunks = (long long)((nbytes + chunk - 1) / chunk);
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long c = 0; c < nchunks; c++) {
        const size_t lo = (size_t)c * chunk, hi = lo + chunk < nbytes ? lo + chunk : nbytes;
        const size_t s_lo = lo * 2, s_hi = hi * 2 < n_sym ? hi * 2 : n_sym;
        orc_nybble_pack(sym + s_lo, s_hi - s_lo, packed + lo);
    }
}

void orc_nybble_unpack_mt(const uint8_t *packed, size_t n_sym, uint8_t *sym, int threads) {
    const size_t nbytes = (n_sym + 1) / 2, chunk = 1u << 16;
    const long long nchunks = (long long)((nbytes + chunk - 1) / chunk);
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long c = 0; c < nchunks; c++) {
        const size_t lo = (size_t)c * chunk, hi = lo + chunk < nbytes ? lo + chunk : nbytes;
        const size_t s_lo = lo * 2, s_hi = hi * 2 < n_sym ? hi * 2 : n_sym;
        orc_nybble_unpack(packed + lo, s_hi - s_lo, sym + s_lo);
    }
}

/* ---------------------------------------------------------------- static-table compressor (row N1) */

/* initialize_dictionary() :546-562: every context holds " etaoins"; with modify=false it never changes,
 * so the context (:825) is irrelevant. */
static const char k_letters[8] = {' ', 'e', 't', 'a', 'o', 'i', 'n', 's'};

static int letter_index(uint8_t s) {
    for (int i = 0; i < 8; i++)
        if ((uint8_t)k_letters[i] == s) return i;
    return -1;
}

/*
 * compress_bytestring(src, dst, false) :887-1038 with compress_byte_index :819-884.
 * dst[0]=0xAF (:903), dst[1]=src[0] (:905); then per byte: table hit -> nibble (8|i) (:877-883);
 * miss at nibble offset 0 -> literal byte (:844-846); miss at offset 1 -> the half-written byte is
 * replaced by the previous source byte as a literal, followed by this byte (:855-857).  A trailing
 * half byte is expanded to a literal (:1000-1009).  If the result is not shorter than the source it
 * becomes ' ' + raw copy (:1018-1037).  Length-explicit: src must not contain 0x00 or bytes >= 0x80 (:910).
 */
size_t orc_nybble_static_compress(const uint8_t *src, size_t n, uint8_t *dst) {
    if (n == 0) { dst[0] = 0; return 0; }
    size_t o = 0;
    dst[o++] = 0xAF;
    dst[o++] = src[0];
    int off = 0;
    for (size_t i = 1; i < n; i++) {
        const int idx = letter_index(src[i]);
        int used;
        if (idx < 0) {
            if (off == 0) { dst[o] = src[i]; used = 2; }
            else { dst[o] = src[i - 1]; dst[o + 1] = src[i]; used = 3; }
        } else {
            const uint8_t nyb = (uint8_t)(idx | 0x8);
            if (off == 0) dst[o] = (uint8_t)(nyb << 4);
            else dst[o] |= nyb;
            used = 1;
        }
        off += used;
        if (off > 1) { o++; off -= 2; }
        if (off > 1) { o++; off -= 2; }
    }
    if (off != 0) dst[o++] = src[n - 1];
    dst[o] = 0;
    if (o >= n) {
        o = 0;
        dst[o++] = ' ';
        memcpy(dst + o, src, n);
        o += n;
        dst[o] = 0;
    }
    return o;
}

/* decompress_bytestring(src, dst, false) :734-817 with decompress_nybble :643-663. */
size_t orc_nybble_static_decompress(const uint8_t *src, size_t n, uint8_t *dst) {
    size_t o = 0;
    if (n == 0) { dst[0] = 0; return 0; }
    if (src[0] == 0xAF) {
        if (n < 2) { dst[0] = 0; return 0; }
        dst[o++] = src[1];
        size_t p = 2;
        int off = 0;
        while (p < n) {
            const unsigned b = src[p], nb = p + 1 < n ? src[p + 1] : 0;
            unsigned nyb, next;
            if (off == 0) { nyb = (b >> 4) & 0xF; next = b & 0xF; }
            else { nyb = b & 0xF; next = (nb >> 4) & 0xF; }
            if (nyb & 0x8) { dst[o++] = (uint8_t)k_letters[nyb & 7]; off += 1; }
            else { dst[o++] = (uint8_t)(((nyb & 7) << 4) + next); off += 2; }
            if (off >= 2) { p++; off -= 2; }
        }
    } else {
        /* ' ' (LITERAL) skips the type byte (:799-805); any other type copies everything (:806-812) */
        size_t p = src[0] == ' ' ? 1 : 0;
        while (p < n) dst[o++] = src[p++];
    }
    dst[o] = 0;
    return o;
}

/* ---------------------------------------------------------------- adaptive compressor (row N3) */

/*
 * modify == true (nybble_compress :1134 / nybble_decompress :1117): 16 contexts, chosen by bits 3..6 of the
 * previous byte (byte_to_context :517-523), each a move-to-front list of 8 letters that starts as
 * " etaoins" (:546-562).  After every byte -- hit or miss, and on both sides -- the byte is moved to the
 * front of its context's list (update_context :665-687): entries in front of its old position (or all
 * but the last, if it was absent) shift back by one.  Emission is the static coder's (:819-884).
 */
typedef struct { uint8_t letter[16][8]; } orc_ctx_table;

static void ctx_init(orc_ctx_table *t) {
    for (int c = 0; c < 16; c++) memcpy(t->letter[c], k_letters, 8);
}
static int ctx_of(uint8_t prev) { return (prev >> 3) & 15; }
static int ctx_find(const orc_ctx_table *t, int c, uint8_t s) {
    for (int i = 0; i < 8; i++)
        if (t->letter[c][i] == s) return i;
    return -1;
}
static void ctx_touch(orc_ctx_table *t, int c, uint8_t s) {
    int at = ctx_find(t, c, s);
    if (at < 0) at = 7;
    memmove(&t->letter[c][1], &t->letter[c][0], (size_t)at);
    t->letter[c][0] = s;
}

size_t orc_nybble_adaptive_compress(const uint8_t *src, size_t n, uint8_t *dst) {
    if (n == 0) { dst[0] = 0; return 0; }
    orc_ctx_table t;
    ctx_init(&t);
    size_t o = 0;
    dst[o++] = 0xAF;
    dst[o++] = src[0];
    int off = 0;
    for (size_t i = 1; i < n; i++) {
        const int c = ctx_of(src[i - 1]);
        const int idx = ctx_find(&t, c, src[i]);
        int used;
        if (idx < 0) {
            if (off == 0) { dst[o] = src[i]; used = 2; }
            else { dst[o] = src[i - 1]; dst[o + 1] = src[i]; used = 3; }
        } else {
            const uint8_t nyb = (uint8_t)(idx | 0x8);
            if (off == 0) dst[o] = (uint8_t)(nyb << 4);
            else dst[o] |= nyb;
            used = 1;
        }
        ctx_touch(&t, c, src[i]);
        off += used;
        if (off > 1) { o++; off -= 2; }
        if (off > 1) { o++; off -= 2; }
    }
    if (off != 0) dst[o++] = src[n - 1];
    dst[o] = 0;
    if (o >= n) {
        o = 0;
        dst[o++] = ' ';
        memcpy(dst + o, src, n);
        o += n;
        dst[o] = 0;
    }
    return o;
}

size_t orc_nybble_adaptive_decompress(const uint8_t *src, size_t n, uint8_t *dst) {
    size_t o = 0;
    if (n == 0) { dst[0] = 0; return 0; }
    if (src[0] == 0xAF) {
        if (n < 2) { dst[0] = 0; return 0; }
        orc_ctx_table t;
        ctx_init(&t);
        dst[o++] = src[1];
        size_t p = 2;
        int off = 0;
        while (p < n) {
            const unsigned b = src[p], nb = p + 1 < n ? src[p + 1] : 0;
            unsigned nyb, next;
            if (off == 0) { nyb = (b >> 4) & 0xF; next = b & 0xF; }
            else { nyb = b & 0xF; next = (nb >> 4) & 0xF; }
            const int c = ctx_of(dst[o - 1]);
            uint8_t out;
            if (nyb & 0x8) { out = t.letter[c][nyb & 7]; off += 1; }
            else { out = (uint8_t)(((nyb & 7) << 4) + next); off += 2; }
            dst[o++] = out;
            ctx_touch(&t, c, out);
            if (off >= 2) { p++; off -= 2; }
        }
    } else {
        size_t p = src[0] == ' ' ? 1 : 0;
        while (p < n) dst[o++] = src[p++];
    }
    dst[o] = 0;
    return o;
}
