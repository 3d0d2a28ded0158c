/*
 * oracle_huff.c -- CPU restatement of the n-ary Huffman path of the reference.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Restates, print-free and with
 * 64-bit counts, what /root/reference/n_ary_huffman.c computes when compiled
 * with -DNDEBUG (the as-written dummy rule is only executable that
 * way).  The tree builder is the two-queue form of the reference's stable
 * bubble-sort merge loop; tests/test_oracle_vs_reference.py checks it against
 * the unmodified reference functions on random, tied and edge histograms.
 *
 * The bit packer / decoder at the bottom are NOT restatements: the reference
 * has none (n_ary_huffman.c:1661, :2081-2089).  They define the payload layout
 * of this repository ("parity unpinned").
 */
#include "oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ histogram */

/* n_ary_huffman.c:461-493: zero h[0..max_symbol_value], then h[*c]++ until NUL. */
void orc_histogram_cstr(const char *text, int max_symbol_value, uint64_t h[]) {
    for (int i = 0; i < max_symbol_value + 1; i++) h[i] = 0;
    const unsigned char *c = (const unsigned char *)text;
    while (*c) {
        h[*c]++;
        c++;
    }
}

void orc_histogram_u8(const uint8_t *in, size_t n, uint64_t h[], int nslots) {
    for (int i = 0; i < nslots; i++) h[i] = 0;
    for (size_t i = 0; i < n; i++) h[in[i]]++;
}

void orc_histogram_u8_mt(const uint8_t *in, size_t n, uint64_t h[], int nslots, int threads) {
    for (int i = 0; i < nslots; i++) h[i] = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        uint64_t local[4][256];
        memset(local, 0, sizeof local);
#ifdef _OPENMP
        int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int t = 0, nt = 1;
#endif
        size_t lo = n / (size_t)nt * (size_t)t, hi = (t == nt - 1) ? n : n / (size_t)nt * (size_t)(t + 1);
        size_t i = lo;
        for (; i + 4 <= hi; i += 4) {
            local[0][in[i]]++;
            local[1][in[i + 1]]++;
            local[2][in[i + 2]]++;
            local[3][in[i + 3]]++;
        }
        for (; i < hi; i++) local[0][in[i]]++;
#pragma omp critical
        for (int b = 0; b < 256; b++) h[b] += local[0][b] + local[1][b] + local[2][b] + local[3][b];
    }
}

/* ------------------------------------------------------------------ tree -> lengths */

typedef struct {
    uint64_t count;
    int index; /* symbol value; >= nsym for dummy leaves (:921-929) */
} orc_leaf;

static int leaf_cmp(const void *a, const void *b) {
    const orc_leaf *x = (const orc_leaf *)a, *y = (const orc_leaf *)b;
    if (x->count != y->count) return x->count < y->count ? -1 : 1;
    return x->index < y->index ? -1 : (x->index > y->index);
}

/*
 * huffman() :1161-1208.
 *  - nz = leaves with non-zero count (:880-886)
 *  - d  = (n-1) - ((nz-1) % (n-1)) dummy leaves of count 1 at indices max_leaf_value+1.. (:900-903,:921-929)
 *  - partial_sort (:672-731) is a stable sort by count only => order (count, node index); a freshly created
 *    internal node sits at the right end, so it lands after every node of equal count (:962-1002)
 *    == two queues, leaf wins ties, internals FIFO.
 *  - lengths = parent hops to the root (:1069-1076).
 */
int orc_huffman(int max_leaf_value, const uint64_t freqs[], int compressed_symbols, int lengths[]) {
    const int n = compressed_symbols;
    const int nsym = max_leaf_value + 1;
    if (n < 2 || nsym < 1) return ORC_ERR_ARG;
    for (int i = 0; i < nsym; i++) lengths[i] = 0;

    int nz = 0;
    for (int i = 0; i < nsym; i++)
        if (freqs[i] != 0) nz++;
    const int k = n - 1;
    const int d = k - ((nz - 1) % k); /* C remainder semantics, exactly as written */
    const int nleaf = nz + d;
    const int ninternal = (nleaf - 1) / k;

    orc_leaf *leaf = (orc_leaf *)malloc(sizeof(orc_leaf) * (size_t)nleaf);
    int *leaf_parent = (int *)malloc(sizeof(int) * (size_t)nleaf);
    uint64_t *icount = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(ninternal + 1));
    int *iparent = (int *)malloc(sizeof(int) * (size_t)(ninternal + 1));
    int *idepth = (int *)malloc(sizeof(int) * (size_t)(ninternal + 1));
    if (!leaf || !leaf_parent || !icount || !iparent || !idepth) {
        free(leaf); free(leaf_parent); free(icount); free(iparent); free(idepth);
        return ORC_ERR_ARG;
    }
    int m = 0;
    for (int i = 0; i < nsym; i++)
        if (freqs[i] != 0) { leaf[m].count = freqs[i]; leaf[m].index = i; m++; }
    for (int j = 0; j < d; j++) { leaf[m].count = 1; leaf[m].index = nsym + j; m++; }
    qsort(leaf, (size_t)nleaf, sizeof(orc_leaf), leaf_cmp);

    int lh = 0, ih = 0, it = 0, remaining = nleaf;
    while (remaining > 1) {
        uint64_t sum = 0;
        for (int c = 0; c < n; c++) {
            const int leaf_ok = lh < nleaf, int_ok = ih < it;
            if (leaf_ok && (!int_ok || leaf[lh].count <= icount[ih])) {
                sum += leaf[lh].count;
                leaf_parent[lh++] = it;
            } else {
                sum += icount[ih];
                iparent[ih++] = it;
            }
        }
        icount[it++] = sum;
        remaining -= k;
    }
    if (it > 0) {
        idepth[it - 1] = 0;
        for (int i = it - 2; i >= 0; i--) idepth[i] = idepth[iparent[i]] + 1;
        for (int i = 0; i < nleaf; i++)
            if (leaf[i].index < nsym) lengths[leaf[i].index] = idepth[leaf_parent[i]] + 1;
    }
    free(leaf); free(leaf_parent); free(icount); free(iparent); free(idepth);
    return ORC_OK;
}

/* ------------------------------------------------------------------ lengths -> canonical values */

/*
 * convert_lengths_to_encode_table() :1382-1612.  array_max/array_min (:1330-1379) and the clearing
 * loop (:1421) scan i < max_symbol_value (the last slot is skipped); the assignment loop (:1547) scans
 * i <= max_symbol_value.  current_code is a C int (:1540); the reference asserts max_len < 16 (:1414).
 */
int orc_convert_lengths_to_encode_table(int max_symbol_value, const int lengths[], int compressed_symbols,
                                        int encode_length_table[], unsigned int encode_value_table[]) {
    int status = ORC_OK;
    int max_len = 0;
    for (int i = 0; i < max_symbol_value; i++)
        if (lengths[i] > max_len) max_len = lengths[i];
    int min_len = 300;
    for (int i = 0; i < max_symbol_value; i++)
        if (lengths[i] != 0 && lengths[i] < min_len) min_len = lengths[i];
    if (max_len >= 16) status = ORC_ERR_CODE_TOO_LONG;
    for (int i = 0; i < max_symbol_value; i++) {
        encode_length_table[i] = 0;
        encode_value_table[i] = 0;
    }
    unsigned long long code = 0;
    for (int cl = min_len; cl <= max_len; cl++) {
        for (int i = 0; i <= max_symbol_value; i++) {
            if (cl == lengths[i]) {
                if (code > (unsigned long long)INT_MAX) status = ORC_ERR_CODE_TOO_LONG;
                encode_length_table[i] = cl;
                encode_value_table[i] = (unsigned int)code;
                code += 1;
            }
        }
        code *= (unsigned long long)compressed_symbols;
        if (code > (1ull << 40)) { status = ORC_ERR_CODE_TOO_LONG; code &= (1ull << 40) - 1; }
    }
    return status;
}

int orc_bits_per_digit(int n) { return n == 2 ? 1 : n == 4 ? 2 : n == 16 ? 4 : 0; }

/* ------------------------------------------------------------------ payload packer (repo-defined) */

static int code_bits(const int elen[], int bpd, unsigned s) { return elen[s] * bpd; }

int orc_pack(const uint8_t *in, size_t n, const int elen[], const unsigned int eval[], int bits_per_digit,
             unsigned bit_phase, uint8_t *out, size_t out_capacity, uint64_t *total_bits) {
    if (bits_per_digit <= 0 || bit_phase > 7) return ORC_ERR_ARG;
    uint64_t acc = 0, bits = 0;
    int nacc = (int)bit_phase;
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        const unsigned s = in[i];
        const int l = code_bits(elen, bits_per_digit, s);
        if (l <= 0) return ORC_ERR_SYMBOL;
        if (l > 32) return ORC_ERR_CODE_TOO_LONG;
        acc = (acc << l) | (uint64_t)eval[s];
        nacc += l;
        bits += (uint64_t)l;
        while (nacc >= 8) {
            if (o >= out_capacity) return ORC_ERR_CAPACITY;
            out[o++] = (uint8_t)(acc >> (nacc - 8));
            nacc -= 8;
        }
        acc &= 0xFF;
    }
    if (bits != 0 && nacc > 0) {
        if (o >= out_capacity) return ORC_ERR_CAPACITY;
        out[o++] = (uint8_t)(acc << (8 - nacc));
    }
    if (total_bits) *total_bits = bits;
    return ORC_OK;
}

/* OR `n` symbols' codes into a zero-initialised buffer at absolute bit position `bitpos`; first and last
 * partially owned bytes use atomic OR so neighbouring blocks may run concurrently. */
static int pack_at(const uint8_t *in, size_t n, const int elen[], const unsigned int eval[], int bpd,
                   uint8_t *out, size_t cap, uint64_t bitpos) {
    size_t o = (size_t)(bitpos >> 3);
    int nacc = (int)(bitpos & 7);
    uint64_t acc = 0;
    size_t i = 0;
    /* lead-in: until the byte shared with the previous block has been emitted (atomic OR) */
    int first = nacc != 0;
    for (; i < n && first; i++) {
        const unsigned s = in[i];
        const int l = code_bits(elen, bpd, s);
        if (l <= 0) return ORC_ERR_SYMBOL;
        if (l > 32) return ORC_ERR_CODE_TOO_LONG;
        acc = (acc << l) | (uint64_t)eval[s];
        nacc += l;
        while (nacc >= 8) {
            if (o >= cap) return ORC_ERR_CAPACITY;
            const uint8_t b = (uint8_t)(acc >> (nacc - 8));
            if (first) { __atomic_fetch_or(&out[o], b, __ATOMIC_RELAXED); first = 0; }
            else out[o] = b;
            o++;
            nacc -= 8;
        }
    }
    /* main: bytes from here on are owned by this block alone; flush 32 bits at a time */
    for (; i < n; i++) {
        const unsigned s = in[i];
        const int l = code_bits(elen, bpd, s);
        if (l <= 0) return ORC_ERR_SYMBOL;
        if (l > 32) return ORC_ERR_CODE_TOO_LONG;
        acc = (acc << l) | (uint64_t)eval[s];
        nacc += l;
        if (nacc >= 32) {
            if (o + 4 > cap) return ORC_ERR_CAPACITY;
            const uint32_t w = __builtin_bswap32((uint32_t)(acc >> (nacc - 32)));
            memcpy(out + o, &w, 4);
            o += 4;
            nacc -= 32;
        }
    }
    while (nacc >= 8) {
        if (o >= cap) return ORC_ERR_CAPACITY;
        const uint8_t b = (uint8_t)(acc >> (nacc - 8));
        if (first) { __atomic_fetch_or(&out[o], b, __ATOMIC_RELAXED); first = 0; }
        else out[o] = b;
        o++;
        nacc -= 8;
    }
    if (nacc > 0 && n != 0) {
        if (o >= cap) return ORC_ERR_CAPACITY;
        __atomic_fetch_or(&out[o], (uint8_t)((acc << (8 - nacc)) & 0xFF), __ATOMIC_RELAXED);
    }
    return ORC_OK;
}

int orc_pack_mt(const uint8_t *in, size_t n, const int elen[], const unsigned int eval[], int bits_per_digit,
                unsigned bit_phase, uint8_t *out, size_t out_capacity, uint64_t *total_bits, int threads,
                uint64_t *block_bit_offsets, size_t block_symbols) {
    if (bits_per_digit <= 0 || bit_phase > 7 || block_symbols == 0) return ORC_ERR_ARG;
    if (threads < 1) threads = 1;
    const size_t nblocks = (n + block_symbols - 1) / block_symbols;
    int status = ORC_OK;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long b = 0; b < (long long)nblocks; b++) {
        const size_t lo = (size_t)b * block_symbols, hi = lo + block_symbols < n ? lo + block_symbols : n;
        uint64_t bits = 0;
        for (size_t i = lo; i < hi; i++) bits += (uint64_t)code_bits(elen, bits_per_digit, in[i]);
        block_bit_offsets[b + 1] = bits;
    }
    block_bit_offsets[0] = 0;
    for (size_t b = 0; b < nblocks; b++) block_bit_offsets[b + 1] += block_bit_offsets[b];
    const uint64_t bits = block_bit_offsets[nblocks];
    const size_t need = (size_t)((bits + bit_phase + 7) >> 3);
    if (bits != 0 && need > out_capacity) return ORC_ERR_CAPACITY;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long i = 0; i < (long long)((need + 4095) / 4096); i++) {
        const size_t lo = (size_t)i * 4096, len = lo + 4096 < need ? 4096 : need - lo;
        memset(out + lo, 0, len);
    }
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long b = 0; b < (long long)nblocks; b++) {
        const size_t lo = (size_t)b * block_symbols, hi = lo + block_symbols < n ? lo + block_symbols : n;
        const int st = pack_at(in + lo, hi - lo, elen, eval, bits_per_digit, out, out_capacity,
                               block_bit_offsets[b] + bit_phase);
        if (st != ORC_OK) {
#pragma omp critical
            status = st;
        }
    }
    if (total_bits) *total_bits = bits;
    return status;
}

/* ------------------------------------------------------------------ payload decoder (repo-defined) */

typedef struct {
    int bpd, n, max_len, min_len;      /* lengths in digits */
    unsigned first_code[64];           /* canonical first code value per length (:1540-1568) */
    unsigned count[64];
    unsigned offset[64];               /* into sorted[] */
    int sorted[1024];                  /* symbols by (length, index) */
    int lut_bits;                      /* 0 = no LUT */
    uint32_t *lut;                     /* sym | bits<<16, 0xFFFFFFFF invalid */
} orc_dec;

static int dec_init(orc_dec *d, int max_symbol_value, const int lengths[], int n) {
    memset(d, 0, sizeof *d);
    d->bpd = orc_bits_per_digit(n);
    d->n = n;
    if (d->bpd == 0 || max_symbol_value + 1 > 1024) return ORC_ERR_ARG;
    d->min_len = 300;
    for (int i = 0; i <= max_symbol_value; i++) {
        if (lengths[i] < 0 || lengths[i] >= 64) return ORC_ERR_CODE_TOO_LONG;
        if (lengths[i] > d->max_len) d->max_len = lengths[i];
        if (lengths[i] && lengths[i] < d->min_len) d->min_len = lengths[i];
        d->count[lengths[i]]++;
    }
    d->count[0] = 0;
    if (d->max_len * d->bpd > 32) return ORC_ERR_CODE_TOO_LONG;
    unsigned long long code = 0;
    unsigned off = 0;
    for (int l = d->min_len; l <= d->max_len; l++) {
        d->first_code[l] = (unsigned)code;
        d->offset[l] = off;
        unsigned k = 0;
        for (int i = 0; i <= max_symbol_value; i++)
            if (lengths[i] == l) d->sorted[off + k++] = i;
        off += d->count[l];
        code = (code + d->count[l]) * (unsigned long long)n;
    }
    const int mb = d->max_len * d->bpd;
    if (mb > 0 && mb <= 16) {
        d->lut_bits = mb;
        d->lut = (uint32_t *)malloc(sizeof(uint32_t) << mb);
        if (!d->lut) return ORC_ERR_ARG;
        memset(d->lut, 0xFF, sizeof(uint32_t) << mb);
        for (int l = d->min_len; l <= d->max_len; l++) {
            const int lb = l * d->bpd;
            for (unsigned k = 0; k < d->count[l]; k++) {
                const unsigned v = d->first_code[l] + k;
                const uint32_t e = (uint32_t)d->sorted[d->offset[l] + k] | ((uint32_t)lb << 16);
                const unsigned base = v << (mb - lb);
                for (unsigned f = 0; f < (1u << (mb - lb)); f++) d->lut[base + f] = e;
            }
        }
    }
    return ORC_OK;
}

static void dec_free(orc_dec *d) { free(d->lut); d->lut = NULL; }

/* k (<=32) bits MSB-first starting at absolute bit position pos; bits at or beyond `limit_bits` read as 0 */
static inline uint32_t peek_bits(const uint8_t *p, uint64_t pos, int k, uint64_t limit_bits) {
    uint64_t w = 0;
    const uint64_t byte0 = pos >> 3, nbytes = (limit_bits + 7) >> 3;
    if (byte0 + 8 <= nbytes) {
        memcpy(&w, p + byte0, 8);
        w = __builtin_bswap64(w);
    } else {
        for (int j = 0; j < 8; j++) {
            const uint64_t b = byte0 + (uint64_t)j;
            w = (w << 8) | (b < nbytes ? p[b] : 0);
        }
    }
    w <<= (pos & 7);
    return (uint32_t)(w >> (64 - k));
}

/* decode one code at pos; returns bits consumed (>0) and *sym, or 0 if the bits are an unused code slot */
static inline int dec_one(const orc_dec *d, const uint8_t *p, uint64_t pos, uint64_t limit, int *sym) {
    if (d->lut) {
        const uint32_t e = d->lut[peek_bits(p, pos, d->lut_bits, limit)];
        if (e == 0xFFFFFFFFu) return 0;
        *sym = (int)(e & 0xFFFF);
        return (int)(e >> 16);
    }
    const uint32_t w = peek_bits(p, pos, 32, limit);
    for (int l = d->min_len; l <= d->max_len; l++) {
        const int lb = l * d->bpd;
        const unsigned v = lb == 32 ? w : (w >> (32 - lb));
        if (d->count[l] && v >= d->first_code[l] && v - d->first_code[l] < d->count[l]) {
            *sym = d->sorted[d->offset[l] + (v - d->first_code[l])];
            return lb;
        }
    }
    return 0;
}

int orc_unpack(const uint8_t *bits, uint64_t bit_start, uint64_t nbits, int max_symbol_value,
               const int lengths[], int compressed_symbols, uint8_t *out, size_t n_out_capacity,
               size_t *n_decoded) {
    orc_dec d;
    int st = dec_init(&d, max_symbol_value, lengths, compressed_symbols);
    if (st != ORC_OK) return st;
    const uint64_t end = bit_start + nbits;
    uint64_t pos = bit_start;
    size_t o = 0;
    while (pos < end && o < n_out_capacity) {
        int sym = 0;
        const int used = dec_one(&d, bits, pos, end, &sym);
        if (used == 0 || pos + (uint64_t)used > end || sym > 255) { st = ORC_ERR_CORRUPT; break; }
        out[o++] = (uint8_t)sym;
        pos += (uint64_t)used;
    }
    if (st == ORC_OK && pos < end) st = ORC_ERR_CAPACITY;
    if (n_decoded) *n_decoded = o;
    dec_free(&d);
    return st;
}

int orc_unpack_mt(const uint8_t *bits, uint64_t bit_start, int max_symbol_value, const int lengths[],
                  int compressed_symbols, uint8_t *out, size_t n, const uint64_t *block_bit_offsets,
                  size_t block_symbols, int threads) {
    orc_dec d;
    int st = dec_init(&d, max_symbol_value, lengths, compressed_symbols);
    if (st != ORC_OK) return st;
    if (threads < 1) threads = 1;
    const size_t nblocks = (n + block_symbols - 1) / block_symbols;
    const uint64_t end = bit_start + block_bit_offsets[nblocks];
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long long b = 0; b < (long long)nblocks; b++) {
        const size_t lo = (size_t)b * block_symbols, hi = lo + block_symbols < n ? lo + block_symbols : n;
        uint64_t pos = bit_start + block_bit_offsets[b];
        for (size_t i = lo; i < hi; i++) {
            int sym = 0;
            const int used = dec_one(&d, bits, pos, end, &sym);
            if (used == 0) {
#pragma omp critical
                st = ORC_ERR_CORRUPT;
                break;
            }
            out[i] = (uint8_t)sym;
            pos += (uint64_t)used;
        }
    }
    dec_free(&d);
    return st;
}

/* ------------------------------------------------------------------ trit payload (n = 3; repo-defined, row N4) */

/*
 * The reference's default radix is 3 (n_ary_huffman.c:2529) and its author sketches the storage at :745-748: "grab 5
 * trits at a time, convert into a number 1..243, and store as an 8-bit octet (which never uses byte 0 or 244..255)".
 * Layout defined here: a code of `len` digits is the len-trit base-3 numeral of encode_value, most significant trit
 * first (as for the power-of-two radices); codes are concatenated in input order; every 5 trits t0..t4 become the byte
 * 1 + t0*81 + t1*27 + t2*9 + t3*3 + t4; the last group is padded with zero trits.
 */
int orc_pack_trits(const uint8_t *in, size_t n, const int elen[], const unsigned int eval[], uint8_t *out,
                   size_t out_capacity, uint64_t *total_trits) {
    uint64_t trits = 0;
    unsigned group = 0;
    int ng = 0;
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        const unsigned s = in[i];
        const int l = elen[s];
        if (l <= 0) return ORC_ERR_SYMBOL;
        if (l > 20) return ORC_ERR_CODE_TOO_LONG;
        unsigned digits[20];
        unsigned v = eval[s];
        for (int k = l - 1; k >= 0; k--) { digits[k] = v % 3u; v /= 3u; }
        for (int k = 0; k < l; k++) {
            group = group * 3u + digits[k];
            if (++ng == 5) {
                if (o >= out_capacity) return ORC_ERR_CAPACITY;
                out[o++] = (uint8_t)(1u + group);
                group = 0;
                ng = 0;
            }
        }
        trits += (uint64_t)l;
    }
    if (ng > 0) {
        for (; ng < 5; ng++) group *= 3u;
        if (o >= out_capacity) return ORC_ERR_CAPACITY;
        out[o++] = (uint8_t)(1u + group);
    }
    if (total_trits) *total_trits = trits;
    return ORC_OK;
}

/* sequential canonical decoder of the trit payload */
int orc_unpack_trits(const uint8_t *packed, uint64_t total_trits, int max_symbol_value, const int lengths[],
                     uint8_t *out, size_t n_out_capacity, size_t *n_decoded) {
    if (max_symbol_value + 1 > 1024) return ORC_ERR_ARG;
    int elen[1024];
    unsigned int eval[1024];
    const int st = orc_convert_lengths_to_encode_table(max_symbol_value, lengths, 3, elen, eval);
    if (st != ORC_OK) return st;
    size_t o = 0;
    unsigned v = 0;
    int l = 0;
    for (uint64_t t = 0; t < total_trits; t++) {
        const unsigned b = packed[t / 5];
        if (b == 0 || b > 243) return ORC_ERR_CORRUPT;
        static const unsigned p3[5] = {81, 27, 9, 3, 1};
        const unsigned digit = ((b - 1u) / p3[t % 5]) % 3u;
        v = v * 3u + digit;
        l++;
        int found = -1;
        for (int s = 0; s <= max_symbol_value && found < 0; s++)
            if (elen[s] == l && eval[s] == v) found = s;
        if (found >= 0) {
            if (o >= n_out_capacity) return ORC_ERR_CAPACITY;
            out[o++] = (uint8_t)found;
            v = 0;
            l = 0;
        } else if (l > 20) {
            return ORC_ERR_CORRUPT;
        }
    }
    if (l != 0) return ORC_ERR_CORRUPT;  /* the stream ends inside a code */
    if (n_decoded) *n_decoded = o;
    return ORC_OK;
}

/* ---------------------------------------------------------------- base64url text form of a binary payload (row N4)
 * int2digit() n_ary_huffman.c:371-426 (the base64url table) and digit2int() :428-455 (which also takes '+' and '/');
 * the unfinished packer (:1646-1671) emits 6 bits per character.  Character k = bits [6k, 6k + 6), most significant
 * first, zero padded. */
static const char k_b64url[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";

size_t orc_base64url_pack(const uint8_t *bits, uint64_t nbits, uint8_t *chars) {
    const uint64_t nchars = (nbits + 5) / 6;
    for (uint64_t k = 0; k < nchars; k++) {
        unsigned v = 0;
        for (int b = 0; b < 6; b++) {
            const uint64_t i = 6 * k + (uint64_t)b;
            const unsigned bit = i < nbits ? (bits[i >> 3] >> (7 - (i & 7))) & 1u : 0u;
            v = (v << 1) | bit;
        }
        chars[k] = (uint8_t)k_b64url[v];
    }
    return (size_t)nchars;
}

int orc_base64url_unpack(const uint8_t *chars, uint64_t nbits, uint8_t *bits) {
    const uint64_t nchars = (nbits + 5) / 6, nbytes = (nbits + 7) / 8;
    memset(bits, 0, (size_t)nbytes);
    for (uint64_t k = 0; k < nchars; k++) {
        const char c = (char)chars[k];
        int v = -1;
        for (int i = 0; i < 64; i++)
            if (k_b64url[i] == c) v = i;
        if (c == '+') v = 62;
        if (c == '/') v = 63;
        if (v < 0) return ORC_ERR_CORRUPT;
        for (int b = 0; b < 6; b++) {
            const uint64_t i = 6 * k + (uint64_t)b;
            if (i < 8 * nbytes && ((v >> (5 - b)) & 1)) bits[i >> 3] |= (uint8_t)(0x80u >> (i & 7));
        }
    }
    return ORC_OK;
}
