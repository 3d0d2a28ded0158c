// k4_decode.cu -- K4: self-synchronising parallel Huffman decode.
//
// The reference has no Huffman block decoder (decompress() case 'X'/'Z' is assert(0),
// n_ary_huffman.c:2081-2089; the intent is prose at :1838-1863, :1915-1928).  This decodes the payload
// layout of k3_encode.cu from the bitstream alone -- no side information besides the code lengths, the
// bit count and the symbol count.
//
// FAST PATH (three launches, no spin-waits).  The stream is cut into 256-bit subsequences (one per lane, held in
// registers), 32 of them form a warp tile (1 KB) and 16 tiles a segment (16 KB, one warp):
//   F1 synchronise+count: a warp walks its segment tile by tile.  Every lane walks its subsequence from a
//      guessed start as if a code began there, recording the position at the end of every 32-bit word; then it
//      takes the exit point of its left neighbour (shuffle) as its start and re-walks only until its position at
//      a word end equals the recorded one -- codes self-synchronise after a few symbols, so the re-walks cost a
//      fifth of the first walk.  Multi-symbol look-ups (every code that lies inside the next 12 bits) keep a walk
//      at ~0.6 table reads per symbol.  Per subsequence the exact start offset and symbol count are stored
//      (2 bytes per 32 bytes of bitstream); per segment the symbol count.  The first code of a segment is found
//      by synchronising over the LAST tile of the previous segment (1/16 extra work); that is an assumption, so
//      F2 checks it: assumed start of segment s == exit of segment s-1, for every s.
//   F2 verify + exclusive scan of the segment counts -> output offsets; total checked against n_out.
//   F3 decode+write: a warp re-walks its segment with exact starts, decodes two symbols per look-up into a
//      shared-memory staging tile and copies it out with aligned 16-byte stores.
// A stream can also be decoded chunk by chunk (DecodeChain): the pipelined host entry point at the end of this
// file uploads, decodes and downloads 32 MiB chunks concurrently.
// ROBUST PATH (the v1 kernels D1..D4 directly below, 128-bit subsequences): used when F2's check fails (a
// stream whose codes do not self-synchronise within 8192 bits); hands the tile starts over iteratively, always
// terminates.
// Unused code slots (the reference's dummy leaves) are legal on speculative paths -- they
// advance by one digit and produce no symbol -- and are reported as DC_ERR_CORRUPT on the true path.
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "chain_scan.cuh"
#include "dc_common.cuh"

namespace dc {

constexpr int kDecThreads = 256;
constexpr int kSubBits = 128;
constexpr int kTileSubs = kDecThreads;
constexpr int kTileWords = kTileSubs * kSubBits / 32;  // 1024
constexpr int kHaloWords = 8;
constexpr int kMaxSymPerSub = kSubBits;                // 1-bit codes

struct DecTables {  // shared-memory copy of the decode part of dc_huff_table
    uint16_t lut[1 << DC_LUT_BITS];
    uint32_t first_code[32], len_count[32], len_offset[32];
    uint16_t sorted[DC_NSLOTS + 1];
    int bpd, min_len, max_len, max_bits, t2;
};

// n = 3 tables (packed_radix == 3): the stream has one 2-bit field per trit.  The first `len` fields of the left-aligned
// window w as a base-3 value, for the canonical search; false if a field is 3 (not a trit).
__device__ __forceinline__ bool t2_window_value(uint32_t w, int len, uint32_t *v) {
    uint32_t r = 0;
    for (int k = 0; k < len && k < 16; k++) {
        const uint32_t d = (w >> (30 - 2 * k)) & 3u;
        if (d == 3u) return false;
        r = r * 3u + d;
    }
    *v = r;
    return true;
}

__device__ __forceinline__ void load_tables(DecTables *t, const dc_huff_table *__restrict__ tab) {
    for (int i = threadIdx.x; i < (1 << DC_LUT_BITS) / 2; i += blockDim.x)
        ((uint32_t *)t->lut)[i] = ((const uint32_t *)tab->lut)[i];
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
        t->first_code[i] = tab->first_code[i];
        t->len_count[i] = tab->len_count[i];
        t->len_offset[i] = tab->len_offset[i];
    }
    for (int i = threadIdx.x; i <= DC_NSLOTS; i += blockDim.x) t->sorted[i] = tab->sorted[i];
    if (threadIdx.x == 0) {
        t->bpd = tab->bits_per_digit;
        t->min_len = tab->min_len;
        t->max_len = tab->max_len;
        t->max_bits = tab->max_bits;
        t->t2 = tab->packed_radix == 3;
    }
}

// decode the code whose bits are left-aligned in w; returns its bit length, 0 for an unused slot
__device__ __forceinline__ int decode_one(const DecTables *t, uint32_t w, int *sym) {
    const uint32_t e = t->lut[w >> (32 - DC_LUT_BITS)];
    if (e) {
        *sym = (int)(e & 0xFFu);
        return (int)(e >> 8);
    }
    if (t->max_bits > DC_LUT_BITS) {
        for (int l = t->min_len; l <= t->max_len; l++) {
            const int lb = l * t->bpd;
            if (lb <= DC_LUT_BITS) continue;
            uint32_t v = lb >= 32 ? w : (w >> (32 - lb));
            if (t->t2 && !t2_window_value(w, l, &v)) continue;
            const uint32_t f = t->first_code[l], c = t->len_count[l];
            if (c && v >= f && v - f < c) {
                *sym = (int)t->sorted[t->len_offset[l] + (v - f)];
                return lb;
            }
        }
    }
    return 0;
}

struct SmemBits {  // a tile of big-endian words in shared memory; `base` = global bit position of word 0
    const uint32_t *w;
    unsigned long long base;
    __device__ __forceinline__ uint32_t peek(unsigned long long p) const {
        const uint32_t q = (uint32_t)(p - base);
        return __funnelshift_l(w[(q >> 5) + 1], w[q >> 5], q & 31);
    }
};
struct GmemBits {  // the bitstream in global memory (little-endian u32 loads, swapped)
    const uint32_t *g;
    unsigned long long nwords;
    __device__ __forceinline__ uint32_t word(unsigned long long i) const { return i < nwords ? bswap32(__ldg(g + i)) : 0u; }
    __device__ __forceinline__ uint32_t peek(unsigned long long p) const {
        const unsigned long long a = p >> 5;
        return __funnelshift_l(word(a + 1), word(a), (uint32_t)(p & 31));
    }
};

// decode subsequence [sub_begin, sub_begin+128) from sub_begin+start; the code crossing the end belongs to it
template <bool WRITE, typename Bits>
__device__ __forceinline__ void decode_sub(const DecTables *t, const Bits &bits, unsigned long long sub_begin, uint32_t start,
                                           unsigned long long end, uint32_t *exit_off, uint32_t *count, uint8_t *dst,
                                           bool *corrupt) {
    const unsigned long long sub_end = sub_begin + kSubBits;
    const unsigned long long limit = sub_end < end ? sub_end : end;
    unsigned long long p = sub_begin + start;
    uint32_t c = 0;
    while (p < limit) {
        int sym = 0;
        int nb = decode_one(t, bits.peek(p), &sym);
        if (nb == 0) {
            nb = t->bpd;
            if (WRITE) *corrupt = true;
        } else {
            if (WRITE) dst[c] = (uint8_t)sym;
            c++;
        }
        p += (unsigned long long)nb;
    }
    if (WRITE && p > end) *corrupt = true;  // the stream ends inside a code
    *exit_off = p >= sub_end ? (uint32_t)(p - sub_end) : 0u;
    *count = c;
}

struct DecWorkspace {
    int32_t *changed;              // D2 quiescence flag
    unsigned long long *total;     // D3 total symbol count
    uint8_t *sub_start, *sub_cnt;  // [nsub]
    uint32_t *tile_start, *tile_exit, *tile_cnt;  // [ntiles]
    unsigned long long *tile_off;  // [ntiles]
};

// load tile `tile` (+halo) of the bitstream into shared memory as big-endian words
__device__ __forceinline__ void load_tile(uint32_t *s_words, const uint8_t *__restrict__ d_bits, unsigned long long tile,
                                          unsigned long long nvec) {
    const uint4 *v = (const uint4 *)d_bits;
    for (int i = threadIdx.x; i < (kTileWords + kHaloWords) / 4; i += kDecThreads) {
        const unsigned long long gv = tile * (kTileWords / 4) + i;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (gv < nvec) x = ldg_stream(v + gv);
        s_words[4 * i + 0] = bswap32(x.x);
        s_words[4 * i + 1] = bswap32(x.y);
        s_words[4 * i + 2] = bswap32(x.z);
        s_words[4 * i + 3] = bswap32(x.w);
    }
}

// ------------------------------------------------------------------------------------------ D1
__global__ void __launch_bounds__(kDecThreads) decode_sync_kernel(const uint8_t *__restrict__ d_bits, unsigned long long bit_start,
                                                                  unsigned long long end, const dc_huff_table *__restrict__ tab,
                                                                  DecWorkspace ws, unsigned long long nsub,
                                                                  unsigned long long ntiles) {
    __shared__ DecTables s_t;
    __shared__ __align__(16) uint32_t s_words[kTileWords + kHaloWords];
    __shared__ uint32_t s_exit[kTileSubs];
    __shared__ uint32_t s_warp[kDecThreads / 32];
    load_tables(&s_t, tab);
    const unsigned long long nvec = ((end + 7) / 8 + 15) / 16;
    const int tid = threadIdx.x;
    for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        load_tile(s_words, d_bits, tile, nvec);
        __syncthreads();
        SmemBits bits = {s_words, tile * (unsigned long long)(kTileSubs * kSubBits)};
        const unsigned long long s = tile * kTileSubs + tid, sub_begin = s * kSubBits;
        const bool active = sub_begin < end;
        uint32_t my_start = s == 0 ? (uint32_t)bit_start : 0u, my_exit = 0, my_cnt = 0;
        if (active) decode_sub<false>(&s_t, bits, sub_begin, my_start, end, &my_exit, &my_cnt, nullptr, nullptr);
        s_exit[tid] = my_exit;
        __syncthreads();
        int any;
        do {
            const uint32_t ns = tid == 0 ? my_start : s_exit[tid - 1];
            const bool redo = active && ns != my_start;
            __syncthreads();
            if (redo) {
                my_start = ns;
                decode_sub<false>(&s_t, bits, sub_begin, my_start, end, &my_exit, &my_cnt, nullptr, nullptr);
                s_exit[tid] = my_exit;
            }
            any = __syncthreads_or(redo);
        } while (any);
        if (s < nsub) {
            ws.sub_start[s] = (uint8_t)my_start;
            ws.sub_cnt[s] = (uint8_t)my_cnt;
        }
        uint32_t sum = my_cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if ((tid & 31) == 0) s_warp[tid >> 5] = sum;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (int i = 0; i < kDecThreads / 32; i++) tot += s_warp[i];
            ws.tile_cnt[tile] = tot;
            ws.tile_start[tile] = tile == 0 ? (uint32_t)bit_start : 0u;
            ws.tile_exit[tile] = s_exit[kTileSubs - 1];
        }
    }
}

// ------------------------------------------------------------------------------------------ D2
__global__ void decode_handoff_kernel(const uint8_t *__restrict__ d_bits, unsigned long long end,
                                      const dc_huff_table *__restrict__ tab, DecWorkspace ws, unsigned long long ntiles) {
    __shared__ DecTables s_t;
    load_tables(&s_t, tab);
    __syncthreads();
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= ntiles) return;
    const uint32_t ns = ws.tile_exit[i - 1];
    if (ns == ws.tile_start[i]) return;
    *ws.changed = 1;
    ws.tile_start[i] = ns;
    GmemBits bits = {(const uint32_t *)d_bits, ((end + 7) / 8 + 3) / 4};
    uint32_t st = ns;
    long long delta = 0;
    for (int j = 0; j < kTileSubs; j++) {
        const unsigned long long sg = i * kTileSubs + j, sub_begin = sg * kSubBits;
        if (sub_begin >= end) break;
        uint32_t e, c;
        decode_sub<false>(&s_t, bits, sub_begin, st, end, &e, &c, nullptr, nullptr);
        delta += (long long)c - (long long)ws.sub_cnt[sg];
        ws.sub_start[sg] = (uint8_t)st;
        ws.sub_cnt[sg] = (uint8_t)c;
        if (j == kTileSubs - 1 || sub_begin + kSubBits >= end) {
            ws.tile_exit[i] = e;
            break;
        }
        if (e == ws.sub_start[sg + 1]) break;  // merged with the recorded path
        st = e;
    }
    ws.tile_cnt[i] = (uint32_t)((long long)ws.tile_cnt[i] + delta);
}

// ------------------------------------------------------------------------------------------ D3
__global__ void __launch_bounds__(1024) decode_scan_kernel(DecWorkspace ws, unsigned long long ntiles, unsigned long long n_out,
                                                           int32_t *__restrict__ d_status) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (unsigned long long base = 0; base < ntiles; base += 1024) {
        const unsigned long long i = base + tid;
        const unsigned long long c = i < ntiles ? ws.tile_cnt[i] : 0ull;
        unsigned long long incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long off = s_carry;
        for (int w = 0; w < warp; w++) off += s_warp[w];
        if (i < ntiles) ws.tile_off[i] = off + incl - c;
        __syncthreads();
        if (tid == 1023) s_carry = off + incl;
        __syncthreads();
    }
    if (tid == 0) {
        *ws.total = s_carry;
        if (s_carry != n_out) set_status(d_status, s_carry > n_out ? DC_ERR_CAPACITY : DC_ERR_CORRUPT);
    }
}

// ------------------------------------------------------------------------------------------ D4
constexpr int kDecStageBytes = kTileSubs * kMaxSymPerSub + 32;

__global__ void __launch_bounds__(kDecThreads) decode_write_kernel(const uint8_t *__restrict__ d_bits, unsigned long long end,
                                                                   const dc_huff_table *__restrict__ tab, DecWorkspace ws,
                                                                   unsigned long long nsub, unsigned long long ntiles,
                                                                   uint8_t *__restrict__ out, unsigned long long n_out,
                                                                   int32_t *__restrict__ d_status) {
    extern __shared__ __align__(16) uint8_t dec_smem[];
    DecTables *s_t = (DecTables *)dec_smem;
    uint32_t *s_words = (uint32_t *)(dec_smem + ((sizeof(DecTables) + 15) & ~(size_t)15));
    uint8_t *s_out = (uint8_t *)(s_words + kTileWords + kHaloWords);
    __shared__ uint32_t s_warp[kDecThreads / 32];
    load_tables(s_t, tab);
    const unsigned long long nvec = ((end + 7) / 8 + 15) / 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        load_tile(s_words, d_bits, tile, nvec);
        const unsigned long long s = tile * kTileSubs + tid, sub_begin = s * kSubBits;
        const bool active = s < nsub && sub_begin < end;
        const uint32_t my_start = active ? ws.sub_start[s] : 0u;
        const uint32_t my_cnt = active ? ws.sub_cnt[s] : 0u;
        uint32_t incl = my_cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t off = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kDecThreads / 32; w++) {
            const uint32_t t = s_warp[w];
            if (w < warp) off += t;
            tile_total += t;
        }
        off += incl - my_cnt;
        const unsigned long long ob = ws.tile_off[tile];
        const uint32_t a = (uint32_t)(((uintptr_t)out + ob) & 15);
        if (active) {
            SmemBits bits = {s_words, tile * (unsigned long long)(kTileSubs * kSubBits)};
            uint32_t e, c;
            bool corrupt = false;
            decode_sub<true>(s_t, bits, sub_begin, my_start, end, &e, &c, s_out + a + off, &corrupt);
            if (corrupt || c != my_cnt) set_status(d_status, DC_ERR_CORRUPT);
        }
        __syncthreads();
        // copy-out: staging byte i <-> out[ob - a + i]; 16-byte words are aligned on both sides
        const uint32_t span = a + tile_total;
        for (uint32_t j = tid; j * 16 < span; j += kDecThreads) {
            const uint32_t lo = j * 16, hi = lo + 16;
            const unsigned long long g = ob - a + lo;  // global index of staging byte lo
            if (lo >= a && hi <= span && g + 16 <= n_out) {
                stg_stream((uint4 *)(out + g), *(const uint4 *)(s_out + lo));
            } else {
                for (uint32_t k = lo < a ? a : lo; k < hi && k < span; k++)
                    if (ob - a + k < n_out) out[ob - a + k] = s_out[k];
            }
        }
    }
}

// ================================================================================================ fast path

constexpr int kF_Threads = 256;
constexpr int kF_Warps = kF_Threads / 32;
constexpr int kF_SubBits = 256;                    // one lane's subsequence
constexpr int kF_SubWords = kF_SubBits / 32;
constexpr int kF_TileVecs = 32 * kF_SubBits / 128; // 16-byte vectors per warp tile (1 KB)
constexpr int kF_SegTiles = 16;                    // 16 KB of bitstream per warp

// decode modes of the fast path (the kernels' template parameter MODE)
//   0  every code fits the 12-bit index: no escape branch
//   1  ESC: codes of 13..16 bits through the second-level table, longer ones by canonical search
//   2  T2 (radix 3, one 2-bit field per trit): the index is the base-3 value of the next 8 trits, every code fits
//   3  T2 + ESC: codes of more than 8 trits by canonical search
//   4  F1 only: tables whose longest code has 13 or 14 bits count through the 14-bit-indexed u16 table (no escapes; an
//      escape test inside the look-up loop costs F1 a third of its speed, F3 nothing: it is bound by the LSU pipe)
__host__ __device__ constexpr bool mode_esc(int m) { return (m & 1) != 0; }
__host__ __device__ constexpr bool mode_t2(int m) { return (m & 2) != 0; }
__host__ __device__ constexpr int mode_window(int m) { return m == 4 ? DC_LUT14_BITS : mode_t2(m) ? 2 * DC_TRIT_WINDOW : DC_LUT_BITS; }  // bits a look-up sees
__host__ __device__ constexpr int mode_lut_words(int m) { return m == 4 ? (1 << DC_LUT14_BITS) / 2 : mode_t2(m) ? DC_LUT_ENTRIES : (1 << DC_LUT_BITS); }
__host__ __device__ constexpr int mode_count_shift(int m) { return m == 4 ? 8 : 16; }   // where a count entry keeps its number of codes
constexpr int kTritLutEntries = 6561;  // 3^DC_TRIT_WINDOW

struct FastHeader {  // the canonical arrays for the escape path
    uint32_t first_code[32], len_count[32], len_offset[32];
    uint16_t sorted[DC_NSLOTS + 1];
    int bpd, min_len, max_len, t2;
};
template <int MODE>
struct FastTables {  // shared-memory copy: one multi-symbol LUT (+ the second level for MODE 1) + the header
    FastHeader h;
    uint32_t lut[mode_lut_words(MODE)];
    uint16_t lut2[MODE == 1 ? (DC_LUT2_SUBTABLES + 1) * 16 : 8];  // codes of 13..16 bits; the last sub-table is empty
};
static size_t fast_tables_bytes(int mode, int lut2_tables) {
    const size_t b = mode == 0 ? sizeof(FastTables<0>) : mode == 1 ? offsetof(FastTables<1>, lut2) + (size_t)lut2_tables * 32
                   : mode == 2 ? sizeof(FastTables<2>) : sizeof(FastTables<3>);
    return (b + 15) & ~(size_t)15;
}

template <int MODE>
__device__ __forceinline__ void load_fast_tables(FastTables<MODE> *t, const dc_huff_table *__restrict__ tab, const uint32_t *lut) {
    constexpr int kWords = MODE == 4 ? mode_lut_words(4) : mode_t2(MODE) ? kTritLutEntries : (1 << DC_LUT_BITS);
    for (int i = threadIdx.x; i < kWords; i += blockDim.x) t->lut[i] = lut[i];
    if (MODE == 1) {
        const int words = min(tab->lut2_used, DC_LUT2_SUBTABLES) * 8;   // 16 u16 entries per table
        for (int i = threadIdx.x; i < words; i += blockDim.x) ((uint32_t *)t->lut2)[i] = ((const uint32_t *)tab->lut2)[i];
    }
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
        t->h.first_code[i] = tab->first_code[i];
        t->h.len_count[i] = tab->len_count[i];
        t->h.len_offset[i] = tab->len_offset[i];
    }
    for (int i = threadIdx.x; i <= DC_NSLOTS; i += blockDim.x) t->h.sorted[i] = tab->sorted[i];
    if (threadIdx.x == 0) {
        t->h.bpd = tab->bits_per_digit;
        t->h.min_len = tab->min_len;
        t->h.max_len = tab->max_len;
        t->h.t2 = tab->packed_radix == 3;
    }
}

// escape path: the index window holds no complete code.  Canonical search over the lengths LONGER than the LUT index
// (a shorter code would have been in the LUT); returns the code's bits, 0 for an unused slot.
__device__ __noinline__ int decode_escape(const FastHeader *t, uint32_t w, int *sym) {
    const int bpd = t->bpd;
    int l = (t->t2 ? DC_TRIT_WINDOW : DC_LUT_BITS / bpd) + 1;
    if (l < t->min_len) l = t->min_len;
    uint32_t v3 = 0;   // radix 3: the base-3 value of the first k two-bit fields, extended as l grows
    int k = 0;
    for (; l <= t->max_len; l++) {
        const int lb = l * bpd;
        uint32_t v;
        if (t->t2) {
            for (; k < l && k < 16; k++) {
                const uint32_t d = (w >> (30 - 2 * k)) & 3u;
                if (d == 3u) return 0;  // not a trit: no code of this or a greater length starts here
                v3 = v3 * 3u + d;
            }
            v = v3;
        } else {
            v = lb >= 32 ? w : (w >> (32 - lb));
        }
        const uint32_t f = t->first_code[l], c = t->len_count[l];
        if (c && v >= f && v - f < c) {
            *sym = (int)t->sorted[t->len_offset[l] + (v - f)];
            return lb;
        }
    }
    return 0;
}

// a look-up that did not resolve (ESC tables): e names a second-level table indexed by the 4 bits behind the window, or
// none (DC_LUT_NO_SUBTABLE: canonical search).  Returns the code's bits (0 = unused slot).
template <int MODE>
__device__ __forceinline__ int escape_code(const FastTables<MODE> *t, uint32_t e, uint32_t x, int *sym) {
    if (MODE == 1 && (e & 0xFFFFu) != DC_LUT_NO_SUBTABLE) {
        const uint32_t e2 = t->lut2[((e & 0xFFFFu) << 4) | ((x >> (28 - DC_LUT_BITS)) & 15u)];
        if (e2) {
            *sym = (int)(e2 & 0xFFu);
            return (int)(e2 >> 8);
        }
    }
    return decode_escape(&t->h, x, sym);
}
__device__ __forceinline__ bool is_escape_count(uint32_t e) { return (int32_t)e < 0; }       // DC_LUT_COUNT_MARK | x
__device__ __forceinline__ bool is_escape_pair(uint32_t e) { return e >= DC_LUT_PAIR_MARK; }  // count field == 3

// LUT entry for the window in the top 12 bits of x: base + (x >> 20) * 4 as one shift and one multiply-add (written in
// PTX so that it is not canonicalised back into shift, mask and add) in front of the LDS
// T2: the window's top 16 bits are 8 two-bit trits; their base-3 value by three SWAR steps (pairs: 4a + b - a,
// then 16a + b - 7a, then 256a + b - 175a), clamped so that a field of 3 (not a trit) cannot leave the table
template <int MODE>
__device__ __forceinline__ uint32_t lds_lut(uint32_t lut, uint32_t x) {
    uint32_t v, idx;
    if (MODE == 4) {  // u16 entries, 14-bit index
        asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %1, 2, %2;\n\tld.shared.u16 %0, [a];\n\t}" : "=r"(v) : "r"(x >> (32 - DC_LUT14_BITS)), "r"(lut));
        return v;
    }
    if (mode_t2(MODE)) {
        const uint32_t s1 = x - ((x >> 2) & 0x33330000u);
        const uint32_t s2 = s1 - 7u * ((s1 >> 4) & 0x0F0F0000u);
        idx = min((s2 >> 16) - 175u * (s2 >> 24), (uint32_t)(kTritLutEntries - 1));
    } else {
        idx = x >> (32 - DC_LUT_BITS);
    }
    asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %1, 4, %2;\n\tld.shared.u32 %0, [a];\n\t}" : "=r"(v) : "r"(idx), "r"(lut));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t shared_addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr));
    return v;
}
__device__ __forceinline__ void ldg_256(const void *p, uint4 &a, uint4 &b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
// c + byte `b` of e (IDP.4A: one instruction instead of shift + mask + add); b is a compile-time constant after unrolling
__device__ __forceinline__ uint32_t add_byte(uint32_t e, int b, uint32_t c) { return __dp4a(e, 1u << (8 * b), c); }
// replace byte `b` of word with the low byte of v (PRMT)
__device__ __forceinline__ uint32_t put_byte(uint32_t word, int b, uint32_t v) {
    return __byte_perm(word, v, (0x3210u & ~(0xFu << (4 * b))) | (0x4u << (4 * b)));
}

// One lane's subsequence lives in registers: w[0..7] are its 256 bits (big-endian words), w[8] the 32 bits that
// follow.  Positions are relative to the subsequence; word k is indexed statically, so the walk is one loop per
// word.  ESC = the table has codes longer than the 12 index bits (entry 0 = escape to the canonical search).
//
// What a walk leaves behind (the lane's record of its current path):
//   chk   byte k = low 8 bits of the position after the multi-symbol loop of word k
//   wc    byte k = symbols counted in word k (word 7 also holds the single-code steps at the end)
//   exit  bits by which the last code overhangs the subsequence
// A re-walk from a corrected start updates the record word by word and stops at the first word whose end
// position equals the recorded one: from there on the two paths are the same path, and the rest of the record
// (and the exit) stands.  The symbol count is the sum of the wc bytes.
struct SyncRecord {
    uint32_t chk[2], wc[2], exit;
    __device__ __forceinline__ uint32_t count() const { return __dp4a(wc[0], 0x01010101u, __dp4a(wc[1], 0x01010101u, 0u)); }
};

template <int MODE>
__device__ __forceinline__ void sync_lookup_multi(const FastTables<MODE> *t, uint32_t lut, uint32_t hi, uint32_t lo, uint32_t &p, uint32_t &csum) {
    const uint32_t x = __funnelshift_l(lo, hi, p);
    const uint32_t e = lds_lut<MODE>(lut, x);
    if (mode_esc(MODE) && is_escape_count(e)) {
        int sym;
        const int nb = escape_code(t, e, x, &sym);
        p += nb ? nb : t->h.bpd;
        csum += nb ? 0x10000u : 0u;
    } else if (MODE == 4) {
        p = add_byte(e, 0, p);
        csum += e & 0x0FFFu;    // byte 1 accumulates the number of codes
    } else {
        p = add_byte(e, 0, p);  // bits of every code inside the window
        csum += e;              // byte 2 accumulates their number (byte 0 only carries into the unused byte 1)
    }
}
// radix 3: the index is computed from ALL eight 2-bit fields of the window, so bits behind the end of the stream (whatever
// the caller's buffer holds there; a field of 3 carries into the trits in front of it) must not reach it.  Only the
// one-code-at-a-time steps of a ragged last tile can look past the end; `end_rel` = bits from the start of the lane's
// subsequence to the end of the STREAM (a code may well reach into the next subsequence).
template <int MODE>
__device__ __forceinline__ uint32_t window_at(uint32_t hi, uint32_t lo, uint32_t p, uint32_t end_rel) {
    uint32_t x = __funnelshift_l(lo, hi, p);
    if (mode_t2(MODE) && end_rel != 0xFFFFFFFFu) {
        const uint32_t valid = end_rel - p;
        if (valid < 32u) x &= ~(0xFFFFFFFFu >> valid);
    }
    return x;
}

template <int MODE>
__device__ __forceinline__ void sync_lookup_single(const FastTables<MODE> *t, uint32_t lut, uint32_t hi, uint32_t lo, uint32_t &p, uint32_t &scnt,
                                                   uint32_t end_rel = 0xFFFFFFFFu) {
    const uint32_t x = window_at<MODE>(hi, lo, p, end_rel);
    const uint32_t e = lds_lut<MODE>(lut, x);
    if (mode_esc(MODE) && is_escape_count(e)) {
        int sym;
        const int nb = escape_code(t, e, x, &sym);
        p += nb ? nb : t->h.bpd;
        scnt += nb ? 1u : 0u;
    } else if (MODE == 4) {
        p += e >> 12;
        scnt += (e & 0x0F00u) ? 1u : 0u;
    } else {
        p += e >> 24;                            // the first code only
        scnt += (e & 0x00FF0000u) ? 1u : 0u;     // an unused slot counts nothing
    }
}

// FIRST: walk the whole subsequence and write the record.  !FIRST: re-walk from `p` until the path merges
// with the recorded one (`merged` lanes take no part; the warp leaves as soon as every lane has merged).
// FULL: lim == 256 for every lane of the warp (all tiles but the stream's last).
template <int MODE, bool FIRST, bool FULL>
__device__ __forceinline__ void sync_walk(const FastTables<MODE> *t, uint32_t lut, const uint32_t (&w)[kF_SubWords + 1], uint32_t p, uint32_t lim,
                                          bool merged, SyncRecord &r, uint32_t end_rel = 0xFFFFFFFFu) {
    constexpr int kWin = mode_window(MODE);
    const int multi_lim = (int)lim - kWin;  // every code inside the index window then starts before lim
#pragma unroll
    for (int k = 0; k < kF_SubWords; k++) {
        if (!FIRST && !__any_sync(0xFFFFFFFFu, !merged)) return;
        uint32_t csum = 0;
        if (FIRST || !merged) {
            const int stop = FULL ? (k == kF_SubWords - 1 ? kF_SubBits - kWin : 32 * (k + 1) - 1) : min(32 * (k + 1) - 1, multi_lim);
            while ((int)p <= stop) sync_lookup_multi<MODE>(t, lut, w[k], w[k + 1], p, csum);
        }
        if (k == kF_SubWords - 1 && (FIRST || !merged)) {
            // the last few bits: one code at a time (in a short subsequence they may begin in any word)
            uint32_t scnt = 0;
            if (FULL) {
                while (p < (uint32_t)kF_SubBits) sync_lookup_single<MODE>(t, lut, w[k], w[k + 1], p, scnt);
            } else {
#pragma unroll
                for (int j = 0; j < kF_SubWords; j++) {
                    const int stop1 = min(32 * (j + 1), (int)lim) - 1;
                    while ((int)p <= stop1) sync_lookup_single<MODE>(t, lut, w[j], w[j + 1], p, scnt, end_rel);
                }
            }
            csum += scnt << mode_count_shift(MODE);
        }
        const uint32_t cw = csum >> mode_count_shift(MODE);  // the count byte (only its low byte is kept: at most 46 symbols per word)
        if (FIRST) {
            if (k < 4) { r.chk[0] = put_byte(r.chk[0], k & 3, p); r.wc[0] = put_byte(r.wc[0], k & 3, cw); }
            else       { r.chk[1] = put_byte(r.chk[1], k & 3, p); r.wc[1] = put_byte(r.wc[1], k & 3, cw); }
        } else if (!merged) {
            const uint32_t old = k < 4 ? r.chk[0] : r.chk[1];
            const uint32_t neu = put_byte(old, k & 3, p);
            if (k < 4) { r.chk[0] = neu; r.wc[0] = put_byte(r.wc[0], k & 3, cw); }
            else       { r.chk[1] = neu; r.wc[1] = put_byte(r.wc[1], k & 3, cw); }
            // same position at the end of word k => same path from here on (the last word always updates the exit)
            if (k < kF_SubWords - 1 && neu == old) merged = true;
        }
        if (k == kF_SubWords - 1 && (FIRST || !merged)) r.exit = p >= (uint32_t)kF_SubBits ? p - kF_SubBits : 0u;
    }
}

// Decoding a stream in several launches (a chunk of the bitstream each, in order): what one chunk hands to the next.
struct DecodeChain {
    unsigned long long base;   // symbols decoded by the chunks before this one = output offset of this chunk
    uint32_t next_start;       // bit offset, relative to the first bit of the next chunk, of that chunk's first code
    int32_t mismatch;          // sticky: some segment of some chunk started on a wrong guess
    uint32_t first_assumed;    // the start the chunk's first segment assumed (== its given start unless it had a lead tile)
    uint32_t reserved;
};
static_assert(sizeof(DecodeChain) == sizeof(dc_shard_summary), "dc_shard_summary is the public face of DecodeChain");

struct FastWorkspace {
    uint16_t *sub_info;                              // [nsub] start offset (7 bits) | symbol count << 7
    uint32_t *seg_cnt, *seg_assumed, *seg_exit;      // [nseg]
    unsigned long long *seg_off;                     // [nseg]
    int32_t *mismatch;                               // F2: some segment started on a wrong guess
    void *fsm;                                       // byte-stepped decoder: header + F1 table + F3 table (k4_fsm.cuh)
    void *chain_slots;                               // F2: CTA totals and flags (chain_scan.cuh)
    __host__ __device__ int32_t *bad_input() const { return mismatch + 2; }  // F1 (T2 streams): a 2-bit field of 3; zeroed per launch
    __host__ __device__ uint32_t *start_slot() const { return (uint32_t *)(mismatch + 3); }  // FSM F1 -> F3: the first segment's start token
};

// A warp's view of its segment: 32-bit positions relative to the segment start, tiles streamed through
// registers one 32-byte load per lane ahead.
struct SegCursor {
    const uint4 *src;      // this lane's two vectors of the NEXT tile to fetch
    uint32_t vec_left;     // vectors of the stream from the segment start on (clamped)
    uint32_t bits_left;    // bits of the stream from the segment start on (clamped)
    uint4 n0, n1;          // prefetched vectors (little-endian words as loaded)
    uint32_t fetched;      // tiles fetched so far
    int lane;

    __device__ __forceinline__ void init(const uint8_t *d_bits, unsigned long long first_tile, unsigned long long nvec,
                                         unsigned long long end, int lane_) {
        lane = lane_;
        const unsigned long long vec0 = first_tile * kF_TileVecs, bit0 = first_tile * (unsigned long long)(32 * kF_SubBits);
        src = (const uint4 *)d_bits + vec0 + 2 * lane;
        const unsigned long long vl = nvec > vec0 ? nvec - vec0 : 0, bl = end > bit0 ? end - bit0 : 0;
        vec_left = vl > 0x40000000ull ? 0x40000000u : (uint32_t)vl;
        bits_left = bl > 0x40000000ull ? 0x40000000u : (uint32_t)bl;
        fetched = 0;
        fetch();
    }
    __device__ __forceinline__ void fetch() {
        n0 = make_uint4(0, 0, 0, 0);
        n1 = make_uint4(0, 0, 0, 0);
        const uint32_t i = fetched * kF_TileVecs + 2 * lane;
        if (i + 1 < vec_left) ldg_256(src, n0, n1);
        else if (i < vec_left) n0 = ldg_stream(src);
        src += kF_TileVecs;
        fetched++;
    }
    // words of tile number `fetched - 1` for this lane + the 32 bits that follow; prefetches the tile after it
    __device__ __forceinline__ void take(uint32_t (&w)[kF_SubWords + 1]) {
        const uint4 c0 = n0, c1 = n1;
        fetch();
        w[0] = bswap32(c0.x); w[1] = bswap32(c0.y); w[2] = bswap32(c0.z); w[3] = bswap32(c0.w);
        w[4] = bswap32(c1.x); w[5] = bswap32(c1.y); w[6] = bswap32(c1.z); w[7] = bswap32(c1.w);
        const uint32_t right = __shfl_down_sync(0xFFFFFFFFu, w[0], 1);
        const uint32_t wrap = __shfl_sync(0xFFFFFFFFu, bswap32(n0.x), 0);
        w[8] = lane == 31 ? wrap : right;
    }
};

}  // namespace dc
#include "k4_fsm.cuh"
namespace dc {

// ------------------------------------------------------------------------------------------ F1
template <int MODE>
__global__ void __launch_bounds__(kF_Threads, mode_t2(MODE) ? 3 : 4) decode_fast_sync_kernel(const uint8_t *__restrict__ d_bits, unsigned long long bit_start,
                                                                      unsigned long long end, const dc_huff_table *__restrict__ tab,
                                                                      FastWorkspace ws, unsigned long long nsub,
                                                                      unsigned long long ntiles, unsigned long long nseg,
                                                                      const DecodeChain *__restrict__ chain, int lead) {
    // lead = 1: tile 0 of d_bits is the last tile of the PREVIOUS shard of a longer stream; the first code of this
    // shard is unknown and segment 0 finds it like every other segment does, by synchronising over the tile in front
    __shared__ FastTables<MODE> s_t;
    load_fast_tables(&s_t, tab, MODE == 4 ? (const uint32_t *)tab->lut14 : tab->lut_count);
    __syncthreads();
    if (chain) bit_start = chain->next_start;  // a later chunk of a stream: its first code starts where the previous chunk's last one ended
    uint32_t lut = (uint32_t)__cvta_generic_to_shared(s_t.lut);
    asm volatile("" : "+r"(lut));  // keep the table's address in a register: otherwise it is re-derived (S2R + LEA) in every word loop
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long nvec = ((end + 7) / 8 + 15) / 16;
    const uint32_t guess = (uint32_t)(bit_start & 7);  // fixed-length-like codes keep the stream's phase
    for (unsigned long long seg = (unsigned long long)blockIdx.x * kF_Warps + warp; seg < nseg;
         seg += (unsigned long long)gridDim.x * kF_Warps) {
        const bool exact = seg == 0 && lead == 0;  // the stream's (or chunk's) first segment starts where it was told
        const int warm = exact ? 0 : 1;            // the others walk the tile in front of them first to find their first code
        const unsigned long long tile0 = (unsigned long long)lead + seg * kF_SegTiles - warm;
        SegCursor cur;
        cur.init(d_bits, tile0, nvec, end, lane);
        const uint32_t ntile = (uint32_t)min((unsigned long long)(kF_SegTiles + warm), ntiles - tile0);
        uint16_t *info = ws.sub_info + tile0 * 32 + lane;
        const unsigned long long sub0 = tile0 * 32;
        const uint32_t sub_left = nsub - sub0 > 0x40000000ull ? 0x40000000u : (uint32_t)(nsub - sub0);
        uint32_t carry = exact ? (uint32_t)bit_start : guess, assumed = carry, total = 0;
        for (uint32_t tt = 0; tt < ntile; tt++, info += 32) {
            uint32_t w[kF_SubWords + 1];
            cur.take(w);
            const uint32_t sub_bit0 = (tt * 32 + lane) * kF_SubBits;             // relative to tile0
            const bool active = sub_bit0 < cur.bits_left;
            const uint32_t lim = active ? min((uint32_t)kF_SubBits, cur.bits_left - sub_bit0) : 0u;
            const bool full = (tt + 1) * 32 * kF_SubBits <= cur.bits_left;       // warp-uniform: every lane has 256 bits
            if (mode_t2(MODE)) {  // a 2-bit field of 3 is no trit: the robust path reports it (F2 reads the flag)
                uint32_t bad3 = 0;
#pragma unroll
                for (int k = 0; k < kF_SubWords; k++) {
                    const int left = (int)lim - 32 * k;
                    const uint32_t m = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ~(0xFFFFFFFFu >> left));
                    bad3 |= w[k] & (w[k] >> 1) & 0x55555555u & m;
                }
                if (bad3) *ws.bad_input() = 1;
            }
            uint32_t start = lane == 0 ? carry : guess;
            SyncRecord r;
            r.chk[0] = r.chk[1] = r.wc[0] = r.wc[1] = r.exit = 0;
            if (full) sync_walk<MODE, true, true>(&s_t, lut, w, start, lim, false, r);
            else if (active) sync_walk<MODE, true, false>(&s_t, lut, w, start, lim, false, r, cur.bits_left - sub_bit0);
            while (true) {
                uint32_t ns = __shfl_up_sync(0xFFFFFFFFu, r.exit, 1);
                if (lane == 0) ns = start;
                const bool redo = active && ns != start;
                if (!__any_sync(0xFFFFFFFFu, redo)) break;
                start = ns;
                if (full) sync_walk<MODE, false, true>(&s_t, lut, w, start, lim, !redo, r);
                else sync_walk<MODE, false, false>(&s_t, lut, w, start, lim, !redo, r, active ? cur.bits_left - sub_bit0 : 0u);
            }
            carry = __shfl_sync(0xFFFFFFFFu, r.exit, 31);
            if (tt < (uint32_t)warm) {
                assumed = carry;
            } else {
                const uint32_t cnt = r.count();
                if (tt * 32 + lane < sub_left) *info = (uint16_t)(start | (cnt << 7));
                total += cnt;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
        if (lane == 0) {
            ws.seg_cnt[seg] = total;
            ws.seg_assumed[seg] = assumed;
            ws.seg_exit[seg] = carry;
        }
    }
}

// ------------------------------------------------------------------------------------------ F2
__global__ void __launch_bounds__(1024) decode_fast_scan_kernel(FastWorkspace ws, unsigned long long nseg, unsigned long long n_out,
                                                                int32_t *__restrict__ d_status, DecodeChain *__restrict__ chain,
                                                                int last_chunk, int check_input,
                                                                DecodeChain *__restrict__ host_chain, ChainSlots slots,
                                                                int32_t *__restrict__ host_flag) {
    // CTA b takes the contiguous segments [lo, hi); totals travel from CTA to CTA (chain_scan.cuh).  A CTA's value carries its
    // "some segment started on a wrong guess" flag above bit 56 (a symbol total stays far below).
    const int tid = threadIdx.x;
    const unsigned long long per = (nseg + gridDim.x - 1) / gridDim.x, lo = min(nseg, per * blockIdx.x), hi = min(nseg, lo + per);
    unsigned long long mine = 0;
    bool bad = false;
    for (unsigned long long i = lo + tid; i < hi; i += 1024) {
        mine += ws.seg_cnt[i];
        if (i > 0) bad |= ws.seg_assumed[i] != ws.seg_exit[i - 1];
    }
    const int any_bad = __syncthreads_or(bad);
    unsigned long long cta_total;
    block_exclusive(mine, &cta_total);
    const unsigned long long before = chain_exclusive(slots, cta_total + ((unsigned long long)(any_bad ? 1 : 0) << 56));
    const unsigned long long start = (chain ? chain->base : 0ull) + (before & ((1ull << 56) - 1));
    unsigned long long carry = start;
    for (unsigned long long base = lo; base < hi; base += 1024) {
        const unsigned long long i = base + tid;
        const unsigned long long c = i < hi ? ws.seg_cnt[i] : 0ull;
        unsigned long long trip_total;
        const unsigned long long off = block_exclusive(c, &trip_total);
        if (i < hi) ws.seg_off[i] = carry + off;
        carry += trip_total;
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {   // the last CTA has seen every total and every flag
        int s_bad = ((before >> 56) != 0 || any_bad) ? 1 : 0;
        if (check_input && *ws.bad_input()) s_bad = 1;
        *ws.mismatch = s_bad;
        if (host_flag) {   // mapped host memory (dc_huff_decode waits for this kernel's event, not for the stream)
            *host_flag = s_bad;
            __threadfence_system();
        }
        if (chain) {
            chain->base = carry;
            chain->next_start = ws.seg_exit[nseg - 1];
            chain->mismatch |= s_bad;
            chain->first_assumed = ws.seg_assumed[0];
            if (host_chain) {  // mapped host memory: the host learns the chunk's symbol count without a copy-engine transfer
                *host_chain = *chain;
                __threadfence_system();
            }
        }
        if (!s_bad && last_chunk && carry != n_out) set_status(d_status, carry > n_out ? DC_ERR_CAPACITY : DC_ERR_CORRUPT);
    }
}

// ------------------------------------------------------------------------------------------ F3
// lut_pair entry: symbol0 | symbol1 << 8 | bits of (up to) two codes << 16 | first code's bits << 24 (5 bits)
//                 | unused-slot flag << 29 | number of symbols (0..2) << 30
constexpr uint32_t kPairUnused = 1u << 29;

template <int MODE>
__device__ __forceinline__ void write_lookup_multi(const FastTables<MODE> *t, uint32_t lut, uint32_t hi, uint32_t lo, uint32_t &p, uint32_t &dst,
                                                   uint32_t &flags) {
    const uint32_t x = __funnelshift_l(lo, hi, p);
    const uint32_t e = lds_lut<MODE>(lut, x);
    if (mode_esc(MODE) && is_escape_pair(e)) {
        int sym = 0;
        const int nb = escape_code(t, e, x, &sym);
        if (nb) {
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst), "r"(sym) : "memory");
            dst++;
        } else {
            flags |= kPairUnused;
        }
        p += nb ? nb : t->h.bpd;
    } else {
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst), "r"(e) : "memory");
        if ((int)e < 0) asm volatile("st.shared.u8 [%0+1], %1;" ::"r"(dst), "r"(e >> 8) : "memory");  // two symbols
        p = add_byte(e, 2, p);
        dst += e >> 30;
        flags |= e;
    }
}
template <int MODE>
__device__ __forceinline__ void write_lookup_single(const FastTables<MODE> *t, uint32_t lut, uint32_t hi, uint32_t lo, uint32_t &p, uint32_t &dst,
                                                    uint32_t &flags, uint32_t end_rel = 0xFFFFFFFFu) {
    const uint32_t x = window_at<MODE>(hi, lo, p, end_rel);
    const uint32_t e = lds_lut<MODE>(lut, x);
    if (mode_esc(MODE) && is_escape_pair(e)) {
        int sym = 0;
        const int nb = escape_code(t, e, x, &sym);
        if (nb) {
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst), "r"(sym) : "memory");
            dst++;
        } else {
            flags |= kPairUnused;
        }
        p += nb ? nb : t->h.bpd;
    } else {
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst), "r"(e) : "memory");
        p += (e >> 24) & 31u;                 // the first code only
        dst += (e >> 30) ? 1u : 0u;           // an unused slot yields no symbol
        flags |= e;
    }
}

// decode bits [p, lim) of the lane's subsequence into shared memory at dst (a shared-space byte address)
template <int MODE, bool FULL>
__device__ __forceinline__ void write_walk(const FastTables<MODE> *t, uint32_t lut, const uint32_t (&w)[kF_SubWords + 1], uint32_t &p, uint32_t lim,
                                           uint32_t &dst, uint32_t &flags, uint32_t end_rel = 0xFFFFFFFFu) {
    constexpr int kWin = mode_window(MODE);
    const int multi_lim = (int)lim - kWin;
#pragma unroll
    for (int k = 0; k < kF_SubWords; k++) {
        const int stop = FULL ? (k == kF_SubWords - 1 ? kF_SubBits - kWin : 32 * (k + 1) - 1) : min(32 * (k + 1) - 1, multi_lim);
        while ((int)p <= stop) write_lookup_multi<MODE>(t, lut, w[k], w[k + 1], p, dst, flags);
    }
    if (FULL) {
        while (p < (uint32_t)kF_SubBits) write_lookup_single<MODE>(t, lut, w[kF_SubWords - 1], w[kF_SubWords], p, dst, flags);
    } else {
#pragma unroll
        for (int j = 0; j < kF_SubWords; j++) {
            const int stop1 = min(32 * (j + 1), (int)lim) - 1;
            while ((int)p <= stop1) write_lookup_single<MODE>(t, lut, w[j], w[j + 1], p, dst, flags, end_rel);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kF_Threads, mode_t2(MODE) ? 3 : 4) decode_fast_write_kernel(const uint8_t *__restrict__ d_bits, unsigned long long end,
                                                                       const dc_huff_table *__restrict__ tab, FastWorkspace ws,
                                                                       unsigned long long nsub, unsigned long long ntiles,
                                                                       unsigned long long nseg, uint8_t *__restrict__ out,
                                                                       unsigned long long n_out, uint32_t stage_bytes,
                                                                       int32_t *__restrict__ d_status, int lead, uint32_t tables_bytes) {
    extern __shared__ __align__(16) uint8_t fast_smem[];
    FastTables<MODE> *s_t = (FastTables<MODE> *)fast_smem;
    uint8_t *s_stage = fast_smem + tables_bytes;  // MODE 1: only the second-level tables in use are resident
    if (*ws.mismatch) return;  // the robust path redoes the stream
    load_fast_tables(s_t, tab, tab->lut_pair);
    __syncthreads();
    uint32_t lut = (uint32_t)__cvta_generic_to_shared(s_t->lut);
    asm volatile("" : "+r"(lut));  // as in F1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *stage = s_stage + (size_t)warp * stage_bytes;  // 16-byte aligned: stage_bytes is a multiple of 16
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(stage);
    const unsigned long long nvec = ((end + 7) / 8 + 15) / 16;
    bool corrupt = false;
    for (unsigned long long seg = (unsigned long long)blockIdx.x * kF_Warps + warp; seg < nseg;
         seg += (unsigned long long)gridDim.x * kF_Warps) {
        const unsigned long long tile0 = (unsigned long long)lead + seg * kF_SegTiles;
        SegCursor cur;
        cur.init(d_bits, tile0, nvec, end, lane);
        const uint32_t ntile = (uint32_t)min((unsigned long long)kF_SegTiles, ntiles - tile0);
        const uint16_t *info = ws.sub_info + tile0 * 32 + lane;
        const unsigned long long sub0 = tile0 * 32;
        const uint32_t sub_left = nsub - sub0 > 0x40000000ull ? 0x40000000u : (uint32_t)(nsub - sub0);
        unsigned long long ob = ws.seg_off[seg];
        uint32_t next_info = lane < sub_left ? *info : 0u;
        for (uint32_t tt = 0; tt < ntile; tt++) {
            uint32_t w[kF_SubWords + 1];
            cur.take(w);
            const uint32_t my_info = next_info;
            info += 32;
            next_info = (tt + 1 < ntile && (tt + 1) * 32 + lane < sub_left) ? *info : 0u;
            const uint32_t start = my_info & 0x7Fu, my_cnt = my_info >> 7;
            uint32_t incl = my_cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += x;
            }
            const uint32_t tile_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const uint32_t a = (uint32_t)(((uintptr_t)out + ob) & 15);
            const uint32_t sub_bit0 = (tt * 32 + lane) * kF_SubBits;
            const bool full = (tt + 1) * 32 * kF_SubBits <= cur.bits_left && (tt + 1) * 32 <= sub_left;  // warp-uniform
            // a record that cannot be true (more symbols than the staging tile holds) must not be walked
            const bool sane = a + tile_total + 2 <= stage_bytes;
            __syncwarp();  // the previous tile's copy-out has read the staging buffer
            if (!sane) {
                corrupt = true;
            } else if (sub_bit0 < cur.bits_left && tt * 32 + lane < sub_left) {
                const uint32_t lim = min((uint32_t)kF_SubBits, cur.bits_left - sub_bit0);
                const uint32_t dst0 = stage_addr + a + (incl - my_cnt);
                uint32_t p = start, dst = dst0, flags = 0;
                if (full) write_walk<MODE, true>(s_t, lut, w, p, lim, dst, flags);
                else write_walk<MODE, false>(s_t, lut, w, p, lim, dst, flags, cur.bits_left - sub_bit0);
                if ((flags & kPairUnused) || dst - dst0 != my_cnt || p > cur.bits_left - sub_bit0) corrupt = true;
            }
            __syncwarp();
            // copy-out: staging byte i <-> out[ob - a + i]; 16-byte words are aligned on both sides
            const uint32_t span = a + tile_total;
            if (!sane) {
                // nothing to copy
            } else if (ob + tile_total <= n_out) {
                const uint32_t jfull = span >> 4;  // staging vectors [head, jfull) are complete
                const uint32_t head = a ? 1u : 0u;
                for (uint32_t j = head + lane; j < jfull; j += 32)
                    stg_stream((uint4 *)(out + (ob - a)) + j, *(const uint4 *)(stage + j * 16));
                // ragged edges, one byte per lane: lanes 0..15 the first vector, lanes 16..31 the last one
                const uint32_t k = lane < 16 ? (uint32_t)lane : jfull * 16 + (lane - 16);
                const bool edge = lane < 16 ? (a != 0 || jfull == 0) : (jfull != 0);
                if (edge && k >= a && k < span) out[ob - a + k] = stage[k];
            } else {  // more symbols than the caller expects (corrupt stream): clamp every byte
                for (uint32_t k = a + lane; k < span; k += 32)
                    if (ob - a + k < n_out) out[ob - a + k] = stage[k];
            }
            ob += tile_total;
        }
    }
    if (corrupt) set_status(d_status, DC_ERR_CORRUPT);
}

static size_t dec_ws_layout(unsigned long long bit_start, unsigned long long nbits, size_t off[12], unsigned long long *nsub_out,
                            unsigned long long *ntiles_out) {
    const unsigned long long end = bit_start + nbits;
    const unsigned long long nsub = (end + kSubBits - 1) / kSubBits;
    const unsigned long long ntiles = (nsub + kTileSubs - 1) / kTileSubs;        // robust path: 256 subsequences
    const unsigned long long nsubf = (end + kF_SubBits - 1) / kF_SubBits;                              // fast path: 256-bit subsequences
    const unsigned long long nwt = (nsubf + 31) / 32, nseg = (nwt + kF_SegTiles - 1) / kF_SegTiles;
    size_t p = 64;
    auto take = [&](size_t bytes) { size_t o = p; p += (bytes + 63) & ~(size_t)63; return o; };
    size_t o[12];
    o[0] = take(nsub);            // sub_start
    o[1] = take(nsub);            // sub_cnt
    o[2] = take(ntiles * 4);      // tile_start
    o[3] = take(ntiles * 4);      // tile_exit
    o[4] = take(ntiles * 4);      // tile_cnt
    o[5] = take(ntiles * 8);      // tile_off
    o[6] = take(nwt * 32 * 2);    // fast: sub_info
    o[7] = take(nseg * 4);        // fast: seg_cnt
    o[8] = take(nseg * 4);        // fast: seg_assumed
    o[9] = take(nseg * 4);        // fast: seg_exit
    o[10] = take(nseg * 8);       // fast: seg_off
    o[11] = take(kFsmWorkspaceBytes + kChainSlotsBytes);  // byte-stepped decoder: its tables (built per call from the code table); F2's chain slots
    if (off) for (int i = 0; i < 12; i++) off[i] = o[i];
    if (nsub_out) *nsub_out = nsub;
    if (ntiles_out) *ntiles_out = ntiles;
    return p;
}

}  // namespace dc

using namespace dc;

extern "C" size_t dc_huff_decode_workspace_bytes(uint64_t bit_start, uint64_t nbits) {
    return dec_ws_layout(bit_start, nbits, nullptr, nullptr, nullptr);
}

// the v1 kernels: exact for any stream, iterates the tile hand-off until quiescent
static int decode_robust(const uint8_t *d_bits, unsigned long long bit_start, unsigned long long end, const dc_huff_table *d_table,
                         uint8_t *d_out, size_t n_out, int32_t *d_status, DecWorkspace ws, unsigned long long nsub,
                         unsigned long long ntiles, cudaStream_t st) {
    const int sms = sm_count();
    const int grid1 = (int)(ntiles < (unsigned long long)sms * 8 ? ntiles : (unsigned long long)sms * 8);
    {
        LaunchScope ls(DC_K_DECODE_SYNC, st);
        decode_sync_kernel<<<grid1, kDecThreads, 0, st>>>(d_bits, bit_start, end, d_table, ws, nsub, ntiles);
    }
    if (ntiles > 1) {
        const unsigned int hb = (unsigned int)((ntiles - 1 + 127) / 128);
        for (int iter = 0;; iter++) {
            DC_CUDA_TRY(cudaMemsetAsync(ws.changed, 0, sizeof(int32_t), st));
            {
                LaunchScope ls(DC_K_DECODE_HANDOFF, st);
                decode_handoff_kernel<<<hb, 128, 0, st>>>(d_bits, end, d_table, ws, ntiles);
            }
            int32_t changed = 0;
            DC_CUDA_TRY(cudaMemcpyAsync(&changed, ws.changed, sizeof changed, cudaMemcpyDeviceToHost, st));
            DC_CUDA_TRY(cudaStreamSynchronize(st));
            if (!changed) break;
            if ((unsigned long long)iter > ntiles + 1) return DC_ERR_CORRUPT;
        }
    }
    {
        LaunchScope ls(DC_K_DECODE_SCAN, st);
        decode_scan_kernel<<<1, 1024, 0, st>>>(ws, ntiles, n_out, d_status);
    }
    const size_t smem4 = ((sizeof(DecTables) + 15) & ~(size_t)15) + (kTileWords + kHaloWords) * 4 + kDecStageBytes;
    DC_CUDA_TRY(ensure_dynamic_smem((const void *)decode_write_kernel, smem4));
    const int grid4 = (int)(ntiles < (unsigned long long)sms * 4 ? ntiles : (unsigned long long)sms * 4);
    LaunchScope ls(DC_K_DECODE_WRITE, st);
    decode_write_kernel<<<grid4, kDecThreads, smem4, st>>>(d_bits, end, d_table, ws, nsub, ntiles, d_out, n_out, d_status);
    return cuda_status(cudaGetLastError());
}


constexpr int kSyncW14 = 0x10;
constexpr int kModeFsmSync = 0x40;   // F1 through the state machine, F3 through the window kernel (k4_fsm.cuh, COMPAT): binary codes with 256 states
constexpr int kModeFsmWide = 0x80;   // with kModeFsm: 4-bit digits, at most two codes end in a byte (the write kernel without its suffix look-up)
constexpr int kModeFsm = 0x20;   // the byte-stepped kernels of k4_fsm.cuh; bits 20..28: states, bits 29..31: min(shortest code's bits, 8) - 1
static int decode_force_mode();
static bool decode_tma();
static bool fsm_rows_fit(const int32_t *tmeta) {
    const int min_bits = tmeta[5] * tmeta[1] > 0 ? tmeta[5] * tmeta[1] : 1;
    const size_t per_lane = (size_t)(kF_SubBits / (min_bits < 8 ? min_bits : 8)) + 2, half = 16 * per_lane + 48, whole = 32 * per_lane + 48;
    const size_t want = half > 2176 ? half : 2176, stage = ((want < whole ? want : whole) + 15) & ~(size_t)15;
    return kFsmHeaderBytes + (size_t)(tmeta[11] + 1 + kFsmEntryRows) * kFsmWriteRowBytes + kFsmWriteXTableBytes + 24 * stage + 256 <= 227 * 1024 - 1024;
}
// which instantiation of the fast kernels a table takes (tmeta = the first ten words of dc_huff_table, lut2_used, fsm_states)
static int fast_mode(const int32_t *tmeta) {
    int mode;
    if (tmeta[9] == 3) {
        mode = tmeta[6] > DC_TRIT_WINDOW ? 3 : 2;   // radix 3: max_len in trits against the 8-trit index
    } else if (nibble_radix(tmeta[0])) {
        mode = 0;                          // (no window LUTs: only the state-machine bits below count -- see decode_mode_usable)
    } else if (tmeta[7] <= DC_LUT_BITS) {
        mode = 0;                          // max_bits against the 12-bit index
    } else {
        const int sub = tmeta[10] < 0 ? 0 : (tmeta[10] > DC_LUT2_SUBTABLES ? DC_LUT2_SUBTABLES : tmeta[10]);
        mode = (tmeta[7] <= DC_LUT14_BITS ? (1 | kSyncW14) : 1) | (sub << 8);   // 13 or 14 bits: F1 counts through the 14-bit table (K2 fills it exactly then)
    }
    // (a table whose write-pass rows do not all fit shared memory keeps the window kernels: a row read from global memory
    // sits in the walk's dependent chain, and with 32 lanes some lane needs one in most steps -- measured 2.5x slower)
    if (tmeta[11] > 0 && tmeta[11] <= kFsmMaxStates && fsm_rows_fit(tmeta) && decode_force_mode() != 3) {
        const int min_bits = tmeta[5] * tmeta[1];
        mode |= kModeFsm | (tmeta[1] == 4 ? kModeFsmWide : 0) | (tmeta[11] << 20) | ((min_bits < 8 ? min_bits : 8) - 1) << 29;
    } else if (tmeta[11] > 0 && tmeta[11] <= kFsmMaxSyncStates && tmeta[1] == 1 && tmeta[9] == 0 && decode_force_mode() != 3) {
        // binary code, too many states for F3's rows: the state machine for F1 only (its table is at most 128 KB)
        const int min_bits = tmeta[5] * tmeta[1];
        mode |= kModeFsmSync | (tmeta[11] << 20) | ((min_bits < 8 ? min_bits : 8) - 1) << 29;
    }
    return mode;
}
// nibble-per-digit radices (5 .. 15) have no window LUTs: only the byte-stepped kernels can decode them
static bool decode_mode_usable(const int32_t *tmeta, int mode) { return !nibble_radix(tmeta[0]) || (mode & kModeFsm) != 0; }
static int write_mode(int mode) { return mode & 3; }
static int lut2_tables(int mode) { return (mode >> 8) & 0x1FF; }   // second-level tables in use (F3 keeps only those in shared memory)
static int sync_mode(int mode) { return (mode & kSyncW14) ? 4 : (mode & 3); }
static int fsm_states(int mode) { return (mode >> 20) & 0x1FF; }
static int fsm_min_bits(int mode) { return (int)(((unsigned)mode >> 29) & 7u) + 1; }
// a stream whose first code does not start on a digit boundary of the byte grid keeps the window kernels
static int mode_for_start(int mode, int bpd, unsigned long long start) {
    if ((mode & (kModeFsm | kModeFsmSync)) && !(start & kFsmToken) && bpd > 0 && start % (unsigned)bpd != 0) mode &= ~(kModeFsm | kModeFsmSync);
    return mode;
}

// staging tile per warp of the write kernel: a lane decodes at most 256 / min_bits symbols, plus the code that crosses its end
static uint32_t fast_stage_bytes(const int32_t *tmeta) {
    const int min_bits = tmeta[5] * tmeta[1] > 0 ? tmeta[5] * tmeta[1] : 1;
    return (uint32_t)((32 * (kF_SubBits / min_bits + 2) + 64 + 15) & ~15);
}

// the write kernel's dynamic shared memory limit only ever grows (the attribute is a limit, not a request: setting it
// lower for one table would make the next launch for a table with shorter codes fail)
static cudaError_t ensure_write_smem(size_t smem3) {
    cudaError_t e = ensure_dynamic_smem((const void *)decode_fast_write_kernel<0>, smem3);
    if (e == cudaSuccess) e = ensure_dynamic_smem((const void *)decode_fast_write_kernel<1>, smem3);
    if (e == cudaSuccess) e = ensure_dynamic_smem((const void *)decode_fast_write_kernel<2>, smem3);
    if (e == cudaSuccess) e = ensure_dynamic_smem((const void *)decode_fast_write_kernel<3>, smem3);
    return e;
}

// ---- the byte-stepped kernels (k4_fsm.cuh)
static uint32_t fsm_stage_bytes(int mode) {
    // half a tile always fits (16 lanes x the most symbols a lane can hold): a tile with more symbols than the staging tile
    // holds is walked in two halves.  That must stay the exception -- a half walk costs a whole one -- so the tile is never
    // smaller than 2176 bytes (a typical 1 KB tile decodes to 1300 .. 1500 symbols) unless a whole worst-case tile is
    const uint32_t per_lane = (uint32_t)(kF_SubBits / fsm_min_bits(mode)) + 2u;
    const uint32_t half = 16u * per_lane + 48u, whole = 32u * per_lane + 48u;
    const uint32_t want = half > 2176u ? half : 2176u;
    return ((want < whole ? want : whole) + 15u) & ~15u;
}
static void launch_fsm_build(const dc_huff_table *d_table, FastWorkspace fw, int mode, cudaStream_t st) {
    const int nstates = fsm_states(mode);
    const bool compat = (mode & kModeFsmSync) != 0;
    LaunchScope ls(DC_K_DECODE_FSM_BUILD, st);
    fsm_build_kernel<<<compat ? nstates : nstates + 1 + kFsmEntryRows + kFsmSuffixRows, 256, 0, st>>>(d_table, fsm_tables_at(fw.fsm), compat ? 1 : 0, nstates);
}
static int launch_fsm_sync(const uint8_t *d_bits, unsigned long long start, unsigned long long end, unsigned long long nsubf,
                           unsigned long long nwt, unsigned long long nseg, const dc_huff_table *d_table, FastWorkspace fw, int mode,
                           DecodeChain *chain, int lead, cudaStream_t st) {
    const int nstates = fsm_states(mode);
    const bool compat = (mode & kModeFsmSync) != 0;
    const FsmTables t = fsm_tables_at(fw.fsm);
    launch_fsm_build(d_table, fw, mode, st);
    // north_star (4): the tile staged by the bulk-copy engine instead of a 32-byte load per lane (DC_DECODE_TMA=1; measured
    // in DESIGN.md -- the default is whichever is faster)
    const bool tma = decode_tma();
    const size_t table = (size_t)(nstates + (!compat && fsm_has_entry_rows(nstates) ? 1 + kFsmEntryRows : 0)) * kFsmSyncRowBytes;
    const uint32_t table_bytes = (uint32_t)((table + 127) & ~(size_t)127);
    // as many CTAs per SM as the table allows (each keeps its own copy), 32 warps per SM in every case
    int threads = 1024, per_sm = 1;
    size_t smem = 0;
    for (int ctas = 4; ctas >= 1; ctas >>= 1) {
        const int warps = 32 / ctas;
        smem = kFsmHeaderBytes + table_bytes + (tma ? (size_t)warps * (1024 + 8) : 0);
        if ((size_t)ctas * (smem + 1024) <= 227 * 1024) { threads = warps * 32; per_sm = ctas; break; }
    }
    const void *fn = compat ? (tma ? (const void *)fsm_sync_kernel<true, true> : (const void *)fsm_sync_kernel<false, true>)
                            : (tma ? (const void *)fsm_sync_kernel<true, false> : (const void *)fsm_sync_kernel<false, false>);
    DC_CUDA_TRY(ensure_dynamic_smem(fn, smem));
    const unsigned long long want = (nseg + (threads / 32) - 1) / (threads / 32), cap = (unsigned long long)sm_count() * per_sm;
    FsmSyncArgs a;
    a.d_bits = d_bits;
    a.end = end;
    a.nsub = nsubf;
    a.ntiles = nwt;
    a.nseg = nseg;
    a.start_token = (uint32_t)start;
    a.lead = lead;
    a.chain = chain;
    LaunchScope ls(DC_K_DECODE_FSM_SYNC, st);
    a.tma_table_bytes = table_bytes;
    const unsigned int grid = (unsigned int)(want < cap ? want : cap);
    if (compat) {
        if (tma) fsm_sync_kernel<true, true><<<grid, threads, smem, st>>>(a, t, fw);
        else fsm_sync_kernel<false, true><<<grid, threads, smem, st>>>(a, t, fw);
    } else {
        if (tma) fsm_sync_kernel<true, false><<<grid, threads, smem, st>>>(a, t, fw);
        else fsm_sync_kernel<false, false><<<grid, threads, smem, st>>>(a, t, fw);
    }
    return cuda_status(cudaGetLastError());
}
static int launch_fsm_write(const uint8_t *d_bits, unsigned long long end, unsigned long long nsubf, unsigned long long nwt,
                            unsigned long long nseg, FastWorkspace fw, uint8_t *d_out, size_t n_out, int32_t *d_status, int mode,
                            int lead, cudaStream_t st) {
    const int rows = fsm_states(mode) + 1 + (fsm_has_entry_rows(fsm_states(mode)) ? kFsmEntryRows : 0);   // + DEAD + the entry rows
    const FsmTables t = fsm_tables_at(fw.fsm);
    const uint32_t stage = fsm_stage_bytes(mode);
    const int warps = 24;
    const size_t smem = kFsmHeaderBytes + (size_t)rows * kFsmWriteRowBytes + kFsmWriteXTableBytes + (size_t)warps * stage + 256;   // (fsm_rows_fit: within the limit)
    const int per_sm = smem <= 56 * 1024 ? 2 : 1;   // (2 x 768 threads: the register file allows no more)
    const bool xtab = !(mode & kModeFsmWide);
    DC_CUDA_TRY(ensure_dynamic_smem(xtab ? (const void *)fsm_write_kernel<true> : (const void *)fsm_write_kernel<false>, smem));
    const unsigned long long want = (nseg + warps - 1) / warps, cap = (unsigned long long)sm_count() * per_sm;
    FsmWriteArgs a;
    a.d_bits = d_bits;
    a.end = end;
    a.nsub = nsubf;
    a.ntiles = nwt;
    a.nseg = nseg;
    a.out = d_out;
    a.n_out = n_out;
    a.lead = lead;
    a.stage_bytes = stage;
    a.rows = (uint32_t)rows;
    a.d_status = d_status;
    LaunchScope ls(DC_K_DECODE_FSM_WRITE, st);
    const unsigned int grid = (unsigned int)(want < cap ? want : cap);
    if (xtab) fsm_write_kernel<true><<<grid, warps * 32, smem, st>>>(a, t, fw);
    else fsm_write_kernel<false><<<grid, warps * 32, smem, st>>>(a, t, fw);
    return cuda_status(cudaGetLastError());
}

// F1 + F2, then F3, over `nwt` warp tiles of the bitstream at d_bits (`end` = end of the STREAM in bits from d_bits; a chunk
// that is not the last one may read a few bytes of the next chunk).  chain == nullptr: a whole stream that starts at
// bit_start.  lead = 1: tile 0 belongs to the previous shard (see F1); nwt and nsubf count it, nseg does not.
static int launch_fast_sync(const uint8_t *d_bits, unsigned long long bit_start, unsigned long long end, unsigned long long nsubf,
                            unsigned long long nwt, unsigned long long nseg, const dc_huff_table *d_table, FastWorkspace fw,
                            size_t n_out, int32_t *d_status, int mode, DecodeChain *chain, int last_chunk, int lead, cudaStream_t st,
                            DecodeChain *host_chain = nullptr, int32_t *host_flag = nullptr) {
    const unsigned long long sms = (unsigned long long)sm_count();
    const unsigned long long want = (nseg + kF_Warps - 1) / kF_Warps;
    if (mode & (kModeFsm | kModeFsmSync)) {
        DC_CUDA_TRY(cudaMemsetAsync(fw.bad_input(), 0, sizeof(int32_t), st));   // set by F1 if the table is not the one the host thinks it is
        const int rc = launch_fsm_sync(d_bits, bit_start, end, nsubf, nwt, nseg, d_table, fw, mode, chain, lead, st);
        if (rc != DC_OK) return rc;
    } else {
    if (mode_t2(write_mode(mode))) DC_CUDA_TRY(cudaMemsetAsync(fw.bad_input(), 0, sizeof(int32_t), st));
    {
        LaunchScope ls(DC_K_DECODE_FAST_SYNC, st);
        const unsigned int g1 = (unsigned int)(want < sms * 8 ? want : sms * 8);
#define DC_F1(M) decode_fast_sync_kernel<M><<<g1, kF_Threads, 0, st>>>(d_bits, bit_start, end, d_table, fw, nsubf, nwt, nseg, chain, lead)
        const int sm = sync_mode(mode);
        if (sm == 0) DC_F1(0); else if (sm == 1) DC_F1(1); else if (sm == 2) DC_F1(2); else if (sm == 3) DC_F1(3); else DC_F1(4);
#undef DC_F1
    }
    }
    {
        ChainSlots slots;
        slots.vals = (unsigned long long *)fw.chain_slots;
        slots.flags = (unsigned int *)(slots.vals + 256);
        DC_CUDA_TRY(cudaMemsetAsync(slots.flags, 0, 256 * sizeof(unsigned int), st));
        const unsigned int g2 = (unsigned int)min((unsigned long long)min(sm_count(), 256), (nseg + 255) / 256);
        LaunchScope ls(DC_K_DECODE_FAST_SCAN, st);
        decode_fast_scan_kernel<<<g2 ? g2 : 1, 1024, 0, st>>>(fw, nseg, n_out, d_status, chain, last_chunk, ((mode & (kModeFsm | kModeFsmSync)) || mode_t2(write_mode(mode))) ? 1 : 0, host_chain, slots, host_flag);
    }
    return cuda_status(cudaGetLastError());
}

static int launch_fast_write(const uint8_t *d_bits, unsigned long long end, unsigned long long nsubf, unsigned long long nwt,
                             unsigned long long nseg, const dc_huff_table *d_table, FastWorkspace fw, uint8_t *d_out, size_t n_out,
                             int32_t *d_status, int mode, uint32_t stage_bytes, int lead, cudaStream_t st) {
    if (mode & kModeFsm) return launch_fsm_write(d_bits, end, nsubf, nwt, nseg, fw, d_out, n_out, d_status, mode, lead, st);
    const unsigned long long sms = (unsigned long long)sm_count();
    const unsigned long long want = (nseg + kF_Warps - 1) / kF_Warps;
    const uint32_t tables_bytes = (uint32_t)fast_tables_bytes(write_mode(mode), lut2_tables(mode));
    mode = write_mode(mode);
    const size_t smem3 = tables_bytes + (size_t)kF_Warps * stage_bytes;
    DC_CUDA_TRY(ensure_write_smem(smem3));
    LaunchScope ls(DC_K_DECODE_FAST_WRITE, st);
    const unsigned int g3 = (unsigned int)(want < sms * 4 ? want : sms * 4);
#define DC_F3(M) decode_fast_write_kernel<M><<<g3, kF_Threads, smem3, st>>>(d_bits, end, d_table, fw, nsubf, nwt, nseg, d_out, n_out, stage_bytes, d_status, lead, tables_bytes)
    if (mode == 0) DC_F3(0); else if (mode == 1) DC_F3(1); else if (mode == 2) DC_F3(2); else DC_F3(3);
#undef DC_F3
    return cuda_status(cudaGetLastError());
}

// test hook: 1 = always take the robust path, 2 = pretend the fast path's guess failed after running it
static int g_decode_force = -1;
static int decode_force_mode() {
    if (g_decode_force < 0) {
        const char *e = getenv("DC_DECODE_FORCE");
        g_decode_force = e ? atoi(e) : 0;
    }
    return g_decode_force;
}
static bool decode_tma() {
    static const bool on = [] { const char *e = getenv("DC_DECODE_TMA"); return e && atoi(e) != 0; }();
    return on;
}
// test hook (not in the public header): 0 = normal, 1 = robust path only, 2 = fast path, then the robust path anyway
extern "C" int dc_debug_decode_mode(int mode) {
    const int old = decode_force_mode();
    g_decode_force = mode;
    return old;
}

extern "C" int dc_huff_decode(const uint8_t *d_bits, uint64_t bit_start, uint64_t nbits, const dc_huff_table *d_table,
                              uint8_t *d_out, size_t n_out, int32_t *d_status, void *d_workspace, size_t workspace_bytes,
                              void *stream) {
    if (!d_table || bit_start >= (uint64_t)kSubBits) return DC_ERR_ARG;
    if (nbits && (!d_bits || !d_workspace || (n_out && !d_out))) return DC_ERR_ARG;
    if ((((uintptr_t)d_bits | (uintptr_t)d_workspace) & 15) != 0) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (nbits == 0) return n_out == 0 ? DC_OK : DC_ERR_CORRUPT;
    size_t off[12];
    unsigned long long nsub, ntiles;
    const size_t need = dec_ws_layout(bit_start, nbits, off, &nsub, &ntiles);
    if (workspace_bytes < need) return DC_ERR_CAPACITY;
    char *w = (char *)d_workspace;
    DecWorkspace ws;
    ws.changed = (int32_t *)w;
    ws.total = (unsigned long long *)(w + 8);
    ws.sub_start = (uint8_t *)(w + off[0]);
    ws.sub_cnt = (uint8_t *)(w + off[1]);
    ws.tile_start = (uint32_t *)(w + off[2]);
    ws.tile_exit = (uint32_t *)(w + off[3]);
    ws.tile_cnt = (uint32_t *)(w + off[4]);
    ws.tile_off = (unsigned long long *)(w + off[5]);
    FastWorkspace fw;
    fw.mismatch = (int32_t *)(w + 16);
    fw.sub_info = (uint16_t *)(w + off[6]);
    fw.seg_cnt = (uint32_t *)(w + off[7]);
    fw.seg_assumed = (uint32_t *)(w + off[8]);
    fw.seg_exit = (uint32_t *)(w + off[9]);
    fw.seg_off = (unsigned long long *)(w + off[10]);
    fw.fsm = w + off[11];
    fw.chain_slots = w + off[11] + kFsmWorkspaceBytes;
    const unsigned long long end = bit_start + nbits;
    const unsigned long long nsubf = (end + kF_SubBits - 1) / kF_SubBits;
    const unsigned long long nwt = (nsubf + 31) / 32, nseg = (nwt + kF_SegTiles - 1) / kF_SegTiles;
    const unsigned long long sms = (unsigned long long)sm_count();

    // the table must be usable before any bit is interpreted
    int32_t tmeta[12];   // the table's first ten words, then lut2_used and fsm_states
    {   // known from the build's mapped words once its event has passed; only a foreign table costs a stream synchronise
        const int rc = table_meta_fetch(d_table, st, tmeta, true);
        if (rc != DC_OK) return rc;
    }
    if (tmeta[8] != DC_OK) return tmeta[8];
    if (tmeta[1] == 0) return DC_ERR_RADIX;

    const int mode = mode_for_start(fast_mode(tmeta), tmeta[1], bit_start);
    if (!decode_mode_usable(tmeta, mode)) return (bit_start % 4) ? DC_ERR_ARG : DC_ERR_RADIX;   // (a nibble code off the nibble grid / too many states)
    const int force = decode_force_mode();
    const bool nib = nibble_radix(tmeta[0]);
    auto serial = [&]() -> int {   // the nibble radices' last resort (their tables have no window LUTs for the robust path)
        launch_fsm_build(d_table, fw, mode, st);
        LaunchScope ls(DC_K_DECODE_SYNC, st);
        fsm_serial_kernel<<<1, 32, 0, st>>>(d_bits, bit_start, end, fsm_tables_at(fw.fsm), d_out, n_out, d_status);
        return cuda_status(cudaGetLastError());
    };
    if (force == 1) return nib ? serial() : decode_robust(d_bits, bit_start, end, d_table, d_out, n_out, d_status, ws, nsub, ntiles, st);

    const uint32_t stage_bytes = fast_stage_bytes(tmeta);
    // did every segment start on a code boundary?  F2 knows and stores the answer in mapped host memory as well; the host waits
    // for F2's event only, so F3 (which returns at once on a mismatch) is already queued and the caller's next launches
    // follow it without a gap.
    int32_t mismatch = 0;
    {
        HostFlag hf;
        const int have_flag = host_flag_acquire(&hf) == DC_OK;
        if (have_flag) *hf.host = -1;
        int rc = launch_fast_sync(d_bits, bit_start, end, nsubf, nwt, nseg, d_table, fw, n_out, d_status, mode, nullptr, 1, 0, st, nullptr,
                                  have_flag ? hf.dev : nullptr);
        if (rc == DC_OK && have_flag && cudaEventRecord(hf.ev, st) != cudaSuccess) rc = DC_ERR_CUDA;
        if (rc == DC_OK) rc = launch_fast_write(d_bits, end, nsubf, nwt, nseg, d_table, fw, d_out, n_out, d_status, mode, stage_bytes, 0, st);
        if (rc == DC_OK && have_flag) {
            if (cudaEventSynchronize(hf.ev) != cudaSuccess) rc = DC_ERR_CUDA;
            else mismatch = *hf.host;
        }
        if (have_flag) host_flag_release(hf);
        if (rc != DC_OK) return rc;
        if (!have_flag || mismatch < 0) {   // no mapped flag to be had: the blocking read
            DC_CUDA_TRY(cudaMemcpyAsync(&mismatch, fw.mismatch, sizeof mismatch, cudaMemcpyDeviceToHost, st));
            DC_CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    if (mismatch || force == 2) {
        if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
        if (nib) return serial();
        return decode_robust(d_bits, bit_start, end, d_table, d_out, n_out, d_status, ws, nsub, ntiles, st);
    }
    return DC_OK;
}

// ================================================================================================ a stream with its index (opt-in)
// The index is what F1 + F2 leave behind -- sub_info (2 bytes per 256-bit subsequence) and seg_off (8 bytes per 16 KB segment) --
// kept by the producer, so that the consumer runs F3 only.  Both calls point the workspace's two arrays INTO the index buffer:
// nothing is copied.
namespace {
constexpr uint64_t kIndexMagic = 0x3149444358464344ull;   // "DCFXIDI1"
struct IndexGeom {
    unsigned long long end, nsubf, nwt, nseg;
    size_t sub_info_off, seg_off_off, total;
};
IndexGeom index_geom(uint64_t bit_start, uint64_t nbits) {
    IndexGeom g;
    g.end = bit_start + nbits;
    g.nsubf = (g.end + kF_SubBits - 1) / kF_SubBits;
    g.nwt = (g.nsubf + 31) / 32;
    g.nseg = (g.nwt + kF_SegTiles - 1) / kF_SegTiles;
    g.sub_info_off = 0;
    g.seg_off_off = ((size_t)g.nwt * 32 * 2 + 63) & ~(size_t)63;
    g.total = g.seg_off_off + (size_t)g.nseg * 8;
    return g;
}
}  // namespace

extern "C" size_t dc_huff_index_bytes(uint64_t bit_start, uint64_t nbits) { return nbits ? index_geom(bit_start, nbits).total : 0; }

static int index_workspace(uint64_t bit_start, uint64_t nbits, void *d_workspace, size_t workspace_bytes, void *d_index, FastWorkspace *fw) {
    size_t off[12];
    unsigned long long nsub, ntiles;
    if (workspace_bytes < dec_ws_layout(bit_start, nbits, off, &nsub, &ntiles)) return DC_ERR_CAPACITY;
    const IndexGeom g = index_geom(bit_start, nbits);
    char *w = (char *)d_workspace, *x = (char *)d_index;
    fw->mismatch = (int32_t *)(w + 16);
    fw->sub_info = (uint16_t *)(x + g.sub_info_off);
    fw->seg_cnt = (uint32_t *)(w + off[7]);
    fw->seg_assumed = (uint32_t *)(w + off[8]);
    fw->seg_exit = (uint32_t *)(w + off[9]);
    fw->seg_off = (unsigned long long *)(x + g.seg_off_off);
    fw->fsm = w + off[11];
    fw->chain_slots = w + off[11] + kFsmWorkspaceBytes;
    return DC_OK;
}

extern "C" int dc_huff_index_build(const uint8_t *d_bits, uint64_t bit_start, uint64_t nbits, const dc_huff_table *d_table, uint64_t n_symbols,
                                   void *d_index, size_t index_bytes, dc_huff_index_info *info, void *d_workspace, size_t workspace_bytes,
                                   void *stream) {
    if (!d_bits || !d_table || !d_index || !info || !d_workspace || nbits == 0 || bit_start >= (uint64_t)kSubBits) return DC_ERR_ARG;
    if ((((uintptr_t)d_bits | (uintptr_t)d_workspace | (uintptr_t)d_index) & 15) != 0) return DC_ERR_ARG;
    const IndexGeom g = index_geom(bit_start, nbits);
    if (index_bytes < g.total) return DC_ERR_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    FastWorkspace fw;
    int rc = index_workspace(bit_start, nbits, d_workspace, workspace_bytes, d_index, &fw);
    if (rc != DC_OK) return rc;
    int32_t tmeta[12];
    rc = table_meta_fetch(d_table, st, tmeta, true);
    if (rc != DC_OK) return rc;
    if (tmeta[8] != DC_OK) return tmeta[8];
    if (tmeta[1] == 0) return DC_ERR_RADIX;
    const int mode = mode_for_start(fast_mode(tmeta), tmeta[1], bit_start);
    if (!decode_mode_usable(tmeta, mode)) return DC_ERR_RADIX;
    int32_t *d_st = (int32_t *)d_workspace;   // (the robust path's scratch word: not used here)
    DC_CUDA_TRY(cudaMemsetAsync(d_st, 0, sizeof(int32_t), st));
    rc = launch_fast_sync(d_bits, bit_start, g.end, g.nsubf, g.nwt, g.nseg, d_table, fw, (size_t)n_symbols, d_st, mode, nullptr, 1, 0, st);
    if (rc != DC_OK) return rc;
    int32_t words[4] = {0, 0, 0, 0};   // mismatch, (mode), bad_input, start token
    int32_t status = 0;
    DC_CUDA_TRY(cudaMemcpyAsync(words, fw.mismatch, sizeof words, cudaMemcpyDeviceToHost, st));
    DC_CUDA_TRY(cudaMemcpyAsync(&status, d_st, sizeof status, cudaMemcpyDeviceToHost, st));
    DC_CUDA_TRY(cudaStreamSynchronize(st));
    if (words[0]) return 1;              // some segment started on a wrong guess: this stream has no index
    if (status != DC_OK) return status;  // the symbol total does not match n_symbols
    info->magic = kIndexMagic;
    info->bit_start = bit_start;
    info->nbits = nbits;
    info->n_symbols = n_symbols;
    info->mode = (uint32_t)mode;
    info->start_token = (uint32_t)words[3];
    info->reserved[0] = info->reserved[1] = 0;
    return DC_OK;
}

extern "C" int dc_huff_decode_indexed(const uint8_t *d_bits, const dc_huff_index_info *info, const dc_huff_table *d_table, const void *d_index,
                                      size_t index_bytes, uint8_t *d_out, size_t n_out, int32_t *d_status, void *d_workspace,
                                      size_t workspace_bytes, void *stream) {
    if (!d_bits || !info || !d_table || !d_index || !d_workspace || (n_out && !d_out)) return DC_ERR_ARG;
    if (info->magic != kIndexMagic || info->nbits == 0 || info->bit_start >= (uint64_t)kSubBits || info->n_symbols != (uint64_t)n_out) return DC_ERR_ARG;
    if ((((uintptr_t)d_bits | (uintptr_t)d_workspace | (uintptr_t)d_index) & 15) != 0) return DC_ERR_ARG;
    const IndexGeom g = index_geom(info->bit_start, info->nbits);
    if (index_bytes < g.total) return DC_ERR_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    FastWorkspace fw;
    int rc = index_workspace(info->bit_start, info->nbits, d_workspace, workspace_bytes, (void *)d_index, &fw);
    if (rc != DC_OK) return rc;
    int32_t tmeta[12];
    rc = table_meta_fetch(d_table, st, tmeta, true);
    if (rc != DC_OK) return rc;
    if (tmeta[8] != DC_OK) return tmeta[8];
    if (tmeta[1] == 0) return DC_ERR_RADIX;
    const int mode = mode_for_start(fast_mode(tmeta), tmeta[1], info->bit_start);
    if ((uint32_t)mode != info->mode || !decode_mode_usable(tmeta, mode)) return DC_ERR_ARG;   // another table (or another build of the library) made this index
    const int32_t words[4] = {0, mode, 0, (int32_t)info->start_token};
    DC_CUDA_TRY(cudaMemcpyAsync(fw.mismatch, words, sizeof words, cudaMemcpyHostToDevice, st));
    if (mode & kModeFsm) launch_fsm_build(d_table, fw, mode, st);
    return launch_fast_write(d_bits, g.end, g.nsubf, g.nwt, g.nseg, d_table, fw, d_out, n_out, d_status, mode, fast_stage_bytes(tmeta), 0, st);
}

// ================================================================================================ shards of a longer stream

namespace {
struct ShardGeom {
    const uint8_t *base;              // d_bits, or d_bits - 1024 with a halo
    unsigned long long end, nsubf, nwt, nseg;
    int lead;
    FastWorkspace fw;
    DecodeChain *chain;
    int mode, bpd;
    bool nib;                         // nibble-per-digit radix: the state machine or nothing
    uint32_t stage_bytes;
};
}  // namespace

static int shard_geometry(const uint8_t *d_bits, int has_halo, uint64_t shard_bits, uint64_t stream_bits_left,
                          const dc_huff_table *d_table, void *d_workspace, size_t workspace_bytes, cudaStream_t st, ShardGeom *g) {
    if (!d_bits || !d_table || !d_workspace || shard_bits == 0 || stream_bits_left < shard_bits) return DC_ERR_ARG;
    if ((((uintptr_t)d_bits | (uintptr_t)d_workspace) & 15) != 0) return DC_ERR_ARG;
    const unsigned long long tile_bits = 32ull * kF_SubBits;
    if (stream_bits_left > shard_bits && shard_bits % tile_bits != 0) return DC_ERR_ARG;  // only the last shard is ragged
    g->lead = has_halo ? 1 : 0;
    g->base = d_bits - (size_t)g->lead * (tile_bits / 8);
    g->end = stream_bits_left + (unsigned long long)g->lead * tile_bits;          // stream end, in bits from g->base
    const unsigned long long own = shard_bits + (unsigned long long)g->lead * tile_bits;
    g->nsubf = (own + kF_SubBits - 1) / kF_SubBits;                               // subsequences incl. the lead tile
    g->nwt = (g->nsubf + 31) / 32;
    g->nseg = (g->nwt - g->lead + kF_SegTiles - 1) / kF_SegTiles;
    size_t off[12];
    unsigned long long nsub, ntiles;
    if (workspace_bytes < dec_ws_layout(0, own, off, &nsub, &ntiles)) return DC_ERR_CAPACITY;
    char *w = (char *)d_workspace;
    g->fw.mismatch = (int32_t *)(w + 16);
    g->fw.sub_info = (uint16_t *)(w + off[6]);
    g->fw.seg_cnt = (uint32_t *)(w + off[7]);
    g->fw.seg_assumed = (uint32_t *)(w + off[8]);
    g->fw.seg_exit = (uint32_t *)(w + off[9]);
    g->fw.seg_off = (unsigned long long *)(w + off[10]);
    g->fw.fsm = w + off[11];
    g->fw.chain_slots = w + off[11] + kFsmWorkspaceBytes;

    g->chain = (DecodeChain *)(w + 32);
    int32_t tmeta[12];   // the table's first ten words, then lut2_used and fsm_states
    {   // known from the build's mapped words once its event has passed; only a foreign table costs a stream synchronise
        const int rc = table_meta_fetch(d_table, st, tmeta, true);
        if (rc != DC_OK) return rc;
    }
    if (tmeta[8] != DC_OK) return tmeta[8];
    if (tmeta[1] == 0) return DC_ERR_RADIX;
    g->mode = fast_mode(tmeta);
    if (!decode_mode_usable(tmeta, g->mode)) return DC_ERR_RADIX;
    g->bpd = tmeta[1];
    g->nib = nibble_radix(tmeta[0]);
    g->stage_bytes = fast_stage_bytes(tmeta);
    return DC_OK;
}

extern "C" int dc_huff_decode_shard_sync(const uint8_t *d_bits, int has_halo, unsigned first_code_bit, uint64_t shard_bits,
                                         uint64_t stream_bits_left, const dc_huff_table *d_table, dc_shard_summary *d_summary,
                                         void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!d_summary || (!has_halo && !(first_code_bit & kFsmToken) && first_code_bit >= (unsigned)kSubBits)) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    ShardGeom g;
    int rc = shard_geometry(d_bits, has_halo, shard_bits, stream_bits_left, d_table, d_workspace, workspace_bytes, st, &g);
    if (rc != DC_OK) return rc;
    if (!has_halo) {
        if ((first_code_bit & kFsmToken) && !(g.mode & (kModeFsm | kModeFsmSync))) return DC_ERR_ARG;   // a state token, but the table has no state machine
        g.mode = mode_for_start(g.mode, g.bpd, first_code_bit);
        if (g.nib && !(g.mode & kModeFsm)) return DC_ERR_ARG;   // a nibble code cannot start off the nibble grid
    }
    DC_CUDA_TRY(cudaMemcpyAsync(g.fw.mismatch + 1, &g.mode, sizeof(int), cudaMemcpyHostToDevice, st));   // the write phase takes the same kernels
    DecodeChain init = {0ull, has_halo ? 0u : first_code_bit, 0, 0u, 0u};
    DC_CUDA_TRY(cudaMemcpyAsync(g.chain, &init, sizeof init, cudaMemcpyHostToDevice, st));
    // (F2 never checks the symbol total here: last_chunk = 0; the caller checks the sum over all shards)
    rc = launch_fast_sync(g.base, first_code_bit, g.end, g.nsubf, g.nwt, g.nseg, d_table, g.fw, 0, nullptr, g.mode, g.chain, 0, g.lead, st);
    if (rc != DC_OK) return rc;
    DC_CUDA_TRY(cudaMemcpyAsync(d_summary, g.chain, sizeof(DecodeChain), cudaMemcpyDeviceToDevice, st));
    return DC_OK;
}

extern "C" int dc_huff_decode_shard_write(const uint8_t *d_bits, int has_halo, uint64_t shard_bits, uint64_t stream_bits_left,
                                          const dc_huff_table *d_table, uint8_t *d_out, size_t n_out, int32_t *d_status,
                                          void *d_workspace, size_t workspace_bytes, void *stream) {
    if (n_out && !d_out) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    ShardGeom g;
    int rc = shard_geometry(d_bits, has_halo, shard_bits, stream_bits_left, d_table, d_workspace, workspace_bytes, st, &g);
    if (rc != DC_OK) return rc;
    DC_CUDA_TRY(cudaMemcpyAsync(&g.mode, g.fw.mismatch + 1, sizeof(int), cudaMemcpyDeviceToHost, st));   // what the sync phase decided
    DC_CUDA_TRY(cudaStreamSynchronize(st));
    return launch_fast_write(g.base, g.end, g.nsubf, g.nwt, g.nseg, d_table, g.fw, d_out, n_out, d_status, g.mode, g.stage_bytes, g.lead, st);
}

// ================================================================================================ pipelined host decompress
// dc_host_huff_decompress on a large stream: the payload goes up in 32 MiB chunks, each chunk is decoded as soon as the
// chunk after it has arrived (its last code may cross into it) and its symbols go back down while the next chunks are
// still coming up, so the two PCIe directions overlap.  Chunks are chained on the device (DecodeChain: first code's bit
// offset, output offset); the host reads the 16-byte chain state after every chunk's scan only to learn how many bytes
// to copy back.  Returns DC_OK, a negative dc_status, or +1 = "not taken, use the one-shot path" (small streams, test
// hooks, or a stream that did not self-synchronise -- everything is then on the device already).
namespace dc {

constexpr unsigned long long kPipeChunkTiles = 32768;  // 32 MiB of bitstream; a multiple of kF_SegTiles

struct PipeState {
    cudaStream_t up = nullptr, cp = nullptr, dn = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_f2, ev_f3;
    DecodeChain *h_ring = nullptr, *d_ring = nullptr;   // mapped pinned memory: F2 stores the chain state of chunk k into slot k
    size_t ring = 0;
    int ensure(size_t k) {
        if (!up) {
            DC_CUDA_TRY(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
            DC_CUDA_TRY(cudaStreamCreateWithFlags(&cp, cudaStreamNonBlocking));
            DC_CUDA_TRY(cudaStreamCreateWithFlags(&dn, cudaStreamNonBlocking));
        }
        while (ev_up.size() < k) {
            cudaEvent_t a, b, c;
            DC_CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            DC_CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            DC_CUDA_TRY(cudaEventCreateWithFlags(&c, cudaEventDisableTiming));
            ev_up.push_back(a); ev_f2.push_back(b); ev_f3.push_back(c);
        }
        if (ring < k) {
            if (h_ring) cudaFreeHost(h_ring);
            h_ring = nullptr; ring = 0;
            DC_CUDA_TRY(cudaHostAlloc((void **)&h_ring, k * sizeof(DecodeChain), cudaHostAllocMapped));
            DC_CUDA_TRY(cudaHostGetDevicePointer((void **)&d_ring, h_ring, 0));
            ring = k;
        }
        return DC_OK;
    }
};
static PipeState g_pipe;

static int host_decompress_pipelined_body(const uint8_t *h_payload, uint64_t total_bits, const dc_huff_table *d_table, uint8_t *d_bits,
                                          uint8_t *d_out, void *d_workspace, size_t workspace_bytes, uint8_t *h_out, size_t n_out,
                                          int32_t *d_status, uint8_t *d_packed) {
    const bool trits = d_packed != nullptr;
    const size_t nbytes = (size_t)((total_bits + 7) / 8);
    static unsigned long long chunk_tiles = 0;
    if (chunk_tiles == 0) {  // DC_PIPE_CHUNK_MIB: tuning knob (a multiple of the 16 KB segment either way)
        const char *e = getenv("DC_PIPE_CHUNK_MIB");
        const unsigned long long mib = e ? strtoull(e, nullptr, 10) : 0;
        chunk_tiles = mib >= 1 && mib <= 1024 ? mib * 1024 : kPipeChunkTiles;
    }
    unsigned long long chunk_bytes = chunk_tiles * (kF_TileVecs * 16ull);
    // radix 3: a chunk of the 2-bit-per-trit stream has to be whole segments (16 KB) AND whole 16-byte groups of the
    // payload (20 bytes of stream each): multiples of 81 920 bytes
    if (trits) chunk_bytes = chunk_bytes / 81920 * 81920;
    if (chunk_bytes == 0) return 1;
    const size_t nchunk = (size_t)((nbytes + chunk_bytes - 1) / chunk_bytes);
    if (nchunk < 3 || decode_force_mode() != 0) return 1;
    size_t off[12];
    unsigned long long nsub, ntiles;
    if (workspace_bytes < dec_ws_layout(0, total_bits, off, &nsub, &ntiles)) return DC_ERR_CAPACITY;
    if (g_pipe.ensure(nchunk) != DC_OK) return DC_ERR_CUDA;
    char *w = (char *)d_workspace;
    FastWorkspace fw;
    fw.mismatch = (int32_t *)(w + 16);
    fw.sub_info = (uint16_t *)(w + off[6]);
    fw.seg_cnt = (uint32_t *)(w + off[7]);
    fw.seg_assumed = (uint32_t *)(w + off[8]);
    fw.seg_exit = (uint32_t *)(w + off[9]);
    fw.seg_off = (unsigned long long *)(w + off[10]);
    fw.fsm = w + off[11];
    fw.chain_slots = w + off[11] + kFsmWorkspaceBytes;
    DecodeChain *d_chain = (DecodeChain *)(w + 32);

    // the table (built on the legacy stream by the caller) must be usable before any bit is interpreted
    int32_t tmeta[12];   // the table's first ten words, then lut2_used and fsm_states
    {
        const int rc = table_meta_fetch(d_table, 0, tmeta, true);
        if (rc != DC_OK) return rc;
    }
    if (tmeta[8] != DC_OK) return tmeta[8];
    if (tmeta[1] == 0) return DC_ERR_RADIX;
    const int mode = fast_mode(tmeta);
    if (!decode_mode_usable(tmeta, mode)) return 1;   // (not taken: the one-shot path reports it)
    const uint32_t stage_bytes = fast_stage_bytes(tmeta);

    cudaStream_t up = g_pipe.up, cp = g_pipe.cp, dn = g_pipe.dn;
    for (size_t k = 0; k < nchunk; k++) {
        // 64 bytes more than the chunk: its last code, and the look-ahead word of its last lane, reach into the next chunk
        // (which uploads the same bytes again); so chunk k can be decoded as soon as upload k is in
        const size_t o = (size_t)(k * chunk_bytes), len = (size_t)min((unsigned long long)(nbytes - o), chunk_bytes + 64);
        if (trits) {  // 4 trits per stream byte, 5 per payload byte
            const size_t po = o / 5 * 4, pbytes = (size_t)((total_bits / 2 + 4) / 5);
            const size_t plen = min(pbytes - po, (size_t)(chunk_bytes / 5 * 4) + 64);
            DC_CUDA_TRY(cudaMemcpyAsync(d_packed + po, h_payload + po, plen, cudaMemcpyHostToDevice, up));
        } else {
            DC_CUDA_TRY(cudaMemcpyAsync(d_bits + o, h_payload + o, len, cudaMemcpyHostToDevice, up));
        }
        DC_CUDA_TRY(cudaEventRecord(g_pipe.ev_up[k], up));
    }
    DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), cp));
    DC_CUDA_TRY(cudaMemsetAsync(d_chain, 0, sizeof(DecodeChain), cp));  // base 0, first code at bit 0, no mismatch
    unsigned long long base = 0;
    bool fallback = false;
    for (size_t k = 0; k < nchunk; k++) {
        const unsigned long long bit0 = k * chunk_bytes * 8ull;
        const unsigned long long end_rel = total_bits - bit0;                    // stream end, from this chunk's first bit
        const unsigned long long bits_here = min(end_rel, chunk_bytes * 8ull);   // bits that belong to this chunk
        const unsigned long long nsubf = (bits_here + kF_SubBits - 1) / kF_SubBits;
        const unsigned long long nwt = (nsubf + 31) / 32, nseg = (nwt + kF_SegTiles - 1) / kF_SegTiles;
        const int last = k + 1 == nchunk;
        DC_CUDA_TRY(cudaStreamWaitEvent(cp, g_pipe.ev_up[k], 0));
        if (trits) {  // this chunk's trits, and 256 of the next chunk's (the 64 bytes of look-ahead), into the stream buffer
            const unsigned long long t0 = k * chunk_bytes * 4ull, ntrits = total_bits / 2;
            const unsigned long long cnt = min(ntrits - t0, chunk_bytes * 4ull + 256ull);
            const int rc = trit_unpack_launch(d_packed + k * chunk_bytes / 5 * 4, cnt, d_bits + k * chunk_bytes, d_status, cp);
            if (rc != DC_OK) return rc;
        }
        // F1 and F2, the chain state for the host, then F3 (launch_fast's order, with the read-back in between)
        const uint8_t *bits_k = d_bits + k * chunk_bytes;
        {
            // F2 also stores the chain state into the mapped ring: a cudaMemcpyAsync of these 24 bytes would queue on the
            // device-to-host copy engine behind the previous chunk's 44 MB of output and lock-step the pipeline with it
            const int rc = launch_fast_sync(bits_k, 0, end_rel, nsubf, nwt, nseg, d_table, fw, n_out, d_status, mode, d_chain, last, 0, cp,
                                            g_pipe.d_ring + k);
            if (rc != DC_OK) return rc;
        }
        DC_CUDA_TRY(cudaEventRecord(g_pipe.ev_f2[k], cp));
        {
            const int rc = launch_fast_write(bits_k, end_rel, nsubf, nwt, nseg, d_table, fw, d_out, n_out, d_status, mode, stage_bytes, 0, cp);
            if (rc != DC_OK) return rc;
        }
        DC_CUDA_TRY(cudaEventRecord(g_pipe.ev_f3[k], cp));
        DC_CUDA_TRY(cudaGetLastError());
        DC_CUDA_TRY(cudaEventSynchronize(g_pipe.ev_f2[k]));  // how many symbols did this chunk hold?
        const DecodeChain c = g_pipe.h_ring[k];
        if (c.mismatch) { fallback = true; break; }
        const unsigned long long hi = min((unsigned long long)n_out, c.base);
        if (hi > base) {
            DC_CUDA_TRY(cudaStreamWaitEvent(dn, g_pipe.ev_f3[k], 0));
            DC_CUDA_TRY(cudaMemcpyAsync(h_out + base, d_out + base, (size_t)(hi - base), cudaMemcpyDeviceToHost, dn));
        }
        base = c.base > base ? c.base : base;
    }
    DC_CUDA_TRY(cudaStreamSynchronize(up));
    DC_CUDA_TRY(cudaStreamSynchronize(cp));
    DC_CUDA_TRY(cudaStreamSynchronize(dn));
    if (fallback) return 1;
    int32_t stt = 0;
    DC_CUDA_TRY(cudaMemcpy(&stt, d_status, sizeof stt, cudaMemcpyDeviceToHost));
    return stt;
}

// Every exit of the body -- the error exits too -- passes through here: the three non-blocking streams may still have
// uploads and kernels queued against arena memory that the next dc_host_* call reuses or frees on the legacy stream, which
// does not order against them.
int host_decompress_pipelined(const uint8_t *h_payload, uint64_t total_bits, const dc_huff_table *d_table, uint8_t *d_bits,
                              uint8_t *d_out, void *d_workspace, size_t workspace_bytes, uint8_t *h_out, size_t n_out,
                              int32_t *d_status, uint8_t *d_packed) {
    const int rc = host_decompress_pipelined_body(h_payload, total_bits, d_table, d_bits, d_out, d_workspace, workspace_bytes, h_out, n_out,
                                                  d_status, d_packed);
    if (g_pipe.up) {
        cudaStreamSynchronize(g_pipe.up);
        cudaStreamSynchronize(g_pipe.cp);
        cudaStreamSynchronize(g_pipe.dn);
        cudaGetLastError();
    }
    return rc;
}

}  // namespace dc
