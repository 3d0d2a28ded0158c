// k6_nybble_text.cu -- K6: the static-table nybble compressor / decompressor.
//
// Replaces compress_bytestring(src, dst, false) (nybble_compression.c:887-1038, with compress_byte_index :819-884)
// and decompress_bytestring(src, dst, false) (:734-817, with decompress_nybble :643-663).  With modify == false the
// 16 contexts all hold " etaoins" (initialize_dictionary :546-562) and never change, so the coder is a two-state
// transducer over the input and both directions are exact parallel scans:
//
//   compress    state q = parity of the run of table hits that ends just before the byte.
//               hit : q == 1 -> emit the pair byte ((8|i_prev) << 4) | (8|i), q = 0;   q == 0 -> q = 1
//               miss: q == 1 -> emit the previous byte as a literal (the reference rewrites its half-written
//                               byte, :855-857), then this byte;  q == 0 -> emit this byte;   q = 0
//               end : q == 1 -> emit the last byte as a literal (:1000-1009)
//               stream = 0xAF, src[0], body, NUL (:903-905); if that is not shorter than the source: ' ' + source (:1018-1037)
//   decompress  units are nibbles, high first (:767-773); state q = 1 inside a literal.
//               nibble with bit 3: q == 0 -> letter[nibble & 7], q == 1 -> literal ((prev & 7) << 4) + nibble;   q = 0
//               nibble without  : q == 0 -> q = 1 (first half of a literal), q == 1 -> literal as above, q = 0
//               end : q == 1 -> the dangling half literal is completed with a zero nibble
//               type byte 0xAF as above, ' ' = copy the rest (:799-805), anything else = copy everything (:806-812)
//
// A block of units is summarised by its transfer function over the two states: symbols emitted and next state, for
// each entry state.  Functions compose associatively, so: S1 one function per 8 KB tile, S2 a one-CTA scan of the
// tile functions (exclusive output offset and entry state of every tile, total length, the fall-back decision),
// S3 every tile re-derives its threads' entry states, emits into shared memory and copies out with aligned
// 16-byte stores.  HBM traffic 2N + C.
#include "dc_common.cuh"

namespace dc {

constexpr int kTxThreads = 256;
constexpr int kTxVec = 2;                      // 16-byte vectors per thread
constexpr int kTxPer = 16 * kTxVec;            // input bytes per thread
constexpr int kTxTile = kTxThreads * kTxPer;   // 8 KB per CTA step
constexpr int kTxWarps = kTxThreads / 32;

// ---- transfer functions, packed: e0 [0,15) | e1 [15,30) | next state from 0 (bit 30) | next state from 1 (bit 31)
__device__ __forceinline__ uint32_t fn_pack(uint32_t e0, uint32_t e1, uint32_t q0, uint32_t q1) {
    return e0 | (e1 << 15) | (q0 << 30) | (q1 << 31);
}
__device__ __forceinline__ uint32_t fn_emit(uint32_t f, uint32_t q) { return (f >> (15 * q)) & 0x7FFFu; }
__device__ __forceinline__ uint32_t fn_next(uint32_t f, uint32_t q) { return (f >> (30 + q)) & 1u; }
__device__ __forceinline__ uint32_t fn_compose(uint32_t a, uint32_t b) {  // a, then b
    const uint32_t a0 = fn_next(a, 0), a1 = fn_next(a, 1);
    return fn_pack(fn_emit(a, 0) + fn_emit(b, a0), fn_emit(a, 1) + fn_emit(b, a1), fn_next(b, a0), fn_next(b, a1));
}
constexpr uint32_t kFnIdentity = 1u << 31;  // emits nothing, keeps the state

// " etaoins" (initialize_dictionary :552-559).  letter_of: byte i of the eight letters (one PRMT).
// The inverse is a 128-entry table in shared memory (8 = not a letter), filled by fill_letter_table().
__device__ __forceinline__ uint32_t letter_of(uint32_t i) { return __byte_perm(0x61746520u, 0x736e696fu, i & 7u) & 0xFFu; }
__device__ __forceinline__ void fill_letter_table(uint8_t *lut) {
    for (int c = threadIdx.x; c < 128; c += blockDim.x) {
        uint32_t idx = 8;
#pragma unroll
        for (uint32_t i = 0; i < 8; i++)
            if (letter_of(i) == (uint32_t)c) idx = i;
        lut[c] = (uint8_t)idx;
    }
}

struct TxWorkspace {
    uint32_t *tile_fn;                 // [ntiles]  S1
    unsigned long long *tile_off;      // [ntiles]  S2: symbols emitted by the tiles before this one
    uint8_t *tile_q;                   // [ntiles]  S2: entry state
    unsigned long long *out_len;       // [1] bytes of the output (without the NUL)
    int32_t *mode;                     // [1] compress: 1 = fall back to ' ' + raw; decompress: 0 = 0xAF, 1 = ' ', 2 = other; -1 = do nothing
};

// ---- a thread's block as bit masks.  Units are bytes (compress, 32 per block) or nibbles in stream order
// (decompress, 64 per block); bit j of a mask = unit j.
//   T  units that toggle the state: table hits (compress) / nibbles without bit 3 (decompress)
//   V  units that take part: they exist and are not the stream's header (byte 0 when compressing, bytes 0-1 when
//      decompressing); a unit outside T resets the state to 0
// The state before every unit follows from T alone: inside a run of T units it alternates, starting from 0 after a
// reset (or from the block's entry state for the run that begins at unit 0).  E = units at an even offset inside
// their run (after such a unit the state is 1); the usual carry trick finds it without a loop.
template <typename M> struct TxBlock {
    M T, V;
    int units;       // units present in the block (header included)
};

template <typename M> __device__ __forceinline__ M even_mask();
template <> __device__ __forceinline__ uint32_t even_mask<uint32_t>() { return 0x55555555u; }
template <> __device__ __forceinline__ unsigned long long even_mask<unsigned long long>() { return 0x5555555555555555ull; }
__device__ __forceinline__ int popcnt(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int popcnt(unsigned long long x) { return __popcll(x); }

// E for entry state 0, and the leading run F (the only part the entry state changes)
template <typename M>
__device__ __forceinline__ void run_parity(M T, M &E, M &F) {
    const M even = even_mask<M>();
    const M S = T & ~(T << 1);                       // run starts
    const M Re = ((T + (S & even)) ^ T) & T;         // runs that start on an even unit
    E = (Re & even) | (T & ~Re & ~even);
    F = T & ~(T + 1);                                // the run that begins at unit 0
}
template <typename M> __device__ __forceinline__ M low_bits(int k) { return k >= (int)(8 * sizeof(M)) ? ~(M)0 : (((M)1 << k) - 1); }

// the two 16-byte vectors of the block (0 beyond n), unpacked; prev = the byte in front of the block (0 at the start)
// byte k of the block (k is a compile-time constant after unrolling: one shift and one mask, no array of bytes kept live)
__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[kTxPer / 4], int k) { return (w[k >> 2] >> (8 * (k & 3))) & 0xFFu; }

__device__ __forceinline__ void load_block(const uint8_t *__restrict__ in, size_t n, size_t base, uint32_t &prev, uint32_t (&w)[kTxPer / 4]) {
    if (base + kTxPer <= n && ((uintptr_t)(in + base) & 15) == 0) {
        uint4 v[kTxVec];
#pragma unroll
        for (int j = 0; j < kTxVec; j++) v[j] = ldg_stream((const uint4 *)(in + base) + j);
#pragma unroll
        for (int j = 0; j < kTxVec; j++) { w[4 * j] = v[j].x; w[4 * j + 1] = v[j].y; w[4 * j + 2] = v[j].z; w[4 * j + 3] = v[j].w; }
    } else {
#pragma unroll
        for (int j = 0; j < kTxPer / 4; j++) {
            uint32_t x = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) x |= (base + 4 * j + k < n ? (uint32_t)in[base + 4 * j + k] : 0u) << (8 * k);
            w[j] = x;
        }
    }
    prev = base > 0 && base - 1 < n ? in[base - 1] : 0u;
}

// masks of a compress block: T = table hits.  li[k] = table index of byte k (8 = not in the table).  ADAPT: the indices
// are the move-to-front positions K8 computed (pos[i] for byte i of the string), not the static table's.
template <bool ADAPT>
__device__ __forceinline__ void compress_masks(const uint8_t *lut, const uint8_t *__restrict__ pos, const uint32_t (&w)[kTxPer / 4], size_t base,
                                               size_t n, TxBlock<uint32_t> &b, uint32_t (&li)[kTxPer], bool &bad) {
    const int valid = (int)min((size_t)kTxPer, n - base), skip = base == 0 ? 1 : 0;
    b.units = valid;
    b.V = low_bits<uint32_t>(valid) & ~low_bits<uint32_t>(skip);
    uint32_t miss = 0;
    uint32_t pw[kTxPer / 4];
    if (ADAPT) {
        uint32_t unused;
        load_block(pos, n, base, unused, pw);
    }
#pragma unroll
    for (int k = 0; k < kTxPer; k++) {
        li[k] = ADAPT ? (k < valid ? byte_of(pw, k) & 0xFu : 8u) : lut[byte_of(w, k) & 0x7Fu];
        miss |= k >= 3 ? (li[k] & 8u) << (k - 3) : (li[k] & 8u) >> (3 - k);
    }
    b.T = ~miss & b.V;
    // assert( source[i] < 0x80 ) :910, and 0x00 would end the C string: a zero byte or a byte with bit 7 among the valid ones
    uint32_t flag = 0;
#pragma unroll
    for (int j = 0; j < kTxPer / 4; j++) {
        const uint32_t x = w[j];
        uint32_t f = (x | ((x - 0x01010101u) & ~x)) & 0x80808080u;   // bit 7 of every byte that is 0 or >= 0x80
        const int nb = valid - 4 * j;                                  // bytes of this word that exist
        if (nb < 4) f &= nb <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - nb)));
        if (j == 0 && skip) f &= ~0xFFu;                               // the verbatim first byte is not looked at
        flag |= f;
    }
    bad |= flag != 0;
}

// masks of a decompress block: T = nibbles WITHOUT bit 3 (first or second halves of literals), in stream order
__device__ __forceinline__ void decompress_masks(const uint32_t (&w)[kTxPer / 4], size_t base, size_t n, TxBlock<unsigned long long> &b) {
    const int valid = (int)min((size_t)kTxPer, n - base), skip = base == 0 ? 2 : 0;
    b.units = 2 * valid;
    b.V = low_bits<unsigned long long>(2 * valid) & ~low_bits<unsigned long long>(2 * min(skip, valid));
    unsigned long long hmask = 0;
#pragma unroll
    for (int j = 0; j < kTxPer / 4; j++) {
        // bit 3 of the 8 nibbles of word j -> 8 consecutive bits, low nibble of byte 0 first; then swap pairs: the stream
        // has the HIGH nibble of every byte first (:767-773)
        uint32_t y = (w[j] >> 3) & 0x11111111u;
        y = (y | (y >> 3)) & 0x03030303u;
        y = (y | (y >> 6)) & 0x000F000Fu;
        y = (y | (y >> 12)) & 0xFFu;
        y = ((y & 0x55u) << 1) | ((y & 0xAAu) >> 1);
        hmask |= (unsigned long long)y << (8 * j);
    }
    b.T = ~hmask & b.V;
}

// transfer function of a block from its masks.  COMPRESS: a hit emits the pending pair (state 1), a miss emits itself and,
// in state 1, the stranded hit in front of it.  else: every nibble emits, except the first half of a literal.
template <bool COMPRESS, typename M>
__device__ __forceinline__ uint32_t block_fn_of(const TxBlock<M> &b) {
    if (b.units == 0) return kFnIdentity;
    M E, F;
    run_parity(b.T, E, F);
    const M E1 = E ^ F;
    const M Q0 = E << 1, Q1 = (E1 << 1) | (M)1;
    uint32_t e0, e1;
    if (COMPRESS) {
        const int misses = popcnt(b.V & ~b.T);
        e0 = (uint32_t)(misses + popcnt(Q0 & b.V));
        e1 = (uint32_t)(misses + popcnt(Q1 & b.V));
    } else {
        e0 = (uint32_t)popcnt(b.V & (~b.T | Q0));
        e1 = (uint32_t)popcnt(b.V & (~b.T | Q1));
    }
    const int last = b.units - 1;
    return fn_pack(e0, e1, (uint32_t)(E >> last) & 1u, (uint32_t)(E1 >> last) & 1u);
}

// ------------------------------------------------------------------------------------------ S1
template <bool COMPRESS, bool ADAPT>
__global__ void __launch_bounds__(kTxThreads) tx_summary_kernel(const uint8_t *__restrict__ in, size_t n, TxWorkspace ws, size_t ntiles,
                                                                int32_t *__restrict__ d_status, const uint8_t *__restrict__ pos) {
    __shared__ uint32_t s_warp[kTxWarps];
    __shared__ uint8_t s_lut[128];
    fill_letter_table(s_lut);
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool bad = false;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * kTxTile + (size_t)tid * kTxPer;
        uint32_t f = kFnIdentity;
        if (base < n) {
            uint32_t prev, w[kTxPer / 4];
            load_block(in, n, base, prev, w);
            if (COMPRESS) {
                TxBlock<uint32_t> b;
                uint32_t li[kTxPer];
                compress_masks<ADAPT>(s_lut, pos, w, base, n, b, li, bad);
                f = block_fn_of<true>(b);
            } else {
                TxBlock<unsigned long long> b;
                decompress_masks(w, base, n, b);
                f = block_fn_of<false>(b);
            }
        }
        // ordered reduction over the warp, then over the CTA's warps
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t g = __shfl_down_sync(0xFFFFFFFFu, f, d);
            if ((lane & (2 * d - 1)) == 0) f = fn_compose(f, g);
        }
        if (lane == 0) s_warp[warp] = f;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = s_warp[0];
#pragma unroll
            for (int w = 1; w < kTxWarps; w++) t = fn_compose(t, s_warp[w]);
            ws.tile_fn[tile] = t;
        }
        __syncthreads();
    }
    if (COMPRESS && bad) set_status(d_status, DC_ERR_SYMBOL);
}

// ------------------------------------------------------------------------------------------ S2
// the same functions with 64-bit counts, for the scan over all tiles
struct Fn64 {
    unsigned long long e0, e1;
    uint32_t q;  // bit 0: next from 0, bit 1: next from 1
};
__device__ __forceinline__ Fn64 fn64_of(uint32_t f) { return Fn64{fn_emit(f, 0), fn_emit(f, 1), (f >> 30) & 3u}; }
__device__ __forceinline__ Fn64 fn64_compose(const Fn64 &a, const Fn64 &b) {
    const uint32_t a0 = a.q & 1u, a1 = (a.q >> 1) & 1u;
    Fn64 r;
    r.e0 = a.e0 + (a0 ? b.e1 : b.e0);
    r.e1 = a.e1 + (a1 ? b.e1 : b.e0);
    r.q = ((b.q >> a0) & 1u) | (((b.q >> a1) & 1u) << 1);
    return r;
}

constexpr int kTxScanThreads = 1024;

template <bool COMPRESS>
__global__ void __launch_bounds__(kTxScanThreads) tx_scan_kernel(const uint8_t *__restrict__ in, size_t n, TxWorkspace ws, size_t ntiles,
                                                                 uint8_t *__restrict__ out, size_t out_cap,
                                                                 unsigned long long *__restrict__ d_out_len, int32_t *__restrict__ d_status) {
    __shared__ Fn64 s_warp[32];
    __shared__ unsigned long long s_off;
    __shared__ uint32_t s_q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_off = 0; s_q = 0; }  // the stream starts in state 0 with nothing emitted
    __syncthreads();
    // blocks of 1024 x 8 tiles: a thread composes its 8 consecutive tiles, the CTA scans the 1024 results by
    // composition, and every thread walks its tiles again from its now known entry state and offset
    constexpr int kPer = 8;
    for (size_t blk = 0; blk < ntiles; blk += (size_t)kTxScanThreads * kPer) {
        const size_t t0 = blk + (size_t)tid * kPer;
        uint32_t f[kPer];
#pragma unroll
        for (int k = 0; k < kPer; k++) f[k] = t0 + k < ntiles ? ws.tile_fn[t0 + k] : kFnIdentity;
        Fn64 mine = fn64_of(f[0]);
#pragma unroll
        for (int k = 1; k < kPer; k++) mine = fn64_compose(mine, fn64_of(f[k]));
        Fn64 incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Fn64 g;
            g.e0 = __shfl_up_sync(0xFFFFFFFFu, incl.e0, d);
            g.e1 = __shfl_up_sync(0xFFFFFFFFu, incl.e1, d);
            g.q = __shfl_up_sync(0xFFFFFFFFu, incl.q, d);
            if (lane >= d) incl = fn64_compose(g, incl);
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        // entry of this thread = (block entry) through the earlier warps, then the earlier lanes of this warp
        unsigned long long off = s_off;
        uint32_t q = s_q;
        for (int w = 0; w < warp; w++) {
            const Fn64 g = s_warp[w];
            off += q ? g.e1 : g.e0;
            q = (g.q >> q) & 1u;
        }
        {
            Fn64 g;
            g.e0 = __shfl_up_sync(0xFFFFFFFFu, incl.e0, 1);
            g.e1 = __shfl_up_sync(0xFFFFFFFFu, incl.e1, 1);
            g.q = __shfl_up_sync(0xFFFFFFFFu, incl.q, 1);
            if (lane != 0) {
                off += q ? g.e1 : g.e0;
                q = (g.q >> q) & 1u;
            }
        }
#pragma unroll
        for (int k = 0; k < kPer; k++) {
            if (t0 + k < ntiles) {
                ws.tile_off[t0 + k] = off;
                ws.tile_q[t0 + k] = (uint8_t)q;
            }
            off += fn_emit(f[k], q);
            q = fn_next(f[k], q);
        }
        __syncthreads();
        if (tid == kTxScanThreads - 1) { s_off = off; s_q = q; }  // the last thread's exit = the block's exit
        __syncthreads();
    }
    const unsigned long long off = s_off;
    const uint32_t q = s_q;
    if (tid == kTxScanThreads - 1) {
        // the thread that owns the last tile (or none: then off/q are the whole stream's) finishes the stream
        const unsigned long long body = off + q;  // a pending half (compress: odd hit; decompress: half literal) is one more symbol
        unsigned long long len;
        int32_t mode;
        if (COMPRESS) {
            len = 2 + body;                   // 0xAF, first byte, body
            mode = 0;
            if (len >= n) { len = n + 1; mode = 1; }   // not shorter: ' ' + raw copy (:1018-1037)
        } else {
            const uint32_t type = in[0];
            if (type == 0xAFu) { len = n >= 2 ? 1 + body : 0; mode = 0; }
            else if (type == ' ') { len = n - 1; mode = 1; }
            else { len = n; mode = 2; }
        }
        if (len > out_cap) {
            set_status(d_status, DC_ERR_CAPACITY);
            mode = -1;
        } else if (len < out_cap) {
            out[len] = 0;  // the reference's strings are NUL-terminated (:1011, :814)
        }
        *ws.out_len = len;
        *ws.mode = mode;
        if (d_out_len) *d_out_len = len;
    }
}

__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// emit the symbols of a compress block from entry state q_in into shared memory at dst (a shared-space address).
// Every byte knows the state in front of it (Q) and its output offset (a popcount), so nothing is carried along.
__device__ __forceinline__ void emit_compress(const TxBlock<uint32_t> &b, const uint32_t (&w)[kTxPer / 4], const uint32_t (&li)[kTxPer], uint32_t prev,
                                              uint32_t prev_li, uint32_t q_in, uint32_t dst, bool ends_here, const uint8_t *__restrict__ in,
                                              size_t n) {
    uint32_t E, F;
    run_parity(b.T, E, F);
    if (q_in) E ^= F;
    const uint32_t Q = (E << 1) | q_in;
    const uint32_t A = Q & b.V;        // a byte in front of which a hit is pending: a hit completes the pair, a miss strands it
    const uint32_t Mm = b.V & ~b.T;    // misses: literals
    // a byte writes at most two output bytes: "first" (the pair it completes, or the stranded hit in front of it) and
    // "second" (itself, if it is a miss).  Two predicated stores per byte, offsets by popcount.
    uint32_t pc = prev, pl8 = 8u | prev_li;
#pragma unroll
    for (int k = 0; k < kTxPer; k++) {
        const uint32_t bit = 1u << k, below = bit - 1u;
        const uint32_t off = dst + __popc(A & below) + __popc(Mm & below);
        const uint32_t ck = byte_of(w, k), l8 = 8u | li[k];
        if (A & bit) sts8(off, (b.T & bit) ? ((pl8 << 4) | l8) : pc);
        if (Mm & bit) sts8(off + ((A >> k) & 1u), ck);
        pc = ck;
        pl8 = l8;
    }
    // trailing half byte -> literal (:1000-1009): one thread of the whole grid
    if (ends_here && b.units > 0 && ((E >> (b.units - 1)) & 1u)) sts8(dst + __popc(A) + __popc(Mm), in[n - 1]);
}

// the same for a decompress block: 64 nibbles, in two halves of 32 so that the masks stay 32-bit.  ADAPT: a hit nibble is
// written as 0x80 | position (K8 resolves it, the letter depends on the bytes decoded before it)
template <bool ADAPT>
__device__ __forceinline__ void emit_decompress(const TxBlock<unsigned long long> &b, const uint32_t (&w)[kTxPer / 4], uint32_t prev, uint32_t q_in,
                                                uint32_t dst, bool ends_here, const uint8_t *__restrict__ in, size_t n) {
    unsigned long long E, F;
    run_parity(b.T, E, F);
    if (q_in) E ^= F;
    const unsigned long long Q = (E << 1) | q_in;
    const unsigned long long Em = b.V & (~b.T | Q);  // every nibble emits, except the first half of a literal
    const uint32_t em[2] = {(uint32_t)Em, (uint32_t)(Em >> 32)}, qq[2] = {(uint32_t)Q, (uint32_t)(Q >> 32)};
    const uint32_t base1 = __popc(em[0]);
    uint32_t pn = prev & 0xFu;
#pragma unroll
    for (int j = 0; j < 2 * kTxPer; j++) {
        const int h = j >> 5;
        const uint32_t bit = 1u << (j & 31), below = bit - 1u;
        const uint32_t off = dst + (h ? base1 : 0u) + __popc(em[h] & below);
        const uint32_t cj = byte_of(w, j >> 1);
        const uint32_t x = (j & 1) ? (cj & 0xFu) : (cj >> 4);
        if (em[h] & bit) sts8(off, (qq[h] & bit) ? ((pn & 7u) << 4) + x : (ADAPT ? 0x80u | (x & 7u) : letter_of(x)));
        pn = x;
    }
    // a dangling half literal at the very end is completed with a zero nibble: one thread of the whole grid
    if (ends_here && b.units > 0 && ((E >> (b.units - 1)) & 1ull)) sts8(dst + base1 + __popc(em[1]), ((uint32_t)in[n - 1] & 7u) << 4);
}

// ------------------------------------------------------------------------------------------ S3
constexpr int kTxStageBytes = kTxThreads * 2 * kTxPer + 48;  // decompress: two symbols per byte; + alignment slack

template <bool COMPRESS, bool ADAPT>
__global__ void __launch_bounds__(kTxThreads) tx_emit_kernel(const uint8_t *__restrict__ in, size_t n, TxWorkspace ws, size_t ntiles,
                                                             uint8_t *__restrict__ out, const uint8_t *__restrict__ pos) {
    __shared__ __align__(16) uint8_t s_stage[kTxStageBytes];
    __shared__ uint32_t s_warp[kTxWarps];
    __shared__ uint8_t s_lut[128];
    fill_letter_table(s_lut);
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t mode = *ws.mode;
    if (mode < 0) return;
    const bool raw = COMPRESS ? mode == 1 : mode != 0;
    const size_t raw_shift = COMPRESS ? 0 : (mode == 1 ? 1 : 0);  // decompress: skip the type byte of a LITERAL block
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t base = tile * kTxTile + (size_t)tid * kTxPer;
        if (raw) {
            if (COMPRESS && base == 0) out[0] = ' ';
            for (int k = 0; k < kTxPer; k++) {
                const size_t i = base + k;
                if (i < n && i + 1 > raw_shift) out[i - raw_shift + (COMPRESS ? 1 : 0)] = in[i];
            }
            continue;
        }
        uint32_t prev = 0, w[kTxPer / 4], f = kFnIdentity;
        uint32_t li[COMPRESS ? kTxPer : 1];
        TxBlock<uint32_t> bc;
        TxBlock<unsigned long long> bd;
        bc.units = bd.units = 0;
        bool bad = false;
        if (base < n) {
            load_block(in, n, base, prev, w);
            if (COMPRESS) {
                compress_masks<ADAPT>(s_lut, pos, w, base, n, bc, (uint32_t(&)[kTxPer])li, bad);
                f = block_fn_of<true>(bc);
            } else {
                decompress_masks(w, base, n, bd);
                f = block_fn_of<false>(bd);
            }
        }
        // inclusive scan over the warp by composition, warp totals to shared memory
        uint32_t incl = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t g = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl = fn_compose(g, incl);
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();  // also: the previous tile's copy-out has read the staging buffer
        uint32_t q = ws.tile_q[tile], off = 0;
        for (int w = 0; w < warp; w++) {
            const uint32_t g = s_warp[w];
            off += fn_emit(g, q);
            q = fn_next(g, q);
        }
        const uint32_t total_q_in = q;  // entry state of this warp
        {
            const uint32_t g = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
            if (lane != 0) { off += fn_emit(g, total_q_in); q = fn_next(g, total_q_in); }
        }
        // tile totals (every thread computes them the same way: needed for the copy-out)
        uint32_t tq = ws.tile_q[tile], tile_total = 0;
        for (int w = 0; w < kTxWarps; w++) {
            const uint32_t g = s_warp[w];
            tile_total += fn_emit(g, tq);
            tq = fn_next(g, tq);
        }
        const bool last_tile = tile == ntiles - 1;
        if (last_tile) tile_total += tq;  // the pending half symbol at the end of the stream
        const unsigned long long ob = (COMPRESS ? 2ull : 1ull) + ws.tile_off[tile];
        const uint32_t a = (uint32_t)(((uintptr_t)out + ob) & 15);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_stage) + a + off;
        if (base < n) {
            const bool ends_here = base + kTxPer >= n;  // this block holds the last byte of the stream
            if (COMPRESS) {
                const uint32_t prev_li = ADAPT ? (base ? pos[base - 1] & 0xFu : 8u) : s_lut[prev & 0x7Fu];
                emit_compress(bc, w, (const uint32_t(&)[kTxPer])li, prev, prev_li, q, dst, ends_here, in, n);
            } else {
                emit_decompress<ADAPT>(bd, w, prev, q, dst, ends_here, in, n);
            }
        }
        if (tile == 0 && tid == 0) {
            if (COMPRESS) { out[0] = 0xAF; out[1] = in[0]; }   // :903, :905
            else if (n >= 2) out[0] = in[1];                      // :750
        }
        __syncthreads();
        // copy-out: staging byte i <-> out[ob - a + i]; 16-byte words are aligned on both sides
        const uint32_t span = a + tile_total;
        const uint32_t jfull = span >> 4, head = a ? 1u : 0u;
        for (uint32_t j = head + tid; j < jfull; j += kTxThreads) stg_stream((uint4 *)(out + (ob - a)) + j, *(const uint4 *)(s_stage + j * 16));
        if (tid < 16) {
            if ((a != 0 || jfull == 0) && (uint32_t)tid >= a && (uint32_t)tid < span) out[ob - a + tid] = s_stage[tid];
        } else if (tid < 32) {
            const uint32_t k = jfull * 16 + (tid - 16);
            if (jfull != 0 && k >= a && k < span) out[ob - a + k] = s_stage[k];
        }
    }
}

static size_t tx_ws_layout(size_t n, bool adapt_compress, size_t off[7]) {
    const size_t ntiles = (n + kTxTile - 1) / kTxTile;
    size_t p = 64;
    auto take = [&](size_t bytes) { size_t o = p; p += (bytes + 63) & ~(size_t)63; return o; };
    size_t o[7];
    o[0] = take(ntiles * 4);   // tile_fn
    o[1] = take(ntiles * 8);   // tile_off
    o[2] = take(ntiles);       // tile_q
    o[3] = 0;                  // out_len (header)
    o[4] = 8;                  // mode (header)
    o[5] = adapt_compress ? take(n + 16) : 0;                   // K8: move-to-front position of every byte
    o[6] = adapt_compress ? take(mtf_workspace_bytes(n)) : 0;   // K8: lists per block / chunk
    if (off) for (int i = 0; i < 7; i++) off[i] = o[i];
    return p;
}

template <bool COMPRESS, bool ADAPT>
static int tx_run(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t cap, uint64_t *d_out_len, int32_t *d_status, void *d_ws,
                  size_t ws_bytes, cudaStream_t st) {
    if ((n && (!d_src || !d_dst || !d_ws)) || ((uintptr_t)d_ws & 15)) return DC_ERR_ARG;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (d_out_len) DC_CUDA_TRY(cudaMemsetAsync(d_out_len, 0, sizeof(uint64_t), st));
    if (n == 0) {
        if (cap && d_dst) DC_CUDA_TRY(cudaMemsetAsync(d_dst, 0, 1, st));  // the empty string
        return DC_OK;
    }
    size_t off[7];
    if (ws_bytes < tx_ws_layout(n, COMPRESS && ADAPT, off)) return DC_ERR_CAPACITY;
    char *w = (char *)d_ws;
    TxWorkspace ws;
    ws.tile_fn = (uint32_t *)(w + off[0]);
    ws.tile_off = (unsigned long long *)(w + off[1]);
    ws.tile_q = (uint8_t *)(w + off[2]);
    ws.out_len = (unsigned long long *)(w + off[3]);
    ws.mode = (int32_t *)(w + off[4]);
    const uint8_t *pos = nullptr;
    if (COMPRESS && ADAPT) {
        const int rc = mtf_positions(d_src, n, (uint8_t *)(w + off[5]), w + off[6], st);
        if (rc != DC_OK) return rc;
        pos = (const uint8_t *)(w + off[5]);
    }
    const size_t ntiles = (n + kTxTile - 1) / kTxTile;
    const unsigned int grid = (unsigned int)min(ntiles, (size_t)sm_count() * 8);
    {
        LaunchScope ls(DC_K_TEXT_SUMMARY, st);
        tx_summary_kernel<COMPRESS, COMPRESS && ADAPT><<<grid, kTxThreads, 0, st>>>(d_src, n, ws, ntiles, d_status, pos);
    }
    {
        LaunchScope ls(DC_K_TEXT_SCAN, st);
        tx_scan_kernel<COMPRESS><<<1, kTxScanThreads, 0, st>>>(d_src, n, ws, ntiles, d_dst, cap, (unsigned long long *)d_out_len, d_status);
    }
    {
        LaunchScope ls(DC_K_TEXT_EMIT, st);
        tx_emit_kernel<COMPRESS, ADAPT><<<grid, kTxThreads, 0, st>>>(d_src, n, ws, ntiles, d_dst, pos);
    }
    if (!COMPRESS && ADAPT) {
        const int rc = mtf_resolve(d_dst, ws.out_len, ws.mode, st);
        if (rc != DC_OK) return rc;
    }
    return cuda_status(cudaGetLastError());
}

}  // namespace dc

using namespace dc;

extern "C" size_t dc_nybble_text_workspace_bytes(size_t n) { return tx_ws_layout(n, false, nullptr); }
extern "C" size_t dc_nybble_adaptive_workspace_bytes(size_t n) { return tx_ws_layout(n, true, nullptr); }

extern "C" int dc_nybble_text_compress(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t dst_capacity, uint64_t *d_out_len,
                                       int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream) {
    return tx_run<true, false>(d_src, n, d_dst, dst_capacity, d_out_len, d_status, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dc_nybble_text_decompress(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t dst_capacity, uint64_t *d_out_len,
                                         int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream) {
    return tx_run<false, false>(d_src, n, d_dst, dst_capacity, d_out_len, d_status, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dc_nybble_adaptive_compress(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t dst_capacity, uint64_t *d_out_len,
                                           int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream) {
    return tx_run<true, true>(d_src, n, d_dst, dst_capacity, d_out_len, d_status, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dc_nybble_adaptive_decompress(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t dst_capacity, uint64_t *d_out_len,
                                             int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream) {
    return tx_run<false, true>(d_src, n, d_dst, dst_capacity, d_out_len, d_status, d_workspace, workspace_bytes, (cudaStream_t)stream);
}
