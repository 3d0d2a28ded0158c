// k3_encode.cu -- K3: Huffman payload encode.
//
// Replaces represent_items_with_codes() (n_ary_huffman.c:1621-1678).  In the reference that function is a
// stub (assert(0) at :1661, returns 32767); its stated intent is "for each input byte append
// encode_length digits of encode_value, completely bit-oriented" (:1651-1657).  Payload layout
// (DESIGN.md): a code of `len` digits is the len*log2(n)-bit big-endian numeral of encode_value; codes are
// concatenated in input order, MSB-first within each byte; the final byte is zero-padded.
//
// The input is cut into CHUNKS of 4 KB (one warp each); 8 chunks form a 32 KB RUN (one CTA).
//
// SINGLE PASS (tables whose longest code is <= 16 bits; see "single pass" below): one launch, one read of the
// input.  A warp stages its whole chunk in shared memory at chunk-local bit offsets, the run's bit offset comes
// from a decoupled look-back over run descriptors, and the staging buffer is copied out funnel-shifted to the
// global bit alignment.  HBM traffic = N + C.
//
// THREE LAUNCHES (tables with codes up to 32 bits; the kernels directly below):
//   E1 count   : bits per chunk = sum of code lengths (one streaming read of the input); per run the
//                exclusive offsets of its 8 chunks and the run total
//   E2 scan    : exclusive scan of the run totals -> global bit offset of every run, total bit count
//   E3 encode  : every WARP walks its chunk in 8 sub-tiles of 512 symbols (32 lanes x 16 symbols):
//        A. 16 code look-ups from a shared-memory copy of the table (64-bit entries)
//        B. warp-wide exclusive scan of the per-lane bit totals (shuffles)
//        C. every lane streams its codes through a 64-bit accumulator into the warp's staging buffer at its
//           bit offset.  When every lane holds >= 32 bits each staging word is shared by at most two
//           neighbouring lanes, so words are written with plain stores and the one shared partial word
//           travels by shuffle; otherwise (very compressible or ragged data) shared-memory OR is used
//        D. coalesced copy-out: staging words are funnel-shifted to the global bit alignment and stored
//           as 16-byte words.  The partial 16-byte word at the end of a sub-tile is carried into the next
//           one in registers; the one shared between two CHUNKS is merged by whichever warp arrives
//           second (both sides deposit their half in the workspace and bump a counter), so every output
//           byte is written exactly once, in any scheduling order, with no pre-zeroed output.
//   HBM traffic = 2N (count + encode reads) + C (write).
// Both paths share the sub-tile steps A-C (lookup_items, emit_bits*, the shuffle scan).
#include "chain_scan.cuh"
#include "dc_common.cuh"

namespace dc {

constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kEncPerThread = 16;
constexpr int kSubTile = 32 * kEncPerThread;           // 512 symbols per warp step
constexpr int kChunkSubs = 8;
constexpr int kChunkBytes = kSubTile * kChunkSubs;      // 4 KB per warp
constexpr int kRunBytes = kChunkBytes * kEncWarps;      // 32 KB per CTA
constexpr int kNarrowBits = 16;                         // code pairs fit 32 bits
constexpr int kPlannedMaxBits = 12;                     // the planned single pass (encode_fast_kernel) takes tables up to this

template <bool WIDE>
struct EncCfg {
    static constexpr int kMaxBits = WIDE ? 32 : kNarrowBits;
    static constexpr int kStageWords = 4 + kSubTile * kMaxBits / 32 + 12;  // carried tail + worst case + pad (per warp)
};

struct EncWorkspace {
    uint32_t *run_bits;            // [nruns]      E1
    uint32_t *chunk_rel;           // [nruns * 8]  E1: bit offset of each chunk inside its run
    unsigned long long *run_off;   // [nruns+1]    E2 (exclusive; last = total)
    uint32_t *bstate;              // [nchunks+1]  arrivals at the boundary word between chunk b-1 and chunk b
    uint4 *bleft, *bright;         // [nchunks+1]  the two halves of that word (big-endian word domain)
    const uint32_t *d_phase;       // != nullptr: the bit phase is read from device memory (a shard learns it from a collective)
};

// bits [bit, bit+32) of a big-endian word array
__device__ __forceinline__ uint32_t stage_word(const uint32_t *stage, uint32_t bit) {
    const uint32_t a = bit >> 5, s = bit & 31;
    return __funnelshift_l(stage[a + 1], stage[a], s);
}

// stage[wi] |= word when pred (predicated RED.OR on shared memory: no branch, no BSSY/BSYNC)
__device__ __forceinline__ void stage_or_if(uint32_t *addr, uint32_t word, bool pred) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(addr);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.or.b32 [%0], %1;\n\t}" ::"r"(a), "r"(word),
                 "r"((uint32_t)pred)
                 : "memory");
}

// append `len` (<= 32) bits of `val` to the 64-bit accumulator (hi:lo), then emit one 32-bit word if at least 32
// bits are pending.  nb = pending bits (< 32 on entry and on exit); wi = next staging word.
__device__ __forceinline__ void emit_bits(uint32_t *stage, uint32_t &hi, uint32_t &lo, uint32_t &nb, uint32_t &wi, uint32_t val,
                                          uint32_t len) {
    hi = __funnelshift_lc(lo, hi, len);
    lo = __funnelshift_lc(0u, lo, len) | val;
    nb += len;
    // pending bits are the low nb bits of hi:lo; the oldest 32 of them are (hi:lo) >> (nb - 32)
    stage_or_if(stage + wi, __funnelshift_r(lo, hi, nb), nb >= 32u);
    wi += nb >> 5;
    nb &= 31u;
}

// same as emit_bits, for the case where every word completed by this thread is owned by it alone (plain store)
__device__ __forceinline__ void emit_bits_owned(uint32_t *stage, uint32_t &hi, uint32_t &lo, uint32_t &nb, uint32_t &wi,
                                                uint32_t val, uint32_t len) {
    hi = __funnelshift_lc(lo, hi, len);
    lo = __funnelshift_lc(0u, lo, len) | val;
    nb += len;
    if (nb >= 32u) stage[wi] = __funnelshift_r(lo, hi, nb);
    wi += nb >> 5;
    nb &= 31u;
}

__device__ __forceinline__ bool table_usable(const dc_huff_table *tab, int32_t *d_status) {
    const int tstatus = tab->status, bpd = tab->bits_per_digit;
    if (tstatus == DC_OK && bpd != 0) return true;
    if (blockIdx.x == 0 && threadIdx.x == 0) set_status(d_status, tstatus != DC_OK ? tstatus : DC_ERR_RADIX);
    return false;
}

// ------------------------------------------------------------------------------------------ E1 count
__global__ void __launch_bounds__(kEncThreads) encode_count_kernel(const uint8_t *__restrict__ in, size_t n,
                                                                   const dc_huff_table *__restrict__ tab, EncWorkspace ws,
                                                                   unsigned int nruns, int32_t *__restrict__ d_status) {
    __shared__ uint32_t s_len[256];
    __shared__ uint32_t s_chunk[kEncWarps];
    if (!table_usable(tab, d_status)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_len[tid] = (uint32_t)(tab->enc64[tid] >> 32);
    __syncthreads();
    bool missing = false;
    for (unsigned int run = blockIdx.x; run < nruns; run += gridDim.x) {
        const size_t base = (size_t)run * kRunBytes + (size_t)warp * kChunkBytes;  // this warp's chunk
        uint32_t sum = 0;
        if (base + kChunkBytes <= n) {
            uint4 v[kChunkSubs];
#pragma unroll
            for (int j = 0; j < kChunkSubs; j++) v[j] = ldg_stream((const uint4 *)(in + base) + j * 32 + lane);
#pragma unroll
            for (int j = 0; j < kChunkSubs; j++) {
                const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t l = s_len[(w[k >> 2] >> (8 * (k & 3))) & 0xFFu];
                    missing |= l == 0;
                    sum += l;
                }
            }
        } else {
            for (size_t i = base + lane; i < n && i < base + kChunkBytes; i += 32) {
                const uint32_t l = s_len[in[i]];
                missing |= l == 0;
                sum += l;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if (lane == 0) s_chunk[warp] = sum;
        __syncthreads();
        if (tid < kEncWarps) {
            uint32_t rel = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kEncWarps; w++) {
                const uint32_t x = s_chunk[w];
                if (w < tid) rel += x;
                tot += x;
            }
            ws.chunk_rel[(size_t)run * kEncWarps + tid] = rel;
            if (tid == 0) ws.run_bits[run] = tot;
        }
        __syncthreads();
    }
    if (missing) set_status(d_status, DC_ERR_SYMBOL);
}

// ------------------------------------------------------------------------------------------ E2 scan
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;

__global__ void __launch_bounds__(kScanThreads) encode_scan_kernel(const dc_huff_table *__restrict__ tab, EncWorkspace ws,
                                                                   unsigned int nruns,
                                                                   unsigned long long *__restrict__ d_total_bits) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ok = tab->status == DC_OK && tab->bits_per_digit != 0;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (unsigned int base = 0; base < nruns; base += kScanThreads * kScanItems) {
        const unsigned int first = base + tid * kScanItems;
        uint32_t item[kScanItems];
        unsigned long long mine = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            item[k] = (ok && first + k < nruns) ? ws.run_bits[first + k] : 0u;
            mine += item[k];
        }
        unsigned long long incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long off = s_carry;
        for (int w = 0; w < warp; w++) off += s_warp[w];
        off += incl - mine;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            if (first + k < nruns) ws.run_off[first + k] = off;
            off += item[k];
        }
        __syncthreads();
        if (tid == kScanThreads - 1) s_carry = off;
        __syncthreads();
    }
    if (tid == 0) {
        ws.run_off[nruns] = s_carry;
        if (d_total_bits) *d_total_bits = s_carry;
    }
}

// ------------------------------------------------------------------------------------------ E3 encode

// deposit one half of the 16-byte word shared by two runs; the second arrival merges and stores it
__device__ __forceinline__ void boundary_merge(const EncWorkspace &ws, unsigned int b, bool left_side, uint4 mine,
                                               uint8_t *__restrict__ out, unsigned long long vec, size_t stream_bytes) {
    uint4 *my_slot = left_side ? ws.bleft + b : ws.bright + b;
    const uint4 *other_slot = left_side ? ws.bright + b : ws.bleft + b;
    *my_slot = mine;
    __threadfence();
    if (atomicAdd(&ws.bstate[b], 1u) == 1u) {
        __threadfence();
        const uint4 o = ld_cg_u128(other_slot);
        uint4 r;
        r.x = bswap32(mine.x | o.x);
        r.y = bswap32(mine.y | o.y);
        r.z = bswap32(mine.z | o.z);
        r.w = bswap32(mine.w | o.w);
        if ((vec + 1) * 16 <= stream_bytes) {
            stg_stream((uint4 *)out + vec, r);
        } else {  // the shared word is also the last, partial word of the stream
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
            for (size_t i = vec * 16; i < stream_bytes; i++) out[i] = (uint8_t)(rw[(i & 15) >> 2] >> (8 * (i & 3)));
        }
    }
}

// ---- A. 16 symbols -> 8 code pairs (narrow) or 16 codes (wide), kept in registers; returns the thread's bit total
template <bool WIDE, bool FULL, typename entry_t, int kItems>
__device__ __forceinline__ uint32_t lookup_items(const entry_t *s_enc, const uint32_t (&w)[4], int valid,
                                                 uint32_t (&item_val)[kItems], uint32_t (&item_len)[kItems]) {
    uint32_t bits = 0;
    if (WIDE) {
#pragma unroll
        for (int k = 0; k < kEncPerThread; k++) {
            unsigned long long e = s_enc[(w[k >> 2] >> (8 * (k & 3))) & 0xFFu];
            if (!FULL && k >= valid) e = 0;
            item_val[k % kItems] = (uint32_t)e;
            item_len[k % kItems] = (uint32_t)(e >> 32);
            bits += (uint32_t)(e >> 32);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kEncPerThread; k += 2) {
            uint32_t e0 = (uint32_t)s_enc[(w[k >> 2] >> (8 * (k & 3))) & 0xFFu];
            uint32_t e1 = (uint32_t)s_enc[(w[k >> 2] >> (8 * ((k + 1) & 3))) & 0xFFu];
            if (!FULL && k >= valid) e0 = 0;
            if (!FULL && k + 1 >= valid) e1 = 0;
            const uint32_t l1 = e1 & 63u;
            item_val[(k / 2) % kItems] = ((e0 >> 6) << l1) | (e1 >> 6);
            item_len[(k / 2) % kItems] = (e0 & 63u) + l1;
            bits += (e0 & 63u) + l1;
        }
    }
    return bits;
}

template <bool WIDE>
__global__ void __launch_bounds__(kEncThreads, WIDE ? 2 : 5) encode_run_kernel(const uint8_t *__restrict__ in, size_t n,
                                                                 const dc_huff_table *__restrict__ tab, uint8_t *__restrict__ out,
                                                                 size_t out_cap, unsigned phase, EncWorkspace ws,
                                                                 unsigned int nruns, int32_t *__restrict__ d_status, int exact) {
    typedef typename std::conditional<WIDE, unsigned long long, uint32_t>::type entry_t;
    constexpr int kStageWords = EncCfg<WIDE>::kStageWords;
    constexpr int kItems = WIDE ? kEncPerThread : kEncPerThread / 2;  // codes (WIDE) or code pairs per lane
    __shared__ __align__(16) uint32_t s_stage[kEncWarps][kStageWords];
    __shared__ entry_t s_enc[256];

    if (tab->status != DC_OK || tab->bits_per_digit == 0) return;  // reported by the count kernel
    if (ws.d_phase) phase = *ws.d_phase & 7u;
    if (WIDE && tab->max_bits <= kNarrowBits) {                   // the single-pass kernels handle this table
        // (exact: launched alone because the host's copy of the table header said so -- that header was not this table's)
        if (exact && blockIdx.x == 0 && threadIdx.x == 0) set_status(d_status, DC_ERR_ARG);
        return;
    }

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // bytes [0, stream_bytes) of `out` are written, nothing else; refuse the whole stream if they do not fit
    const size_t stream_bytes = (size_t)(((unsigned long long)phase + ws.run_off[nruns] + 7) >> 3);
    if (stream_bytes > out_cap) {
        if (blockIdx.x == 0 && tid == 0) set_status(d_status, DC_ERR_CAPACITY);
        return;
    }
    s_enc[tid] = WIDE ? (entry_t)tab->enc64[tid] : (entry_t)tab->enc[tid];
    __syncthreads();
    uint32_t *stage = s_stage[warp];
    const size_t nchunks = (n + kChunkBytes - 1) / kChunkBytes;

    __shared__ uint32_t s_cb[kEncWarps];
    for (unsigned int run = blockIdx.x; run < nruns; run += gridDim.x) {
        const size_t chunk = (size_t)run * kEncWarps + warp;
        // bit offsets of the run's chunks: a count of this warp's chunk first (the run's own offset is planned, the chunks'
        // are not; the second read of the chunk comes out of L2)
        uint32_t my_cb = 0;
        if (chunk < nchunks) {
            const size_t cbase = chunk * kChunkBytes, cend = min(n, cbase + (size_t)kChunkBytes);
            for (size_t i = cbase + (size_t)lane * 16; i < cend; i += 512) {
                if (i + 16 <= cend) {
                    const uint4 v = ldg_stream((const uint4 *)(in + i));
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const entry_t e = s_enc[(wv[k >> 2] >> (8 * (k & 3))) & 0xFFu];
                        my_cb += WIDE ? (uint32_t)((unsigned long long)e >> 32) : ((uint32_t)e & 63u);
                    }
                } else {
                    for (size_t j = i; j < cend; j++) {
                        const entry_t e = s_enc[in[j]];
                        my_cb += WIDE ? (uint32_t)((unsigned long long)e >> 32) : ((uint32_t)e & 63u);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_cb += __shfl_xor_sync(0xFFFFFFFFu, my_cb, o);
        }
        __syncthreads();   // the previous run's s_cb has been read
        if (lane == 0) s_cb[warp] = my_cb;
        __syncthreads();
        uint32_t chunk_rel = 0;
#pragma unroll
        for (int w2 = 0; w2 < kEncWarps; w2++) chunk_rel += w2 < warp ? s_cb[w2] : 0u;
        if (chunk >= nchunks) continue;
        const size_t chunk_base = chunk * kChunkBytes;
        const size_t chunk_len = min((size_t)kChunkBytes, n - chunk_base);
        const int nsub = (int)((chunk_len + kSubTile - 1) / kSubTile);
        const int nfull = (int)(chunk_len / kSubTile);  // sub-tiles in which every lane has 16 symbols
        unsigned long long g = (unsigned long long)phase + ws.run_off[run] + chunk_rel;  // next bit to write
        const bool last_chunk = chunk == nchunks - 1;
        uint32_t carried = 0;  // lanes 0..3: the last 128 bits of the previous sub-tile of this chunk
        const uint8_t *src = in + chunk_base + (size_t)lane * kEncPerThread;  // this lane's 16 symbols of sub-tile 0

        uint4 next = make_uint4(0, 0, 0, 0);
        if (nfull > 0) next = ldg_stream((const uint4 *)src);

#pragma unroll 1
        for (int t = 0; t < nsub; t++, src += kSubTile) {
            uint32_t item_val[kItems], item_len[kItems];
            uint32_t my_bits;
            if (t < nfull) {
                const uint32_t w[4] = {next.x, next.y, next.z, next.w};
                if (t + 1 < nfull) next = ldg_stream((const uint4 *)(src + kSubTile));
                my_bits = lookup_items<WIDE, true, entry_t, kItems>(s_enc, w, kEncPerThread, item_val, item_len);
            } else {  // the ragged last sub-tile of the stream
                const size_t base = chunk_base + (size_t)t * kSubTile + (size_t)lane * kEncPerThread;
                const int valid = base < n ? (int)min((size_t)kEncPerThread, n - base) : 0;
                uint32_t w[4] = {0, 0, 0, 0};
                for (int k = 0; k < valid; k++) w[k >> 2] |= (uint32_t)src[k] << (8 * (k & 3));
                my_bits = lookup_items<WIDE, false, entry_t, kItems>(s_enc, w, valid, item_val, item_len);
            }

            // ---- B. exclusive scan of bit totals over the warp
            uint32_t incl = my_bits;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += x;
            }
            const uint32_t tile_bits = __shfl_sync(0xFFFFFFFFu, incl, 31);
            // every lane holds >= 32 bits: each staging word is then shared by at most two NEIGHBOURING lanes
            const bool fast = __all_sync(0xFFFFFFFFu, my_bits >= 32u);

            // ---- C. stream the codes into the staging buffer (own bits start at staging bit 128)
            __syncwarp();  // the previous sub-tile's copy-out has read the staging buffer
            if (lane < 4) stage[lane] = carried;
            const uint32_t pos = 128u + incl - my_bits;
            uint32_t wi = pos >> 5, nb = pos & 31, hi = 0, lo = 0;
            if (fast) {
                // plain stores: a lane writes every word it completes; the leading `nb` bits of its first word
                // belong to its left neighbour, whose trailing partial word arrives by shuffle afterwards
                const uint32_t first_wi = wi, lead = nb;
#pragma unroll
                for (int k = 0; k < kItems; k++) emit_bits_owned(stage, hi, lo, nb, wi, item_val[k], item_len[k]);
                const uint32_t my_tail = nb ? lo << (32u - nb) : 0u;  // trailing partial word, left-aligned
                const uint32_t left_tail = __shfl_up_sync(0xFFFFFFFFu, my_tail, 1);
                if (lane != 0 && lead != 0) stage[first_wi] |= left_tail;
                if (lane == 31) {  // zero padding behind the last bit: the copy-out reads up to 160 bits past it
                    stage[wi] = my_tail;
#pragma unroll
                    for (int k = 1; k <= 5; k++) stage[wi + k] = 0;
                }
            } else {
                for (int i = 4 + lane; i < kStageWords; i += 32) stage[i] = 0;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < kItems; k++) emit_bits(stage, hi, lo, nb, wi, item_val[k], item_len[k]);
                stage_or_if(stage + wi, lo << ((32u - nb) & 31u), nb != 0u);
            }
            __syncwarp();

            // ---- D. copy-out at the global alignment
            const unsigned long long gend = g + tile_bits;
            const unsigned long long v0 = g >> 7, v1 = gend >> 7;  // 16-byte words [v0, v1) end inside this sub-tile
            const uint32_t r = (uint32_t)(g & 127);
            const bool last_sub = t == nsub - 1;
            const bool stream_end = last_chunk && last_sub;
            const bool shared_first = t == 0 && chunk != 0 && r != 0;  // first word also holds the previous chunk's bits
            {
                const uint32_t nvec = (uint32_t)(v1 - v0);
                uint4 *dst = (uint4 *)out + v0;
                for (uint32_t j = (shared_first ? 1u : 0u) + lane; j < nvec; j += 32) {
                    const uint32_t sbit = j * 128u + (128u - r);
                    uint4 o;
                    o.x = bswap32(stage_word(stage, sbit));
                    o.y = bswap32(stage_word(stage, sbit + 32));
                    o.z = bswap32(stage_word(stage, sbit + 64));
                    o.w = bswap32(stage_word(stage, sbit + 96));
                    stg_stream(dst + j, o);
                }
            }
            if (shared_first && lane == 8) {  // (v1 == v0 only for a tiny final chunk: the word is shared AND last)
                const uint32_t sbit = 128u - r;
                const uint4 m = make_uint4(stage_word(stage, sbit), stage_word(stage, sbit + 32),
                                           stage_word(stage, sbit + 64), stage_word(stage, sbit + 96));
                boundary_merge(ws, chunk, false, m, out, v0, stream_bytes);
            }
            if (last_sub && (gend & 127) != 0 && !(shared_first && v1 == v0)) {
                const uint32_t sbit = (uint32_t)(v1 - v0) * 128u + (128u - r);
                if (stream_end) {  // trailing partial 16-byte word of the stream, byte by byte (zero padded)
                    const uint32_t rem_bytes = (uint32_t)(((gend & 127) + 7) >> 3);
                    if (lane < (int)rem_bytes) out[v1 * 16 + lane] = (uint8_t)(stage_word(stage, sbit + 8u * lane) >> 24);
                } else if (lane == 16) {  // the word shared with the next chunk
                    const uint4 m = make_uint4(stage_word(stage, sbit), stage_word(stage, sbit + 32),
                                               stage_word(stage, sbit + 64), stage_word(stage, sbit + 96));
                    boundary_merge(ws, chunk + 1, true, m, out, v1, stream_bytes);
                }
            }
            if (lane < 4) carried = stage_word(stage, tile_bits + 32u * lane);  // last 128 bits of this sub-tile
            g = gend;
        }
    }
}


// ================================================================================================ shared pieces of the single-pass kernels
constexpr int kSpZeroPrefix = 4;                                            // words of zeros in front of the data

struct SpWorkspace {
    uint32_t *bstate;              // [nchunks+1]  arrivals at the boundary word between chunk b-1 and chunk b
    uint4 *bleft, *bright;         // [nchunks+1]
};

// The 16-byte word shared by the last chunk of one run and the first chunk of the next: both sides deposit
// their half (big-endian word domain) and bump a counter; the second arrival stores the word.  `ends` != 0
// (right side only): the stream ends inside this word and only that many of its bytes exist.
__device__ __forceinline__ void sp_boundary_merge(const SpWorkspace &ws, unsigned int b, bool left_side, uint4 mine, uint32_t ends,
                                                  uint8_t *__restrict__ out, unsigned long long vec, size_t out_cap) {
    uint4 *my_slot = left_side ? ws.bleft + b : ws.bright + b;
    const uint4 *other_slot = left_side ? ws.bright + b : ws.bleft + b;
    *my_slot = mine;
    __threadfence();
    const uint32_t old = atomicAdd(&ws.bstate[b], 1u + (ends << 8));
    if ((old & 0xFFu) == 1u) {
        __threadfence();
        const uint4 o = ld_cg_u128(other_slot);
        uint4 rr;
        rr.x = bswap32(mine.x | o.x);
        rr.y = bswap32(mine.y | o.y);
        rr.z = bswap32(mine.z | o.z);
        rr.w = bswap32(mine.w | o.w);
        const uint32_t nbytes = (old >> 8) | ends;  // 0 = the whole word
        if (nbytes == 0 && (vec + 1) * 16 <= out_cap) {
            stg_stream((uint4 *)out + vec, rr);
        } else {
            const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
            const size_t end = min((size_t)out_cap, (size_t)vec * 16 + (nbytes ? nbytes : 16u));
            for (size_t i = vec * 16; i < end; i++) out[i] = (uint8_t)(rw[(i & 15) >> 2] >> (8 * (i & 3)));
        }
    }
}

// ================================================================================================ planned single pass
// Tables whose longest code has at most 12 bits.  The bit offset of every 32 KB run is known BEFORE the launch (E1 + E2
// from the input, or the run histograms K1 left behind: dc_histogram_u8_runs + encode_plan_kernel), so a run needs
// nothing from any other CTA: no tickets, no descriptors, no spinning.  What is left is one barrier among the 16 warps
// of a run (2 KB chunks) to add up their bit counts.
//
// The inner loop is built around the two pipes that bound the first version (ncu, round 1: integer ALU pipe 63 %, one
// LDS per symbol at 2.7 wavefronts):
//   * lane-private table: row b (256 bytes) holds byte b's entry once per lane, lane L at +4L, so a look-up never has a
//     bank conflict and its address is ONE instruction: PRMT(input word, 4 * lane) = byte << 8 | 4 * lane.
//   * entry = code LEFT-aligned | 1 << 7 | length.  A pair is combined with one shift whose amount is the first
//     entry's low 5 bits, taken straight from the entry, and one LOP3: V = (e0 | e1 >> l0) & ~0xFF (<= 24 bits).
//     The running position P += e0 + e1 is exact in its low 7 bits (they never see the code bits above), which is all
//     the shifts (P mod 32) and the word-completion test (bit 5 toggles) need; the markers count the symbols that
//     had a code in bits 7..10 of the same sums.
//   * emit: cur |= V >> P; next = V << (32 - P) (funnel shift of V:0); when bit 5 of P toggles the word is complete:
//     predicated STS, cur = next, wi += 4.
// Two instantiations.  PAIR: codes up to 12 bits, two symbols per emit step, 16 warps of 2 KB per run and two runs in flight
// per CTA.  Single symbols: codes of 13 .. 16 bits (a pair no longer fits the 24 bits above the marker), 32 warps of 1 KB
// per run -- the staging buffer of a chunk must hold 16 bits per symbol and all of them must lie in front of the table.
template <int MAXBITS>
struct FwCfg {
    static constexpr bool kPair = MAXBITS <= kPlannedMaxBits;
    static constexpr int kWarps = kPair ? 16 : 32;                 // warps (= chunks) per run
    static constexpr int kGroups = 32 / kWarps;                    // runs in flight per CTA (they share the table)
    static constexpr int kThreads = 1024;
    static constexpr int kChunkBytes = kRunBytes / kWarps;         // 2 KB / 1 KB
    static constexpr int kChunkSubs = kChunkBytes / kSubTile;      // 4 / 2
    static constexpr int kStageWords = kSpZeroPrefix + kChunkBytes * MAXBITS / 32 + 12;   // per warp
    static constexpr size_t kStageBytes = (size_t)32 * kStageWords * 4;
    static constexpr size_t kXchgBytes = (size_t)kGroups * 2 * kWarps * 8 * 4;            // per group: 2 copies x warps x 8 words
};
constexpr int kFwTableBytes = 256 * 256;
// The table sits at the ABSOLUTE shared-memory address 0x20000, so that the address of a look-up is one PRMT and nothing
// else: byte 0 = 4 * lane, byte 1 = the input byte, byte 2 = 0x02 (from the lane constant).  The staging buffers lie in
// front of it; the kernel computes the padding from its own window base (1 KB of system-reserved memory on sm_100).
constexpr uint32_t kFwTableAddr = 0x20000u;
constexpr size_t kFwSmemBytes = kFwTableAddr + kFwTableBytes - 1024;                  // dynamic bytes when the window starts at 1 KB

__device__ __forceinline__ uint32_t fw_prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// the entry of byte K of w for this lane (lanec = 0x20000 | 4 * lane): one PRMT builds the whole shared-memory address
template <int K>
__device__ __forceinline__ uint32_t fw_entry(uint32_t w, uint32_t lanec) {
    uint32_t e;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(fw_prmt(w, lanec, 0x7604u | ((uint32_t)K << 4))));
    return e;
}

struct FwWorkspace {
    const unsigned long long *run_off;   // [nruns + 1] exclusive bit offsets, last = total
    uint32_t *bstate;
    uint4 *bleft, *bright;
    const uint32_t *d_phase;
};

// one sub-tile (32 lanes x 16 symbols) appended to the warp's staging buffer at bit position `bitpos`
template <bool FULL>
__device__ __forceinline__ void fw_subtile(uint32_t *stage, const uint32_t (&w)[4], int valid, int lane, uint32_t lanec,
                                           uint32_t &bitpos, uint32_t &bad) {
    uint32_t V[8], S[8];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t e0 = fw_entry<0>(w[q], lanec), e1 = fw_entry<1>(w[q], lanec);
        uint32_t e2 = fw_entry<2>(w[q], lanec), e3 = fw_entry<3>(w[q], lanec);
        if (!FULL) {   // positions behind the end of the input: no bits, but counted as "has a code"
            if (4 * q + 0 >= valid) e0 = 0x80u;
            if (4 * q + 1 >= valid) e1 = 0x80u;
            if (4 * q + 2 >= valid) e2 = 0x80u;
            if (4 * q + 3 >= valid) e3 = 0x80u;
        }
        V[2 * q] = (e0 | __funnelshift_r(e1, 0u, e0)) & 0xFFFFFF00u;        // e1 >> (e0 & 31)
        V[2 * q + 1] = (e2 | __funnelshift_r(e3, 0u, e2)) & 0xFFFFFF00u;
        S[2 * q] = e0 + e1;
        S[2 * q + 1] = e2 + e3;
    }
    const uint32_t sa = S[0] + S[1] + S[2] + S[3], sb = S[4] + S[5] + S[6] + S[7];
    bad |= ~(sa & sb);                         // bit 10 stays clear only if all 8 + 8 symbols had a code
    const uint32_t my_bits = (sa & 127u) + (sb & 127u);
    uint32_t incl = my_bits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += x;
    }
    const uint32_t tile_bits = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const bool fast = __all_sync(0xFFFFFFFFu, my_bits >= 32u);
    const uint32_t pos = bitpos + incl - my_bits;
    if (fast) {
        // plain stores: a lane writes every word it completes.  The leading `lead` bits of its first word are its left
        // neighbour's trailing bits (lane 0: the previous sub-tile's, already in the buffer)
        const uint32_t lead = pos & 31u;
        uint32_t wi = (uint32_t)__cvta_generic_to_shared(stage + (pos >> 5));
        const uint32_t first_wi = wi;
        uint32_t left_tail = 0;
        if (lane == 0 && lead != 0) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(left_tail) : "r"(first_wi) : "memory");
        uint32_t cur = 0, P = pos;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile(
                "{\n\t.reg .pred q;\n\t.reg .u32 h, nx, p2, x;\n\t"
                "shf.r.wrap.b32 h, %4, 0, %2;\n\t"        // V >> (P & 31)
                "or.b32 %1, %1, h;\n\t"
                "shf.r.wrap.b32 nx, 0, %4, %2;\n\t"       // V << (32 - (P & 31)), 0 when P & 31 == 0
                "add.u32 p2, %2, %3;\n\t"
                "xor.b32 x, p2, %2;\n\t"
                "and.b32 x, x, 32;\n\t"
                "setp.ne.u32 q, x, 0;\n\t"
                "@q st.shared.u32 [%0], %1;\n\t"
                "@q mov.u32 %1, nx;\n\t"
                "@q add.u32 %0, %0, 4;\n\t"
                "mov.u32 %2, p2;\n\t}"
                : "+r"(wi), "+r"(cur), "+r"(P)
                : "r"(S[k]), "r"(V[k])
                : "memory");
        }
        const uint32_t lt = __shfl_up_sync(0xFFFFFFFFu, cur, 1);   // cur = the trailing partial word, left-aligned
        if (lane != 0) left_tail = lt;
        if (lead != 0) {
            uint32_t t;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(first_wi) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(first_wi), "r"(t | left_tail) : "memory");
        }
        if (lane == 31) asm volatile("st.shared.u32 [%0], %1;" ::"r"(wi), "r"(cur) : "memory");
    } else {
        // very short codes or a ragged tile: shared-memory OR on zeroed words (emit_bits)
        const uint32_t z0 = (bitpos >> 5) + 1, z1 = ((bitpos + tile_bits) >> 5) + 1;
        for (uint32_t i = z0 + lane; i <= z1; i += 32) stage[i] = 0;
        __syncwarp();
        uint32_t wi = pos >> 5, nb = pos & 31, hi = 0, lo = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t len = S[k] & 127u;
            emit_bits(stage, hi, lo, nb, wi, len ? V[k] >> (32u - len) : 0u, len);
        }
        stage_or_if(stage + wi, lo << ((32u - nb) & 31u), nb != 0u);
    }
    __syncwarp();
    bitpos += tile_bits;
}

// the same for tables with codes of 13 .. 16 bits: entry = code left-aligned | 1 << 9 | length, one symbol per emit step.
// The low 9 bits of the sum of the 16 entries are the lane's bit count (<= 256), bits 9..13 count the symbols that had a code.
template <bool FULL>
__device__ __forceinline__ void fw_subtile_single(uint32_t *stage, const uint32_t (&w)[4], int valid, int lane, uint32_t lanec,
                                                  uint32_t &bitpos, uint32_t &bad) {
    uint32_t E[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        E[4 * q + 0] = fw_entry<0>(w[q], lanec);
        E[4 * q + 1] = fw_entry<1>(w[q], lanec);
        E[4 * q + 2] = fw_entry<2>(w[q], lanec);
        E[4 * q + 3] = fw_entry<3>(w[q], lanec);
    }
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (!FULL && k >= valid) E[k] = 0x200u;   // behind the end of the input: no bits, counted as "has a code"
        sum += E[k];
    }
    bad |= ~sum >> 3;                             // bit 13 of the sum (16 symbols with a code) lands on bit 10 of `bad`
    const uint32_t my_bits = sum & 511u;
    uint32_t incl = my_bits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += x;
    }
    const uint32_t tile_bits = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const bool fast = __all_sync(0xFFFFFFFFu, my_bits >= 32u);
    const uint32_t pos = bitpos + incl - my_bits;
    if (fast) {
        const uint32_t lead = pos & 31u;
        uint32_t wi = (uint32_t)__cvta_generic_to_shared(stage + (pos >> 5));
        const uint32_t first_wi = wi;
        uint32_t left_tail = 0;
        if (lane == 0 && lead != 0) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(left_tail) : "r"(first_wi) : "memory");
        uint32_t cur = 0, P = pos;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            asm volatile(
                "{\n\t.reg .pred q;\n\t.reg .u32 v, h, nx, p2, x;\n\t"
                "and.b32 v, %3, 0xFFFF0000;\n\t"
                "shf.r.wrap.b32 h, v, 0, %2;\n\t"
                "or.b32 %1, %1, h;\n\t"
                "shf.r.wrap.b32 nx, 0, v, %2;\n\t"
                "add.u32 p2, %2, %3;\n\t"
                "xor.b32 x, p2, %2;\n\t"
                "and.b32 x, x, 32;\n\t"
                "setp.ne.u32 q, x, 0;\n\t"
                "@q st.shared.u32 [%0], %1;\n\t"
                "@q mov.u32 %1, nx;\n\t"
                "@q add.u32 %0, %0, 4;\n\t"
                "mov.u32 %2, p2;\n\t}"
                : "+r"(wi), "+r"(cur), "+r"(P)
                : "r"(E[k])
                : "memory");
        }
        const uint32_t lt = __shfl_up_sync(0xFFFFFFFFu, cur, 1);
        if (lane != 0) left_tail = lt;
        if (lead != 0) {
            uint32_t t;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(first_wi) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(first_wi), "r"(t | left_tail) : "memory");
        }
        if (lane == 31) asm volatile("st.shared.u32 [%0], %1;" ::"r"(wi), "r"(cur) : "memory");
    } else {
        const uint32_t z0 = (bitpos >> 5) + 1, z1 = ((bitpos + tile_bits) >> 5) + 1;
        for (uint32_t i = z0 + lane; i <= z1; i += 32) stage[i] = 0;
        __syncwarp();
        uint32_t wi = pos >> 5, nb = pos & 31, hi = 0, lo = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t len = E[k] & 31u;
            emit_bits(stage, hi, lo, nb, wi, len ? E[k] >> (32u - len) : 0u, len);
        }
        stage_or_if(stage + wi, lo << ((32u - nb) & 31u), nb != 0u);
    }
    __syncwarp();
    bitpos += tile_bits;
}

template <int THREADS>
__device__ __forceinline__ void fw_group_barrier(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(THREADS) : "memory");
}

template <int MAXBITS>
__device__ __forceinline__ void encode_fast_body(const uint8_t *__restrict__ in, size_t n,
                                                                    const dc_huff_table *__restrict__ tab, uint8_t *__restrict__ out,
                                                                    size_t out_cap, unsigned phase, FwWorkspace ws, unsigned int nruns,
                                                                    int32_t *__restrict__ d_status, uint32_t smem_bytes, bool exact) {
    typedef FwCfg<MAXBITS> Cfg;
    constexpr int kFwWarps = Cfg::kWarps, kFwGroups = Cfg::kGroups, kFwThreads = Cfg::kThreads, kFwChunkBytes = Cfg::kChunkBytes;
    constexpr int kFwChunkSubs = Cfg::kChunkSubs, kFwStageWords = Cfg::kStageWords;
    constexpr size_t kFwStageBytes = Cfg::kStageBytes, kFwXchgBytes = Cfg::kXchgBytes;
    extern __shared__ __align__(16) uint8_t fw_smem[];   // [exchange | staging buffers | padding | table at 0x20000]
    if (!table_usable(tab, d_status)) return;
    {   // this instantiation's tables: [1, 12] (pairs) or (12, 16] (single symbols); longer codes take the 64-bit-entry kernel
        const int mb = tab->max_bits;
        if (mb > MAXBITS || (!Cfg::kPair && mb <= kPlannedMaxBits)) {
            // launched alone because the host's copy of the table header said so: that header was not this table's
            if (exact && blockIdx.x == 0 && threadIdx.x == 0) set_status(d_status, DC_ERR_ARG);
            return;
        }
    }
    if (ws.d_phase) phase = *ws.d_phase & 7u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, group = warp / kFwWarps, gw = warp % kFwWarps;
    const size_t stream_bytes = (size_t)(((unsigned long long)phase + ws.run_off[nruns] + 7) >> 3);
    if (stream_bytes > out_cap) {
        if (blockIdx.x == 0 && tid == 0) set_status(d_status, DC_ERR_CAPACITY);
        return;
    }
    const uint32_t window = (uint32_t)__cvta_generic_to_shared(fw_smem);
    if (window + kFwXchgBytes + kFwStageBytes > kFwTableAddr || kFwTableAddr + kFwTableBytes > window + smem_bytes) {
        if (blockIdx.x == 0 && tid == 0) set_status(d_status, DC_ERR_CUDA);   // the window does not start where sm_100 puts it
        return;
    }
    uint32_t *s_tab = (uint32_t *)(fw_smem + (kFwTableAddr - window));
    for (int i = tid; i < 256 * 32; i += kFwThreads) {
        const int b = i >> 5;
        const uint32_t e = tab->enc[b], l = e & 63u;
        s_tab[b * 64 + (i & 31)] = l ? (((e >> 6) << (32u - l)) | (Cfg::kPair ? 0x80u : 0x200u) | l) : 0u;
    }
    __syncthreads();
    // what a warp shows its run: its bit count and the first 128 bits of its chunk (the left neighbour completes the 16-byte
    // word the two share).  Two copies, used alternately, so that ONE barrier per run is enough: a warp that is already
    // staging the next run writes the other copy.
    uint32_t *xchg_base = (uint32_t *)fw_smem + group * (2 * kFwWarps * 8);
    uint32_t *stage = (uint32_t *)(fw_smem + kFwXchgBytes) + warp * kFwStageWords;
    const uint32_t lanec = kFwTableAddr | (4u * lane);
    const size_t nchunks = (n + kFwChunkBytes - 1) / kFwChunkBytes;
    const unsigned int stride = gridDim.x * kFwGroups;
    uint32_t bad = 0, parity = 0;
    uint4 v[kFwChunkSubs];
    unsigned int run = blockIdx.x * kFwGroups + group;
    auto full_chunk = [&](unsigned int r) { return r < nruns && ((size_t)r * kFwWarps + gw + 1) * kFwChunkBytes <= n; };
    auto load_chunk = [&](unsigned int r) {
        const uint4 *src = (const uint4 *)(in + ((size_t)r * kFwWarps + gw) * kFwChunkBytes) + lane;
#pragma unroll
        for (int t = 0; t < kFwChunkSubs; t++) v[t] = ldg_stream(src + t * 32);
    };
    if (full_chunk(run)) load_chunk(run);
    for (; run < nruns; run += stride, parity ^= 1u) {
        const size_t chunk = (size_t)run * kFwWarps + gw;
        const bool have_chunk = chunk < nchunks;
        const size_t chunk_base = chunk * kFwChunkBytes;
        const size_t chunk_len = have_chunk ? min((size_t)kFwChunkBytes, n - chunk_base) : 0;
        const bool last_chunk = chunk == nchunks - 1;
        const unsigned long long run_excl = ws.run_off[run];   // (needed behind the barrier: in flight until then)

        // ---- 1. encode the chunk into the staging buffer (chunk-local bit offsets behind a zero prefix)
        __syncwarp();
        if (lane <= kSpZeroPrefix) stage[lane] = 0;
        uint32_t bitpos = 32u * kSpZeroPrefix;
        __syncwarp();
        if (chunk_len == (size_t)kFwChunkBytes) {
#pragma unroll
            for (int t = 0; t < kFwChunkSubs; t++) {
                const uint32_t w[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
                if (Cfg::kPair) fw_subtile<true>(stage, w, kEncPerThread, lane, lanec, bitpos, bad);
                else fw_subtile_single<true>(stage, w, kEncPerThread, lane, lanec, bitpos, bad);
            }
        } else {  // the ragged last chunk of the stream (or no chunk at all)
            const int nsub = (int)((chunk_len + kSubTile - 1) / kSubTile);
#pragma unroll 1
            for (int t = 0; t < nsub; t++) {
                const size_t base = chunk_base + (size_t)t * kSubTile + (size_t)lane * kEncPerThread;
                const int valid = base < n ? (int)min((size_t)kEncPerThread, n - base) : 0;
                uint32_t w[4] = {0, 0, 0, 0};
                for (int k = 0; k < valid; k++) w[k >> 2] |= (uint32_t)in[base + k] << (8 * (k & 3));
                if (Cfg::kPair) fw_subtile<false>(stage, w, valid, lane, lanec, bitpos, bad);
                else fw_subtile_single<false>(stage, w, valid, lane, lanec, bitpos, bad);
            }
        }
        // the next run's input is on its way while this one waits for its neighbours and is copied out
        if (full_chunk(run + stride)) load_chunk(run + stride);
        if (lane < 8) stage[(bitpos >> 5) + 1 + lane] = 0;  // the copy-out reads up to 5 words past the last bit
        const uint32_t chunk_bits = bitpos - 32u * kSpZeroPrefix;
        __syncwarp();
        uint32_t *xchg = xchg_base + parity * (kFwWarps * 8);
        if (lane < 5) xchg[gw * 8 + lane] = lane == 0 ? chunk_bits : stage[kSpZeroPrefix - 1 + lane];
        fw_group_barrier<kFwWarps * 32>(group);

        // ---- 2. the chunk's place in the stream: the run's planned offset + the chunks in front of it
        unsigned long long excl = run_excl;
        {
            uint32_t mine = lane < gw ? xchg[lane * 8] : 0u;   // kFwWarps <= 32
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
            excl += __shfl_sync(0xFFFFFFFFu, mine, 0);
        }

        // ---- 3. copy-out at the global alignment
        if (have_chunk) {
            const unsigned long long g = (unsigned long long)phase + excl, gend = g + chunk_bits;
            const size_t stream_bytes_here = (size_t)((gend + 7) >> 3);
            const unsigned long long v0 = g >> 7, v1 = gend >> 7;  // 16-byte words [v0, v1) end inside this chunk
            const uint32_t r = (uint32_t)(g & 127), t = (uint32_t)(gend & 127);
            const bool shared_first = chunk != 0 && r != 0;           // first word also holds the previous chunk's bits
            if (!(v1 == v0 && !last_chunk)) {   // (< 128 bits from 2048 symbols: symbols without codes, reported below)
                const uint32_t nvec = (uint32_t)(v1 - v0);
                const uint32_t sbit0 = 32u * kSpZeroPrefix - r;            // staging bit of the first bit of 16-byte word v0
                const uint32_t sbit1 = nvec * 128u + sbit0;                // ... of 16-byte word v1
                SpWorkspace sp;
                sp.bstate = ws.bstate;
                sp.bleft = ws.bleft;
                sp.bright = ws.bright;
                // 3a. the 16-byte words shared with the neighbouring chunks: inside the run the LEFT chunk owns the word (it
                // takes the right chunk's leading bits from what that warp published); across runs both sides deposit their
                // half and the second arrival stores the word
                if (shared_first && gw == 0 && lane == 8) {
                    const uint4 m = make_uint4(stage_word(stage, sbit0), stage_word(stage, sbit0 + 32), stage_word(stage, sbit0 + 64),
                                               stage_word(stage, sbit0 + 96));
                    const uint32_t ends = (last_chunk && v1 == v0) ? (uint32_t)(stream_bytes_here - v0 * 16) : 0u;
                    sp_boundary_merge(sp, run, false, m, ends, out, v0, out_cap);   // boundary `run`: between run - 1 and run
                }
                if (t != 0 && !(shared_first && v1 == v0)) {
                    if (last_chunk) {  // trailing partial 16-byte word of the stream, byte by byte (zero padded)
                        const uint32_t rem_bytes = (t + 7) >> 3;
                        if (lane < (int)rem_bytes) out[v1 * 16 + lane] = (uint8_t)(stage_word(stage, sbit1 + 8u * lane) >> 24);
                    } else if (gw == kFwWarps - 1) {
                        if (lane == 16) {
                            const uint4 m = make_uint4(stage_word(stage, sbit1), stage_word(stage, sbit1 + 32), stage_word(stage, sbit1 + 64),
                                                       stage_word(stage, sbit1 + 96));
                            sp_boundary_merge(sp, run + 1, true, m, 0u, out, v1, out_cap);
                        }
                    } else if (lane < 16) {
                        // own trailing t bits | the right chunk's first 128 - t bits: bits [128 - t, 256 - t) of (128 zero bits, its
                        // first 128 bits), i.e. its published words shifted right by t
                        const uint32_t *rw = xchg + (gw + 1) * 8;   // [bits, w0, w1, w2, w3]
                        const uint32_t q = (128u - t) + 32u * (lane >> 2), a = q >> 5, sft = q & 31u;   // window word a, a + 1 of (0,0,0,0,w0..w3)
                        const uint32_t hi = a >= 4 ? rw[a - 3] : 0u, lo = a + 1 >= 4 ? (a + 1 < 8 ? rw[a - 2] : 0u) : 0u;
                        const uint32_t word = stage_word(stage, sbit1 + 32u * (lane >> 2)) | __funnelshift_l(lo, hi, sft);
                        const unsigned long long rend = gend + rw[0];
                        size_t limit = out_cap;
                        if (chunk + 1 == nchunks - 1) limit = min(limit, (size_t)((rend + 7) >> 3));
                        const size_t byte = (size_t)v1 * 16 + lane;
                        if (byte < limit) out[byte] = (uint8_t)(word >> (24 - 8 * (lane & 3)));
                    }
                }
                // 3b. the words that are this chunk's alone.  Output vector j needs staging words s0 + 4 j .. s0 + 4 j + 4 with
                // s0 = sbit0 / 32 = 4 q + d: two ALIGNED 16-byte loads (vectors q + j and q + j + 1: consecutive lanes read
                // consecutive vectors, no bank conflict) and a warp-uniform choice of the five words by d -- five scalar
                // loads at a stride of four words would be 4-way conflicts each
                uint4 *dst = (uint4 *)out + v0;
                const uint32_t sh = sbit0 & 31, s0 = sbit0 >> 5, d = s0 & 3u;
                const uint4 *sv = (const uint4 *)stage + (s0 >> 2);
                for (uint32_t j = (shared_first ? 1u : 0u) + lane; j < nvec; j += 32) {
                    const uint4 A = sv[j], B = sv[j + 1];
                    uint32_t a0, a1, a2, a3, a4;
                    if (d == 0) { a0 = A.x; a1 = A.y; a2 = A.z; a3 = A.w; a4 = B.x; }
                    else if (d == 1) { a0 = A.y; a1 = A.z; a2 = A.w; a3 = B.x; a4 = B.y; }
                    else if (d == 2) { a0 = A.z; a1 = A.w; a2 = B.x; a3 = B.y; a4 = B.z; }
                    else { a0 = A.w; a1 = B.x; a2 = B.y; a3 = B.z; a4 = B.w; }
                    uint4 o;
                    o.x = bswap32(__funnelshift_l(a1, a0, sh));
                    o.y = bswap32(__funnelshift_l(a2, a1, sh));
                    o.z = bswap32(__funnelshift_l(a3, a2, sh));
                    o.w = bswap32(__funnelshift_l(a4, a3, sh));
                    stg_stream(dst + j, o);
                }
            }
        }
    }
    if (((bad >> 10) & 1u) != 0u) set_status(d_status, DC_ERR_SYMBOL);
}

// WHICH = 12 / 16: the host knows the table's longest code (table_meta_fetch) and launches the one instantiation that takes it.
// WHICH = 0: it does not yet (the build is still queued): one launch that picks by the table it finds.
template <int WHICH>
__global__ void __launch_bounds__(1024, 1) encode_fast_kernel(const uint8_t *__restrict__ in, size_t n,
                                                              const dc_huff_table *__restrict__ tab, uint8_t *__restrict__ out,
                                                              size_t out_cap, unsigned phase, FwWorkspace ws, unsigned int nruns,
                                                              int32_t *__restrict__ d_status, uint32_t smem_bytes) {
    if (WHICH == kPlannedMaxBits) {
        encode_fast_body<kPlannedMaxBits>(in, n, tab, out, out_cap, phase, ws, nruns, d_status, smem_bytes, true);
    } else if (WHICH == kNarrowBits) {
        encode_fast_body<kNarrowBits>(in, n, tab, out, out_cap, phase, ws, nruns, d_status, smem_bytes, true);
    } else {
        if (tab->max_bits <= kPlannedMaxBits) encode_fast_body<kPlannedMaxBits>(in, n, tab, out, out_cap, phase, ws, nruns, d_status, smem_bytes, false);
        else encode_fast_body<kNarrowBits>(in, n, tab, out, out_cap, phase, ws, nruns, d_status, smem_bytes, false);
    }
}

// ------------------------------------------------------------------------------------------ plan (run offsets from run histograms)
// bits of run r = sum over symbols of run_hist[r][s] * bits(s): one warp per run (a lane takes 8 of the 256 u16 counts with one
// 16-byte load); the CTA that finishes last scans the run totals into exclusive offsets (E2's arithmetic) -- one launch.
constexpr int kPlanThreads = 1024;
__global__ void __launch_bounds__(kPlanThreads) encode_plan_kernel(const uint16_t *__restrict__ run_hist, const dc_huff_table *__restrict__ tab,
                                                                   EncWorkspace ws, unsigned int nruns, ChainSlots slots,
                                                                   unsigned long long *__restrict__ d_total_bits, int32_t *__restrict__ d_status) {
    // CTA b takes the contiguous runs [lo, hi): a warp per run for the bit counts, a scan inside the CTA, CTA totals from CTA
    // to CTA (chain_scan.cuh)
    __shared__ uint32_t s_len[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ok = table_usable(tab, d_status);
    if (tid < 256) s_len[tid] = ok ? (uint32_t)(tab->enc64[tid] >> 32) : 0u;
    __syncthreads();
    const unsigned int per = (nruns + gridDim.x - 1) / gridDim.x, lo = min(nruns, per * blockIdx.x), hi = min(nruns, lo + per);
    uint32_t *s_run = ws.run_bits + lo;   // bits of this CTA's runs (global, L2-resident: written and read inside the CTA)
    bool missing = false;
    for (unsigned int run = lo + warp; run < hi; run += kPlanThreads / 32) {
        const uint4 v = *((const uint4 *)(run_hist + (size_t)run * 256) + lane);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t c = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu, l = s_len[8 * lane + k];
            missing |= c != 0 && l == 0;
            sum += c * l;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if (lane == 0) s_run[run - lo] = sum;
    }
    if (missing && ok) set_status(d_status, DC_ERR_SYMBOL);
    __syncthreads();
    unsigned long long mine = 0;
    for (unsigned int i = tid; i < hi - lo; i += kPlanThreads) mine += s_run[i];
    unsigned long long cta_total;
    block_exclusive(mine, &cta_total);
    unsigned long long carry = chain_exclusive(slots, cta_total);
    for (unsigned int base = 0; base < hi - lo; base += kPlanThreads) {
        const unsigned int i = base + tid;
        const unsigned long long c = i < hi - lo ? s_run[i] : 0ull;
        unsigned long long trip_total;
        const unsigned long long off = block_exclusive(c, &trip_total);
        if (i < hi - lo) ws.run_off[lo + i] = carry + off;
        carry += trip_total;
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        ws.run_off[nruns] = carry;
        if (d_total_bits) *d_total_bits = carry;
    }
}

static size_t enc_ws_layout(size_t n, size_t off[10]) {
    const size_t nruns = (n + kRunBytes - 1) / kRunBytes, nchunks = (n + kChunkBytes - 1) / kChunkBytes;
    size_t p = 64;
    auto take = [&](size_t bytes) { size_t o = p; p += (bytes + 63) & ~(size_t)63; return o; };
    size_t o[10];
    o[0] = take(nruns * 4);              // run_bits
    o[1] = take(nruns * kEncWarps * 4);  // chunk_rel                 (64-bit-entry path)
    o[2] = take((nruns + 1) * 8);        // run_off
    o[3] = take((nchunks + 1) * 4 + 64); // bstate: arrival counters of the 16-byte words that two runs (64-bit-entry path: chunks) share
    o[6] = o[7] = 0;                     // (slots of the look-back encoder of round 1: gone)
    o[4] = take((nchunks + 1) * 16);     // bleft
    o[5] = take((nchunks + 1) * 16);     // bright
    o[8] = take(nruns * 512);            // run histograms (dc_histogram_u8_runs -> dc_huff_encode_planned)
    o[9] = take(kChainSlotsBytes);       // the plan kernel's CTA totals and flags (chain_scan.cuh)
    if (off) for (int i = 0; i < 10; i++) off[i] = o[i];
    return p;
}

}  // namespace dc

using namespace dc;

static int launch_encode_body(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity,
                              unsigned bit_phase, EncWorkspace ws, char *w, const size_t *off, unsigned int nruns,
                              unsigned long long *d_total_bits, int32_t *d_status, cudaStream_t st, bool planned);

extern "C" size_t dc_huff_encode_workspace_bytes(size_t n) { return enc_ws_layout(n, nullptr); }

static int encode_entry(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity,
                        unsigned bit_phase, uint64_t *d_total_bits, int32_t *d_status, void *d_workspace, size_t workspace_bytes,
                        void *stream, bool planned, const uint32_t *d_phase = nullptr) {
    if (!d_table || (n && (!d_in || !d_out || !d_workspace)) || bit_phase > 7) return DC_ERR_ARG;
    if ((((uintptr_t)d_in | (uintptr_t)d_out | (uintptr_t)d_workspace) & 15) != 0) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (d_total_bits) DC_CUDA_TRY(cudaMemsetAsync(d_total_bits, 0, sizeof(uint64_t), st));
    if (n == 0) return DC_OK;
    size_t off[10];
    const size_t need = enc_ws_layout(n, off);
    if (workspace_bytes < need) return DC_ERR_CAPACITY;
    const size_t nruns64 = (n + kRunBytes - 1) / kRunBytes, nchunks = (n + kChunkBytes - 1) / kChunkBytes;
    if (nruns64 > 0x0FFFFFF0ull) return DC_ERR_ARG;
    const unsigned int nruns = (unsigned int)nruns64;
    char *w = (char *)d_workspace;
    EncWorkspace ws;
    ws.run_bits = (uint32_t *)(w + off[0]);
    ws.chunk_rel = (uint32_t *)(w + off[1]);
    ws.run_off = (unsigned long long *)(w + off[2]);
    ws.bstate = (uint32_t *)(w + off[3]);
    ws.bleft = (uint4 *)(w + off[4]);
    ws.bright = (uint4 *)(w + off[5]);
    ws.d_phase = d_phase;
    // the arrival counters of the 16-byte words that two runs share
    DC_CUDA_TRY(cudaMemsetAsync(w + off[3], 0, (nchunks + 1) * 4 + 64, st));
    const unsigned int sms = (unsigned int)sm_count();
    if (planned) {
        // the bit offset of every run from the run histograms K1 left in this workspace (dc_histogram_u8_runs)
        {
            ChainSlots slots;
            slots.vals = (unsigned long long *)(w + off[9]);
            slots.flags = (unsigned int *)(slots.vals + 256);
            DC_CUDA_TRY(cudaMemsetAsync(slots.flags, 0, 256 * sizeof(unsigned int), st));
            const unsigned int grid = min(min(sms, 256u), (nruns + 63) / 64);
            LaunchScope ls(DC_K_ENCODE_PLAN, st);
            encode_plan_kernel<<<grid, kPlanThreads, 0, st>>>((const uint16_t *)(w + off[8]), d_table, ws, nruns, slots,
                                                                                   (unsigned long long *)d_total_bits, d_status);
        }
        return launch_encode_body(d_in, n, d_table, d_out, out_capacity, bit_phase, ws, w, off, nruns, (unsigned long long *)d_total_bits, d_status, st, true);
    }
    // the bit offset of every run, from the input: E1 (bits per run) + E2 (exclusive scan, total)
    {
        LaunchScope ls(DC_K_ENCODE_COUNT, st);
        encode_count_kernel<<<min(nruns, sms * 8u), kEncThreads, 0, st>>>(d_in, n, d_table, ws, nruns, d_status);
    }
    {
        LaunchScope ls(DC_K_ENCODE_SCAN, st);
        encode_scan_kernel<<<1, kScanThreads, 0, st>>>(d_table, ws, nruns, (unsigned long long *)d_total_bits);
    }
    return launch_encode_body(d_in, n, d_table, d_out, out_capacity, bit_phase, ws, w, off, nruns, (unsigned long long *)d_total_bits, d_status, st, false);
}

extern "C" int dc_huff_encode(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity,
                              unsigned bit_phase, uint64_t *d_total_bits, int32_t *d_status, void *d_workspace, size_t workspace_bytes,
                              void *stream) {
    return encode_entry(d_in, n, d_table, d_out, out_capacity, bit_phase, d_total_bits, d_status, d_workspace, workspace_bytes, stream, false);
}

extern "C" int dc_huff_encode_planned(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity,
                                      unsigned bit_phase, uint64_t *d_total_bits, int32_t *d_status, void *d_workspace,
                                      size_t workspace_bytes, void *stream) {
    return encode_entry(d_in, n, d_table, d_out, out_capacity, bit_phase, d_total_bits, d_status, d_workspace, workspace_bytes, stream, true);
}

namespace dc {
// dc_huff_encode_planned with the bit phase in device memory (shard_nccl.cu: the phase comes out of a collective)
int encode_planned_device_phase(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity,
                                const uint32_t *d_phase, uint64_t *d_total_bits, int32_t *d_status, void *d_workspace,
                                size_t workspace_bytes, cudaStream_t st) {
    return encode_entry(d_in, n, d_table, d_out, out_capacity, 0, d_total_bits, d_status, d_workspace, workspace_bytes, (void *)st, true, d_phase);
}
}  // namespace dc

namespace dc {
int histogram_runs_edges(const uint8_t *d_in, size_t n, unsigned long long *d_hist, void *d_encode_workspace, size_t workspace_bytes,
                         unsigned long long *d_edge, cudaStream_t st) {
    if (!d_hist || (n && (!d_in || !d_encode_workspace))) return DC_ERR_ARG;
    if ((((uintptr_t)d_in | (uintptr_t)d_encode_workspace) & 15) != 0) return DC_ERR_ARG;
    size_t off[10];
    if (workspace_bytes < enc_ws_layout(n, off)) return DC_ERR_CAPACITY;
    return launch_histogram_runs(d_in, n, d_hist, (uint16_t *)((char *)d_encode_workspace + off[8]), st, d_edge);
}
}  // namespace dc

extern "C" int dc_histogram_u8_runs(const uint8_t *d_in, size_t n, uint64_t *d_hist, void *d_encode_workspace, size_t workspace_bytes,
                                    void *stream) {
    if (!d_hist || (n && (!d_in || !d_encode_workspace))) return DC_ERR_ARG;
    if ((((uintptr_t)d_in | (uintptr_t)d_encode_workspace) & 15) != 0) return DC_ERR_ARG;
    size_t off[10];
    if (workspace_bytes < enc_ws_layout(n, off)) return DC_ERR_CAPACITY;
    return launch_histogram_runs(d_in, n, (unsigned long long *)d_hist, (uint16_t *)((char *)d_encode_workspace + off[8]), (cudaStream_t)stream);
}

// the kernels behind the planned run offsets (ws.run_off): exactly one of them does the work, the others return at once
static int launch_encode_body(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity,
                              unsigned bit_phase, EncWorkspace ws, char *w, const size_t *off, unsigned int nruns,
                              unsigned long long *d_total_bits, int32_t *d_status, cudaStream_t st, bool planned) {
    (void)planned; (void)w; (void)off; (void)d_total_bits;
    const unsigned int sms = (unsigned int)sm_count();
    FwWorkspace fw;
    fw.run_off = ws.run_off;
    fw.bstate = ws.bstate;
    fw.bleft = ws.bleft;
    fw.bright = ws.bright;
    fw.d_phase = ws.d_phase;
    // the table's longest code, if the host knows it already (the build's event has passed); else the kernels decide
    int32_t tmeta[kTableMetaWords];
    const bool known = table_meta_fetch(d_table, st, tmeta, false) == DC_OK;
    const int mb = known ? tmeta[7] : -1;
    if (!known || mb <= kNarrowBits) {
        // codes up to 12 bits: two symbols per emit step; 13 .. 16 bits: one symbol per emit step
        const bool pair = known && mb <= kPlannedMaxBits;
        const void *fn = !known ? (const void *)encode_fast_kernel<0> : pair ? (const void *)encode_fast_kernel<kPlannedMaxBits> : (const void *)encode_fast_kernel<kNarrowBits>;
        DC_CUDA_TRY(ensure_dynamic_smem(fn, kFwSmemBytes));
        const unsigned int grid = pair ? min((nruns + 1) / 2, sms) : min(nruns, sms);
        LaunchScope ls(known && !pair ? DC_K_ENCODE_MID : DC_K_ENCODE_FAST, st);
        if (!known) encode_fast_kernel<0><<<grid, 1024, kFwSmemBytes, st>>>(d_in, n, d_table, d_out, out_capacity, bit_phase, fw, nruns, d_status, (uint32_t)kFwSmemBytes);
        else if (pair) encode_fast_kernel<kPlannedMaxBits><<<grid, 1024, kFwSmemBytes, st>>>(d_in, n, d_table, d_out, out_capacity, bit_phase, fw, nruns, d_status, (uint32_t)kFwSmemBytes);
        else encode_fast_kernel<kNarrowBits><<<grid, 1024, kFwSmemBytes, st>>>(d_in, n, d_table, d_out, out_capacity, bit_phase, fw, nruns, d_status, (uint32_t)kFwSmemBytes);
    }
    if (!known || mb > kNarrowBits) {   // longer codes: 64-bit entries (returns at once for any other table)
        LaunchScope ls(DC_K_ENCODE_WIDE, st);
        encode_run_kernel<true><<<min(nruns, known ? sms * 4u : sms), kEncThreads, 0, st>>>(d_in, n, d_table, d_out, out_capacity, bit_phase,
                                                                             ws, nruns, d_status, known ? 1 : 0);
    }
    return cuda_status(cudaGetLastError());
}
