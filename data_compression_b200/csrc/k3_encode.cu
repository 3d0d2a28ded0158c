// k3_encode.cu -- K3: Huffman payload encode in a single pass over the input.
//
// Replaces represent_items_with_codes() (n_ary_huffman.c:1621-1678).  In the reference that function is a
// stub (assert(0) at :1661, returns 32767); its stated intent is "for each input byte append
// encode_length digits of encode_value, completely bit-oriented" (:1651-1657).  Payload layout
// (DESIGN.md): a code of `len` digits is the len*log2(n)-bit big-endian numeral of encode_value; codes are
// concatenated in input order, MSB-first within each byte; the final byte is zero-padded.
//
// One CTA per tile of 4096 symbols (256 threads x 16 symbols = one 16-byte load per thread):
//   A. per-symbol code lookup from a shared-memory copy of the table; per-thread bit total
//   B. CTA-wide exclusive scan of bit totals; the tile total is published at once (decoupled look-back)
//   C. every thread streams its codes through a 64-bit accumulator into a shared-memory staging buffer at
//      its tile-relative bit offset (whole words: plain stores; the two edge words: shared atomicOr)
//   D. the tile publishes its last 128 bits; warp 0 resolves the tile's global bit offset by look-back
//   E. coalesced copy-out: every thread funnel-shifts staging words to the global alignment and stores
//      16 bytes.  A 16-byte output word that straddles two tiles is written by the LATER tile, which
//      pulls the missing leading bits from its predecessor's published tail -- no global atomics, no
//      pre-zeroed output, every output byte written exactly once.
// HBM traffic = N (read) + C (write) + 24 bytes of descriptor per tile.
#include "dc_common.cuh"

namespace dc {

constexpr int kEncThreads = 256;
constexpr int kEncPerThread = 16;
constexpr int kEncTile = kEncThreads * kEncPerThread;        // 4096 symbols
constexpr int kEncStageWords = 4 + kEncTile + 8;             // pred. tail + 32 bits/symbol worst case + pad

constexpr unsigned long long kDescAgg = 1ull << 62, kDescPrefix = 2ull << 62, kDescTail = 1ull << 61;
constexpr unsigned long long kDescValue = (1ull << 61) - 1;

struct EncWorkspace {
    unsigned int *ticket;        // dynamic tile id (tiles must start in id order for the look-back)
    unsigned long long *desc;    // [ntiles] status | tail-ready | bits
    uint4 *tails;                // [ntiles] last 128 bits of each tile, big-endian word domain
};

// bits [bit, bit+32) of a big-endian word array
__device__ __forceinline__ uint32_t stage_word(const uint32_t *stage, uint32_t bit) {
    const uint32_t a = bit >> 5, s = bit & 31;
    return __funnelshift_l(stage[a + 1], stage[a], s);
}

template <bool WIDE>
__device__ __forceinline__ void encode_tile(const uint8_t *__restrict__ in, size_t n, const dc_huff_table *__restrict__ tab,
                                            uint8_t *__restrict__ out, size_t out_cap, unsigned phase,
                                            unsigned long long *__restrict__ d_total_bits, int32_t *__restrict__ d_status,
                                            const EncWorkspace &ws, unsigned int ntiles, uint32_t *smem) {
    typedef typename std::conditional<WIDE, unsigned long long, uint32_t>::type entry_t;
    uint32_t *stage = smem;                                        // [kEncStageWords]
    entry_t *s_enc = (entry_t *)(smem + kEncStageWords);           // [256]
    __shared__ uint32_t s_warp_bits[kEncThreads / 32];
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_excl;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ws.ticket, 1u);
    s_enc[tid] = WIDE ? (entry_t)tab->enc64[tid] : (entry_t)tab->enc[tid];
    for (int i = tid; i < kEncStageWords / 4; i += kEncThreads) ((uint4 *)stage)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const unsigned int tile = s_tile;
    const size_t base = (size_t)tile * kEncTile + (size_t)tid * kEncPerThread;

    // ---- A. load 16 symbols, look their codes up
    int valid = 0;
    uint32_t w[4] = {0, 0, 0, 0};
    if (base + kEncPerThread <= n) {
        const uint4 v = ldg_stream((const uint4 *)(in + base));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        valid = kEncPerThread;
    } else if (base < n) {
        valid = (int)(n - base);
        for (int k = 0; k < valid; k++) w[k >> 2] |= (uint32_t)in[base + k] << (8 * (k & 3));
    }
    entry_t e[kEncPerThread];
    uint32_t my_bits = 0;
    bool missing = false;
#pragma unroll
    for (int k = 0; k < kEncPerThread; k++) {
        const uint32_t b = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
        entry_t x = s_enc[b];
        if (k >= valid) x = 0;
        const uint32_t len = WIDE ? (uint32_t)(x >> 32) : ((uint32_t)x & 63u);
        missing |= (k < valid) && (len == 0);
        my_bits += len;
        e[k] = x;
    }
    if (missing) set_status(d_status, DC_ERR_SYMBOL);

    // ---- B. exclusive scan of bit totals over the CTA
    uint32_t incl = my_bits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp_bits[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, tile_bits = 0;
#pragma unroll
    for (int i = 0; i < kEncThreads / 32; i++) {
        const uint32_t t = s_warp_bits[i];
        if (i < warp) warp_off += t;
        tile_bits += t;
    }
    const unsigned long long first_status = tile == 0 ? kDescPrefix : kDescAgg;
    if (tid == 0) st_release_u64(&ws.desc[tile], first_status | tile_bits);

    // ---- C. stream this thread's codes into the staging buffer (own bits start at staging bit 128)
    if (my_bits) {
        const uint32_t pos = 128u + warp_off + incl - my_bits;
        uint32_t wi = pos >> 5;
        uint32_t nb = pos & 31;
        bool shared_word = nb != 0;
        unsigned long long acc = 0;
#pragma unroll
        for (int k = 0; k < kEncPerThread; k++) {
            const uint32_t len = WIDE ? (uint32_t)(e[k] >> 32) : ((uint32_t)e[k] & 63u);
            const uint32_t val = WIDE ? (uint32_t)e[k] : ((uint32_t)e[k] >> 6);
            acc = (acc << len) | val;
            nb += len;
            if (nb >= 32) {
                const uint32_t word = (uint32_t)(acc >> (nb - 32));
                if (shared_word) { atomicOr(&stage[wi], word); shared_word = false; }
                else stage[wi] = word;
                wi++;
                nb -= 32;
            }
        }
        if (nb) atomicOr(&stage[wi], (uint32_t)(acc << (32 - nb)));
    }
    __syncthreads();

    // ---- D. publish the tail, resolve the global bit offset, fetch the predecessor's tail
    if (warp == 0) {
        if (lane == 0) {
            uint4 t;
            t.x = stage_word(stage, tile_bits);
            t.y = stage_word(stage, tile_bits + 32);
            t.z = stage_word(stage, tile_bits + 64);
            t.w = stage_word(stage, tile_bits + 96);
            ws.tails[tile] = t;
            __threadfence();
            st_release_u64(&ws.desc[tile], first_status | kDescTail | tile_bits);
        }
        unsigned long long excl = 0;
        if (tile != 0) {
            long long idx = (long long)tile - 1 - lane;
            while (true) {
                unsigned long long d;
                do {
                    d = idx >= 0 ? ld_acquire_u64(&ws.desc[idx]) : kDescPrefix;
                } while (__any_sync(0xFFFFFFFFu, (d >> 62) == 0));
                const unsigned prefix_mask = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
                unsigned long long v = d & kDescValue;
                if (prefix_mask) {
                    const int first = __ffs(prefix_mask) - 1;  // nearest predecessor that knows its prefix
                    if (lane > first) v = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                excl += v;
                if (prefix_mask) break;
                idx -= 32;
            }
            if (lane == 0) {
                st_release_u64(&ws.desc[tile], kDescPrefix | kDescTail | (excl + tile_bits));
                while ((ld_acquire_u64(&ws.desc[tile - 1]) & kDescTail) == 0) {}
            }
            __syncwarp();
            if (lane == 0) {
                const uint4 t = ld_cg_u128(&ws.tails[tile - 1]);
                stage[0] = t.x; stage[1] = t.y; stage[2] = t.z; stage[3] = t.w;
            }
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();

    // ---- E. copy-out at the global alignment
    const unsigned long long g = (unsigned long long)phase + s_excl;   // global bit position of the tile's first bit
    const unsigned long long gend = g + tile_bits;
    const unsigned long long v0 = g >> 7, v1 = gend >> 7;              // 16-byte words [v0, v1) end inside this tile
    const uint32_t r = (uint32_t)(g & 127);
    const bool last = tile == ntiles - 1;
    const size_t need = (size_t)((last ? gend + 7 : v1 * 128) >> 3);
    if (need > out_cap) {
        if (tid == 0) set_status(d_status, DC_ERR_CAPACITY);
    } else {
        for (unsigned long long v = v0 + tid; v < v1; v += kEncThreads) {
            const uint32_t sbit = (uint32_t)(v - v0) * 128u + (128u - r);
            uint4 o;
            o.x = bswap32(stage_word(stage, sbit));
            o.y = bswap32(stage_word(stage, sbit + 32));
            o.z = bswap32(stage_word(stage, sbit + 64));
            o.w = bswap32(stage_word(stage, sbit + 96));
            stg_stream((uint4 *)out + v, o);
        }
        if (last) {  // trailing partial 16-byte word of the stream, byte by byte (zero padded)
            const uint32_t rem_bytes = (uint32_t)(((gend & 127) + 7) >> 3);
            const uint32_t sbit = (uint32_t)(v1 - v0) * 128u + (128u - r);
            if (tid < (int)rem_bytes) out[v1 * 16 + tid] = (uint8_t)(stage_word(stage, sbit + 8u * tid) >> 24);
        }
    }
    if (last && tid == 0 && d_total_bits) *d_total_bits = s_excl + tile_bits;
}

__global__ void __launch_bounds__(kEncThreads) encode_kernel(const uint8_t *__restrict__ in, size_t n,
                                                             const dc_huff_table *__restrict__ tab, uint8_t *__restrict__ out,
                                                             size_t out_cap, unsigned phase,
                                                             unsigned long long *__restrict__ d_total_bits,
                                                             int32_t *__restrict__ d_status, EncWorkspace ws,
                                                             unsigned int ntiles) {
    extern __shared__ __align__(16) uint32_t enc_smem[];
    const int tstatus = tab->status, bpd = tab->bits_per_digit, max_bits = tab->max_bits;
    if (tstatus != DC_OK || bpd == 0) {  // uniform across the grid: nobody takes a ticket
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            set_status(d_status, tstatus != DC_OK ? tstatus : DC_ERR_RADIX);
            if (d_total_bits) *d_total_bits = 0;
        }
        return;
    }
    if (max_bits <= 26) encode_tile<false>(in, n, tab, out, out_cap, phase, d_total_bits, d_status, ws, ntiles, enc_smem);
    else encode_tile<true>(in, n, tab, out, out_cap, phase, d_total_bits, d_status, ws, ntiles, enc_smem);
}

static size_t enc_ws_layout(size_t n, size_t *desc_off, size_t *tails_off) {
    const size_t ntiles = (n + kEncTile - 1) / kEncTile;
    const size_t d = 16;
    const size_t t = d + ((ntiles * 8 + 15) & ~(size_t)15);
    if (desc_off) *desc_off = d;
    if (tails_off) *tails_off = t;
    return t + ntiles * 16;
}

}  // namespace dc

using namespace dc;

extern "C" size_t dc_huff_encode_workspace_bytes(size_t n) { return enc_ws_layout(n, nullptr, nullptr); }

extern "C" int dc_huff_encode(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out,
                              size_t out_capacity, unsigned bit_phase, uint64_t *d_total_bits, int32_t *d_status,
                              void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!d_table || (n && (!d_in || !d_out || !d_workspace)) || bit_phase > 7) return DC_ERR_ARG;
    if ((((uintptr_t)d_in | (uintptr_t)d_out | (uintptr_t)d_workspace) & 15) != 0) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (d_total_bits) DC_CUDA_TRY(cudaMemsetAsync(d_total_bits, 0, sizeof(uint64_t), st));
    if (n == 0) return DC_OK;
    size_t desc_off, tails_off;
    const size_t need = enc_ws_layout(n, &desc_off, &tails_off);
    if (workspace_bytes < need) return DC_ERR_CAPACITY;
    const size_t ntiles = (n + kEncTile - 1) / kEncTile;
    if (ntiles > 0xFFFFFFF0ull) return DC_ERR_ARG;
    // ticket + descriptors must start at zero; the tails are written before they are read
    DC_CUDA_TRY(cudaMemsetAsync(d_workspace, 0, tails_off, st));
    EncWorkspace ws;
    ws.ticket = (unsigned int *)d_workspace;
    ws.desc = (unsigned long long *)((char *)d_workspace + desc_off);
    ws.tails = (uint4 *)((char *)d_workspace + tails_off);
    const size_t smem = (size_t)kEncStageWords * 4 + 256 * 8;
    LaunchScope ls(DC_K_ENCODE, st);
    encode_kernel<<<(unsigned int)ntiles, kEncThreads, smem, st>>>(d_in, n, d_table, d_out, out_capacity, bit_phase,
                                                                  (unsigned long long *)d_total_bits, d_status, ws,
                                                                  (unsigned int)ntiles);
    return cuda_status(cudaGetLastError());
}
