// k8_mtf.cu -- K8: the move-to-front contexts of the adaptive nybble compressor.
//
// With modify == true (nybble_compress nybble_compression.c:1134, nybble_decompress :1117) each of the 16 contexts --
// chosen by bits 3..6 of the previous byte, byte_to_context :517-523 -- is a move-to-front list of 8 letters that starts
// as " etaoins" (initialize_dictionary :546-562) and is touched after every byte (update_context :665-687).  A byte is a
// table hit when it is in its context's list, and the nibble it emits is 8 | position (compress_byte_index :819-884).
//
// Compress.  The contexts are known from the input, and a list is nothing but "the 8 most recently seen distinct bytes
// of this context, most recent first; behind them what is left of the older list".  That makes the effect of a block of
// input on a list a monoid element: the block's own recency list (<= 8 distinct bytes), and
//     (older list) . (block list)  =  block list, then the older entries not in it, cut to 8.
// So the hit positions of every byte come from a scan, not from a serial walk over the string:
//   M1  one thread per 512-byte block walks the 16 lists from "holes" (eight place-holders 0x80..0x87, which no 7-bit
//       input byte equals): what it leaves is the block's element, holes marking what the older list fills in
//   M2  reduce over chunks of 64 blocks, a scan over the chunk elements (one CTA per context), apply: every
//       block's element is replaced by the exact lists in front of it
//   M3  the M1 walk again from the exact lists, writing one byte per input byte: position 0..7, or 8 = not in the list
// K6's transducer (k6_nybble_text.cu) then runs on these positions instead of its static table.
//
// Decompress.  The nibble parse is K6's, but which letter a hit nibble means depends on the context, i.e. on the byte
// decoded just before, and on every list update before that: a serial chain through the output ("replicas
// only").  K6 writes hit nibbles as 0x80 | position; mtf_resolve_kernel walks the output once and replaces them --
// one lane does the chain, the warp moves the bytes.  About 11 MB/s: correct, and only parallel across strings.
#include "dc_common.cuh"

namespace dc {

constexpr int kMtfThreads = 128;
constexpr int kMtfBlock = 512;   // input bytes per thread
constexpr int kMtfChunk = 64;    // blocks per chunk
constexpr int kMtfCtx = 16;
constexpr unsigned long long kMtfInit = 0x736e696f61746520ull;   // " etaoins", byte 0 = front of the list
constexpr unsigned long long kMtfHoles = 0x8786858483828180ull;

__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) { return (x - 0x01010101u) & ~x & 0x80808080u; }  // lowest flag is exact

// position of b in the list (8 = absent)
__device__ __forceinline__ int mtf_find(unsigned long long w, uint32_t b) {
    const uint32_t bb = b * 0x01010101u;
    const uint32_t zl = zero_bytes((uint32_t)w ^ bb), zh = zero_bytes((uint32_t)(w >> 32) ^ bb);
    return zl ? (__ffs(zl) - 1) >> 3 : (zh ? 4 + ((__ffs(zh) - 1) >> 3) : 8);
}
// b to the front; the entries in front of its old position move back by one (absent: the last one falls off) :665-687
__device__ __forceinline__ unsigned long long mtf_touch(unsigned long long w, uint32_t b, int pos) {
    const int at = pos < 8 ? pos : 7;
    const unsigned long long low = (1ull << (8 * at)) - 1ull;
    const unsigned long long keep = ~((low << 8) | 0xFFull);
    return ((w & low) << 8) | (w & keep) | (unsigned long long)b;
}
__device__ __forceinline__ uint32_t mtf_ctx(uint32_t prev) { return (prev >> 3) & 15u; }

// (older list s) . (block element w)
__device__ __forceinline__ unsigned long long mtf_compose(unsigned long long s, unsigned long long w) {
    unsigned long long hm = w & 0x8080808080808080ull;
    if (hm == 0) return w;
    hm = (hm >> 7) * 0xFFull;                   // 0xFF in every hole
    const unsigned long long wr = w | hm;        // holes match nothing
    unsigned long long out = w & ~hm;
    int cnt = 8 - (__popcll(hm) >> 3);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t e = (uint32_t)(s >> (8 * j)) & 0xFFu;
        if (cnt < 8 && mtf_find(wr, e) == 8) {
            out |= (unsigned long long)e << (8 * cnt);
            cnt++;
        }
    }
    return out;
}

// ------------------------------------------------------------------------------------------ M1 / M3
// lists[block][ctx]: M1 (CODES == false) writes the block's element, M3 reads the lists in front of the block.
template <bool CODES>
__global__ void __launch_bounds__(kMtfThreads) mtf_walk_kernel(const uint8_t *__restrict__ in, size_t n, unsigned long long *__restrict__ lists,
                                                               size_t nblocks, uint8_t *__restrict__ pos_out) {
    __shared__ unsigned long long s_list[kMtfCtx][kMtfThreads];  // [ctx][thread]: conflict-free for any mix of contexts
    const int tid = threadIdx.x;
    for (size_t g = (size_t)blockIdx.x * kMtfThreads + tid; g < nblocks; g += (size_t)gridDim.x * kMtfThreads) {
        unsigned long long *mine = lists + g * kMtfCtx;
#pragma unroll
        for (int c = 0; c < kMtfCtx; c++) s_list[c][tid] = CODES ? mine[c] : kMtfHoles;
        const size_t lo = g * kMtfBlock, hi = min(n, lo + kMtfBlock);
        uint32_t prev = lo ? in[lo - 1] : 0u;
        const bool vec = (((uintptr_t)in | (CODES ? (uintptr_t)pos_out : 0)) & 15) == 0;
        for (size_t p = lo; p < hi; p += 16) {
            uint32_t w[4], o[4] = {0, 0, 0, 0};
            const int cnt = (int)min((size_t)16, hi - p);
            if (vec && cnt == 16) {
                const uint4 v = __ldg((const uint4 *)(in + p));
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t x = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) x |= (4 * j + k < cnt ? (uint32_t)in[p + 4 * j + k] : 0u) << (8 * k);
                    w[j] = x;
                }
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t b = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                uint32_t at = 8;
                if (k < cnt && p + k != 0) {   // the first byte of the string is copied, not coded (:905)
                    const uint32_t c = mtf_ctx(prev);
                    const unsigned long long l = s_list[c][tid];
                    at = (uint32_t)mtf_find(l, b);
                    s_list[c][tid] = mtf_touch(l, b, (int)at);
                }
                if (CODES) o[k >> 2] |= at << (8 * (k & 3));
                prev = b;
            }
            if (CODES) {
                if (vec && cnt == 16) {
                    *(uint4 *)(pos_out + p) = make_uint4(o[0], o[1], o[2], o[3]);
                } else {
                    for (int k = 0; k < cnt; k++) pos_out[p + k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
                }
            }
        }
        if (!CODES) {
#pragma unroll
            for (int c = 0; c < kMtfCtx; c++) mine[c] = s_list[c][tid];
        }
    }
}

// ------------------------------------------------------------------------------------------ M2
// one thread per (chunk, context): the chunk's element
__global__ void __launch_bounds__(256) mtf_reduce_kernel(const unsigned long long *__restrict__ lists, size_t nblocks,
                                                         unsigned long long *__restrict__ chunks, size_t nchunks) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= nchunks * kMtfCtx) return;
    const size_t chunk = t / kMtfCtx, c = t % kMtfCtx;
    const size_t b0 = chunk * kMtfChunk, b1 = min(nblocks, b0 + kMtfChunk);
    // eight loads in flight, then the eight compositions (the chain is serial, the loads are not)
    unsigned long long acc = kMtfHoles;
    for (size_t b = b0; b < b1; b += 8) {
        unsigned long long e[8];
#pragma unroll
        for (int j = 0; j < 8; j++) e[j] = b + j < b1 ? lists[(b + j) * kMtfCtx + c] : kMtfHoles;
#pragma unroll
        for (int j = 0; j < 8; j++) acc = mtf_compose(acc, e[j]);
    }
    chunks[t] = acc;
}

// one CTA per context: chunk elements -> the exact list in front of every chunk (in place)
constexpr int kMtfTopThreads = 1024;
__global__ void __launch_bounds__(kMtfTopThreads) mtf_top_kernel(unsigned long long *__restrict__ chunks, size_t nchunks) {
    __shared__ unsigned long long s_warp[kMtfTopThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, c = blockIdx.x;
    const size_t per = (nchunks + kMtfTopThreads - 1) / kMtfTopThreads;
    const size_t k0 = min(nchunks, (size_t)tid * per), k1 = min(nchunks, k0 + per);
    unsigned long long acc = kMtfHoles;   // the identity: nothing seen
    for (size_t k = k0; k < k1; k++) acc = mtf_compose(acc, chunks[k * kMtfCtx + c]);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long left = __shfl_up_sync(0xFFFFFFFFu, acc, d);
        if (lane >= d) acc = mtf_compose(left, acc);
    }
    if (lane == 31) s_warp[warp] = acc;
    __syncthreads();
    unsigned long long st = kMtfInit;     // the list in front of this thread's first chunk
    for (int w = 0; w < warp; w++) st = mtf_compose(st, s_warp[w]);
    const unsigned long long left = __shfl_up_sync(0xFFFFFFFFu, acc, 1);
    if (lane) st = mtf_compose(st, left);
    for (size_t k = k0; k < k1; k++) {
        const unsigned long long e = chunks[k * kMtfCtx + c];
        chunks[k * kMtfCtx + c] = st;
        st = mtf_compose(st, e);
    }
}

// one thread per (chunk, context): block elements -> the exact list in front of every block (in place)
__global__ void __launch_bounds__(256) mtf_apply_kernel(unsigned long long *__restrict__ lists, size_t nblocks,
                                                        const unsigned long long *__restrict__ chunks, size_t nchunks) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= nchunks * kMtfCtx) return;
    const size_t chunk = t / kMtfCtx, c = t % kMtfCtx;
    const size_t b0 = chunk * kMtfChunk, b1 = min(nblocks, b0 + kMtfChunk);
    unsigned long long st = chunks[t];
    for (size_t b = b0; b < b1; b += 8) {
        unsigned long long e[8];
#pragma unroll
        for (int j = 0; j < 8; j++) e[j] = b + j < b1 ? lists[(b + j) * kMtfCtx + c] : kMtfHoles;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (b + j < b1) lists[(b + j) * kMtfCtx + c] = st;
            st = mtf_compose(st, e[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------ decompress: the serial chain
// buf[0] is the verbatim first byte; buf[1..len): a literal (< 0x80) or 0x80 | position.  One warp: 512-byte tiles
// through shared memory, lane 0 resolves.
__global__ void __launch_bounds__(32) mtf_resolve_kernel(uint8_t *__restrict__ buf, const unsigned long long *__restrict__ d_len,
                                                         const int32_t *__restrict__ d_mode) {
    __shared__ unsigned long long s_list[kMtfCtx];
    __shared__ uint8_t s_tile[512];
    if (*d_mode != 0) return;   // not an 0xAF stream: nothing was coded
    const unsigned long long len = *d_len;
    const int lane = threadIdx.x;
    if (lane < kMtfCtx) s_list[lane] = kMtfInit;
    uint32_t prev = len ? buf[0] : 0u;
    __syncwarp();
    for (unsigned long long base = 1; base < len; base += 512) {
        const int cnt = (int)min((unsigned long long)512, len - base);
        for (int k = lane; k < cnt; k += 32) s_tile[k] = buf[base + k];
        __syncwarp();
        if (lane == 0) {
            for (int k = 0; k < cnt; k++) {
                const uint32_t x = s_tile[k], c = mtf_ctx(prev);
                const unsigned long long l = s_list[c];
                uint32_t b, at;
                if (x & 0x80u) { at = x & 7u; b = (uint32_t)(l >> (8 * at)) & 0xFFu; }
                else { b = x; at = (uint32_t)mtf_find(l, b); }
                s_list[c] = mtf_touch(l, b, (int)at);
                s_tile[k] = (uint8_t)b;
                prev = b;
            }
        }
        __syncwarp();
        for (int k = lane; k < cnt; k += 32) buf[base + k] = s_tile[k];
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ replicas: many strings per call
// SURVEY 8e gives the adaptive compressor "replicas only" as its parallelism; the reference's own use is many short
// strings (messages of a microcontroller program).  One thread per string walks it exactly as the reference does --
// compress_bytestring :887-1038 / decompress_bytestring :734-817, either mode -- with the 16 lists in local memory.
struct MtfLists {
    unsigned long long l[kMtfCtx];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int c = 0; c < kMtfCtx; c++) l[c] = kMtfInit;
    }
};

// returns the compressed length; status: DC_ERR_SYMBOL for a byte outside 0x01..0x7F, DC_ERR_CAPACITY if cap < n + 2
__device__ size_t text_compress_one(const uint8_t *__restrict__ src, size_t n, uint8_t *__restrict__ dst, size_t cap, bool modify, int *status) {
    if (n == 0) { if (cap) dst[0] = 0; return 0; }
    if (cap < n + 2) { *status = DC_ERR_CAPACITY; return 0; }
    MtfLists t;
    t.init();
    size_t o = 0;
    dst[o++] = 0xAF;
    dst[o++] = src[0];
    int off = 0;
    bool bad = false;
    for (size_t i = 1; i < n; i++) {
        const uint32_t s = src[i], c = mtf_ctx(src[i - 1]);
        bad |= s == 0 || s >= 0x80u;
        const int at = mtf_find(t.l[c], s);
        int used;
        if (at == 8) {
            if (off == 0) { dst[o] = (uint8_t)s; used = 2; }
            else { dst[o] = src[i - 1]; dst[o + 1] = (uint8_t)s; used = 3; }   // the half-written byte becomes a literal (:855-857)
        } else {
            const uint32_t nyb = 8u | (uint32_t)at;
            if (off == 0) dst[o] = (uint8_t)(nyb << 4);
            else dst[o] |= (uint8_t)nyb;
            used = 1;
        }
        if (modify) t.l[c] = mtf_touch(t.l[c], s, at);
        off += used;
        if (off > 1) { o++; off -= 2; }
        if (off > 1) { o++; off -= 2; }
    }
    if (off != 0) dst[o++] = src[n - 1];                 // trailing half byte -> literal (:1000-1009)
    if (bad) { *status = DC_ERR_SYMBOL; return 0; }
    if (o >= n) {                                        // not shorter: ' ' + raw copy (:1018-1037)
        dst[0] = ' ';
        for (size_t i = 0; i < n; i++) dst[1 + i] = src[i];
        o = n + 1;
    }
    if (o < cap) dst[o] = 0;
    return o;
}

__device__ size_t text_decompress_one(const uint8_t *__restrict__ src, size_t n, uint8_t *__restrict__ dst, size_t cap, bool modify, int *status) {
    size_t o = 0;
    if (n == 0) { if (cap) dst[0] = 0; return 0; }
    if (src[0] == 0xAFu) {
        if (n < 2) { if (cap) dst[0] = 0; return 0; }
        if (cap < 2 * n) { *status = DC_ERR_CAPACITY; return 0; }   // at most 2n - 3 bytes
        MtfLists t;
        t.init();
        dst[o++] = src[1];
        size_t p = 2;
        int off = 0;
        while (p < n) {
            const uint32_t b = src[p], nb = p + 1 < n ? src[p + 1] : 0u;
            const uint32_t nyb = off == 0 ? (b >> 4) & 15u : b & 15u, next = off == 0 ? b & 15u : (nb >> 4) & 15u;
            const uint32_t c = mtf_ctx(dst[o - 1]);
            uint32_t out;
            int at;
            if (nyb & 8u) { at = (int)(nyb & 7u); out = (uint32_t)(t.l[c] >> (8 * at)) & 0xFFu; off += 1; }
            else { out = ((nyb & 7u) << 4) + next; at = mtf_find(t.l[c], out); off += 2; }
            dst[o++] = (uint8_t)out;
            if (modify) t.l[c] = mtf_touch(t.l[c], out, at);
            if (off >= 2) { p++; off -= 2; }
        }
    } else {
        size_t p = src[0] == ' ' ? 1 : 0;                 // LITERAL skips the type byte (:799-805), anything else copies (:806-812)
        if (cap < n) { *status = DC_ERR_CAPACITY; return 0; }
        while (p < n) dst[o++] = src[p++];
    }
    if (o < cap) dst[o] = 0;
    return o;
}

template <bool COMPRESS>
__global__ void __launch_bounds__(128) text_batch_kernel(const uint8_t *__restrict__ src, const unsigned long long *__restrict__ src_off,
                                                         size_t count, int modify, uint8_t *__restrict__ dst,
                                                         const unsigned long long *__restrict__ dst_off,
                                                         unsigned long long *__restrict__ out_len, int32_t *__restrict__ d_status) {
    const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= count) return;
    const unsigned long long s0 = src_off[i], s1 = src_off[i + 1], d0 = dst_off[i], d1 = dst_off[i + 1];
    int st = DC_OK;
    size_t len = 0;
    if (s1 < s0 || d1 < d0) st = DC_ERR_ARG;
    else len = COMPRESS ? text_compress_one(src + s0, (size_t)(s1 - s0), dst + d0, (size_t)(d1 - d0), modify != 0, &st)
                        : text_decompress_one(src + s0, (size_t)(s1 - s0), dst + d0, (size_t)(d1 - d0), modify != 0, &st);
    out_len[i] = len;
    if (st != DC_OK) set_status(d_status, st);
}

// ------------------------------------------------------------------------------------------ host side (internal)
static size_t mtf_blocks(size_t n) { return (n + kMtfBlock - 1) / kMtfBlock; }
static size_t mtf_chunks(size_t n) { return (mtf_blocks(n) + kMtfChunk - 1) / kMtfChunk; }

size_t mtf_workspace_bytes(size_t n) {
    return ((mtf_blocks(n) * kMtfCtx * 8 + 63) & ~(size_t)63) + ((mtf_chunks(n) * kMtfCtx * 8 + 63) & ~(size_t)63) + 64;
}

int mtf_positions(const uint8_t *d_src, size_t n, uint8_t *d_pos, void *d_ws, cudaStream_t st) {
    if (n == 0) return DC_OK;
    const size_t nblocks = mtf_blocks(n), nchunks = mtf_chunks(n);
    unsigned long long *lists = (unsigned long long *)d_ws;
    unsigned long long *chunks = (unsigned long long *)((char *)d_ws + ((nblocks * kMtfCtx * 8 + 63) & ~(size_t)63));
    const unsigned int walk_grid = (unsigned int)min((nblocks + kMtfThreads - 1) / kMtfThreads, (size_t)sm_count() * 12);
    const unsigned int cc_grid = (unsigned int)((nchunks * kMtfCtx + 255) / 256);
    {
        LaunchScope ls(DC_K_MTF_WALK, st);
        mtf_walk_kernel<false><<<walk_grid, kMtfThreads, 0, st>>>(d_src, n, lists, nblocks, nullptr);
    }
    {
        LaunchScope ls(DC_K_MTF_SCAN, st);
        mtf_reduce_kernel<<<cc_grid, 256, 0, st>>>(lists, nblocks, chunks, nchunks);
    }
    {
        LaunchScope ls(DC_K_MTF_SCAN, st);
        mtf_top_kernel<<<kMtfCtx, kMtfTopThreads, 0, st>>>(chunks, nchunks);
    }
    {
        LaunchScope ls(DC_K_MTF_SCAN, st);
        mtf_apply_kernel<<<cc_grid, 256, 0, st>>>(lists, nblocks, chunks, nchunks);
    }
    {
        LaunchScope ls(DC_K_MTF_WALK, st);
        mtf_walk_kernel<true><<<walk_grid, kMtfThreads, 0, st>>>(d_src, n, lists, nblocks, d_pos);
    }
    return cuda_status(cudaGetLastError());
}

int mtf_resolve(uint8_t *d_buf, const unsigned long long *d_len, const int32_t *d_mode, cudaStream_t st) {
    LaunchScope ls(DC_K_MTF_RESOLVE, st);
    mtf_resolve_kernel<<<1, 32, 0, st>>>(d_buf, d_len, d_mode);
    return cuda_status(cudaGetLastError());
}

}  // namespace dc

using namespace dc;

template <bool COMPRESS>
static int text_batch(const uint8_t *d_src, const uint64_t *d_src_off, size_t count, int modify, uint8_t *d_dst, const uint64_t *d_dst_off,
                      uint64_t *d_out_len, int32_t *d_status, void *stream) {
    if (count && (!d_src_off || !d_dst_off || !d_out_len || !d_src || !d_dst)) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (count == 0) return DC_OK;
    LaunchScope ls(DC_K_TEXT_BATCH, st);
    text_batch_kernel<COMPRESS><<<(unsigned int)((count + 127) / 128), 128, 0, st>>>(d_src, (const unsigned long long *)d_src_off, count, modify, d_dst,
                                                                                    (const unsigned long long *)d_dst_off,
                                                                                    (unsigned long long *)d_out_len, d_status);
    return cuda_status(cudaGetLastError());
}

extern "C" int dc_nybble_text_compress_batch(const uint8_t *d_src, const uint64_t *d_src_off, size_t count, int modify, uint8_t *d_dst,
                                             const uint64_t *d_dst_off, uint64_t *d_out_len, int32_t *d_status, void *stream) {
    return text_batch<true>(d_src, d_src_off, count, modify, d_dst, d_dst_off, d_out_len, d_status, stream);
}

extern "C" int dc_nybble_text_decompress_batch(const uint8_t *d_src, const uint64_t *d_src_off, size_t count, int modify, uint8_t *d_dst,
                                               const uint64_t *d_dst_off, uint64_t *d_out_len, int32_t *d_status, void *stream) {
    return text_batch<false>(d_src, d_src_off, count, modify, d_dst, d_dst_off, d_out_len, d_status, stream);
}
