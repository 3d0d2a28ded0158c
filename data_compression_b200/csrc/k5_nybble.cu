// k5_nybble.cu -- K5: nibble pack / unpack as pure streaming kernels (HBM-bound, 1.5 N bytes of traffic).
//
// Replaces the stream form of write_nybble() (nybble_compression.c:1091-1114, #else branches: offset 0 is
// the HIGH nibble) and the decoder's split (nybble_compression.c:767-769: hi first, then lo).
//
// pack  : thread reads 2 x 16 B (32 symbols, one per byte) and writes 16 B (32 nibbles)
// unpack: thread reads 16 B and writes 2 x 16 B
// 4 independent vectors per thread per loop trip keep >= 8 x 16 B loads in flight per thread.
#include "dc_common.cuh"

namespace dc {

constexpr int kNybThreads = 256;
constexpr int kNybUnroll = 4;

// bytes (s0,s1,s2,s3) of w -> byte0 = s0<<4|s1, byte2 = s2<<4|s3
__device__ __forceinline__ uint32_t pack_word(uint32_t w) { return ((w & 0x000F000Fu) << 4) | ((w >> 8) & 0x000F000Fu); }
__device__ __forceinline__ uint32_t pack_pair(uint32_t w0, uint32_t w1) {
    return __byte_perm(pack_word(w0), pack_word(w1), 0x6420);
}
// packed bytes (b_lo, b_hi) selected by `sel` -> 4 symbols: b_lo>>4, b_lo&15, b_hi>>4, b_hi&15
__device__ __forceinline__ uint32_t unpack_half(uint32_t p, uint32_t sel) {
    const uint32_t d = __byte_perm(p, 0, sel);
    return ((d >> 4) & 0x000F000Fu) | (d & 0x0F000F00u);
}

__global__ void __launch_bounds__(kNybThreads) nybble_pack_kernel(const uint4 *__restrict__ sym, size_t nvec_out,
                                                                  uint4 *__restrict__ packed,
                                                                  int32_t *__restrict__ d_status) {
    const size_t stride = (size_t)gridDim.x * kNybThreads;
    uint32_t bad = 0;
    size_t i = (size_t)blockIdx.x * kNybThreads + threadIdx.x;
    // main loop: kNybUnroll output vectors per trip, all loads issued before the first store
    for (; i + (kNybUnroll - 1) * stride < nvec_out; i += kNybUnroll * stride) {
        uint4 a[kNybUnroll], b[kNybUnroll];
#pragma unroll
        for (int u = 0; u < kNybUnroll; u++) {
            a[u] = ldg_stream(sym + 2 * (i + u * stride));
            b[u] = ldg_stream(sym + 2 * (i + u * stride) + 1);
        }
#pragma unroll
        for (int u = 0; u < kNybUnroll; u++) {
            bad |= (a[u].x | a[u].y | a[u].z | a[u].w | b[u].x | b[u].y | b[u].z | b[u].w);
            uint4 o;
            o.x = pack_pair(a[u].x, a[u].y);
            o.y = pack_pair(a[u].z, a[u].w);
            o.z = pack_pair(b[u].x, b[u].y);
            o.w = pack_pair(b[u].z, b[u].w);
            stg_stream(packed + i + u * stride, o);
        }
    }
    for (; i < nvec_out; i += stride) {
        const uint4 a = ldg_stream(sym + 2 * i), b = ldg_stream(sym + 2 * i + 1);
        bad |= (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w);
        uint4 o;
        o.x = pack_pair(a.x, a.y);
        o.y = pack_pair(a.z, a.w);
        o.z = pack_pair(b.x, b.y);
        o.w = pack_pair(b.z, b.w);
        stg_stream(packed + i, o);
    }
    if (bad & 0xF0F0F0F0u) set_status(d_status, DC_ERR_SYMBOL);
}

// byte-granular path: ragged tail (< 32 symbols) and unaligned buffers; one output byte per thread
__global__ void nybble_pack_bytes_kernel(const uint8_t *__restrict__ sym, size_t sym_begin, size_t n_sym,
                                         uint8_t *__restrict__ packed, int32_t *__restrict__ d_status) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t s = sym_begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); s < n_sym; s += 2 * stride) {
        const uint32_t hi = sym[s], lo = (s + 1 < n_sym) ? sym[s + 1] : 0u;
        if ((hi | lo) & 0xF0u) set_status(d_status, DC_ERR_SYMBOL);
        packed[s >> 1] = (uint8_t)(((hi & 0xFu) << 4) | (lo & 0xFu));
    }
}

// one 32-byte store: a thread's two output vectors are contiguous, and two separate 16-byte stores would each write half
// of every 32-byte sector the warp touches
__device__ __forceinline__ void stg_stream_256(uint4 *p, const uint4 &a, const uint4 &b) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z),
                 "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

template <bool WIDE>  // WIDE: the symbol buffer is 32-byte aligned
__global__ void __launch_bounds__(kNybThreads) nybble_unpack_kernel(const uint4 *__restrict__ packed, size_t nvec_in,
                                                                    uint4 *__restrict__ sym) {
    const size_t stride = (size_t)gridDim.x * kNybThreads;
    size_t i = (size_t)blockIdx.x * kNybThreads + threadIdx.x;
    for (; i + (kNybUnroll - 1) * stride < nvec_in; i += kNybUnroll * stride) {
        uint4 p[kNybUnroll];
#pragma unroll
        for (int u = 0; u < kNybUnroll; u++) p[u] = ldg_stream(packed + i + u * stride);
#pragma unroll
        for (int u = 0; u < kNybUnroll; u++) {
            uint4 lo, hi;
            lo.x = unpack_half(p[u].x, 0x1100);
            lo.y = unpack_half(p[u].x, 0x3322);
            lo.z = unpack_half(p[u].y, 0x1100);
            lo.w = unpack_half(p[u].y, 0x3322);
            hi.x = unpack_half(p[u].z, 0x1100);
            hi.y = unpack_half(p[u].z, 0x3322);
            hi.z = unpack_half(p[u].w, 0x1100);
            hi.w = unpack_half(p[u].w, 0x3322);
            if (WIDE) {
                stg_stream_256(sym + 2 * (i + u * stride), lo, hi);
            } else {
                stg_stream(sym + 2 * (i + u * stride), lo);
                stg_stream(sym + 2 * (i + u * stride) + 1, hi);
            }
        }
    }
    for (; i < nvec_in; i += stride) {
        const uint4 p = ldg_stream(packed + i);
        uint4 lo, hi;
        lo.x = unpack_half(p.x, 0x1100);
        lo.y = unpack_half(p.x, 0x3322);
        lo.z = unpack_half(p.y, 0x1100);
        lo.w = unpack_half(p.y, 0x3322);
        hi.x = unpack_half(p.z, 0x1100);
        hi.y = unpack_half(p.z, 0x3322);
        hi.z = unpack_half(p.w, 0x1100);
        hi.w = unpack_half(p.w, 0x3322);
        stg_stream(sym + 2 * i, lo);
        stg_stream(sym + 2 * i + 1, hi);
    }
}

__global__ void nybble_unpack_bytes_kernel(const uint8_t *__restrict__ packed, size_t sym_begin, size_t n_sym,
                                           uint8_t *__restrict__ sym) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t s = sym_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sym; s += stride) {
        const uint32_t b = packed[s >> 1];
        sym[s] = (uint8_t)((s & 1) ? (b & 0xFu) : (b >> 4));
    }
}

static int stream_grid(size_t items_per_thread_total) {
    const size_t want = (items_per_thread_total + kNybThreads - 1) / kNybThreads;
    const size_t cap = (size_t)sm_count() * 16;  // 16 resident CTAs of 256 threads = 2 waves of 8 per SM
    return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace dc

using namespace dc;

// Unaligned buffers.  Symbols 2i, 2i + 1 live in packed byte i, so peeling an even head of h symbols moves the symbol pointer
// by h and the packed pointer by h / 2: both become 16-byte aligned iff sym == 2 * packed (mod 16) -- which is what a shard
// of an aligned stream looks like (symbol offset lo, byte offset lo / 2).  Then only the head and the tail go through the
// byte kernels; any other combination of misalignments has no common vector grid and takes the byte kernel as a whole.
static size_t nybble_head(const uint8_t *sym, const uint8_t *packed, size_t n_sym, bool *vector_ok) {
    const uintptr_t s = (uintptr_t)sym, p = (uintptr_t)packed;
    *vector_ok = ((s - 2 * p) & 15) == 0;
    if (!*vector_ok) return 0;
    const size_t h = (size_t)((32 - ((2 * p) & 31)) & 31);
    return h < n_sym ? h : n_sym;
}

extern "C" int dc_nybble_pack(const uint8_t *d_sym, size_t n_sym, uint8_t *d_packed, int32_t *d_status, void *stream) {
    if ((!d_sym || !d_packed) && n_sym) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (n_sym == 0) return DC_OK;
    bool vector_ok;
    const size_t head = nybble_head(d_sym, d_packed, n_sym, &vector_ok);
    const size_t nvec = vector_ok ? (n_sym - head) / 32 : 0;
    if (head) {
        LaunchScope ls(DC_K_NYBBLE_TAIL, st);
        nybble_pack_bytes_kernel<<<1, kNybThreads, 0, st>>>(d_sym, 0, head, d_packed, d_status);
    }
    if (nvec) {
        LaunchScope ls(DC_K_NYBBLE_PACK, st);
        nybble_pack_kernel<<<stream_grid(nvec), kNybThreads, 0, st>>>((const uint4 *)(d_sym + head), nvec, (uint4 *)(d_packed + head / 2), d_status);
    }
    const size_t done = head + nvec * 32;
    if (done < n_sym) {
        const size_t bytes = (n_sym - done + 1) / 2;
        LaunchScope ls(DC_K_NYBBLE_TAIL, st);
        nybble_pack_bytes_kernel<<<stream_grid(bytes), kNybThreads, 0, st>>>(d_sym, done, n_sym, d_packed, d_status);
    }
    return cuda_status(cudaGetLastError());
}

extern "C" int dc_nybble_unpack(const uint8_t *d_packed, size_t n_sym, uint8_t *d_sym, void *stream) {
    if ((!d_sym || !d_packed) && n_sym) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_sym == 0) return DC_OK;
    bool vector_ok;
    const size_t head = nybble_head(d_sym, d_packed, n_sym, &vector_ok);
    const size_t nvec = vector_ok ? (n_sym - head) / 32 : 0;
    if (head) {
        LaunchScope ls(DC_K_NYBBLE_TAIL, st);
        nybble_unpack_bytes_kernel<<<1, kNybThreads, 0, st>>>(d_packed, 0, head, d_sym);
    }
    if (nvec) {
        LaunchScope ls(DC_K_NYBBLE_UNPACK, st);
        const uint4 *pv = (const uint4 *)(d_packed + head / 2);
        uint4 *sv = (uint4 *)(d_sym + head);
        if (((uintptr_t)sv & 31) == 0) nybble_unpack_kernel<true><<<stream_grid(nvec), kNybThreads, 0, st>>>(pv, nvec, sv);
        else nybble_unpack_kernel<false><<<stream_grid(nvec), kNybThreads, 0, st>>>(pv, nvec, sv);
    }
    const size_t done = head + nvec * 32;
    if (done < n_sym) {
        LaunchScope ls(DC_K_NYBBLE_TAIL, st);
        nybble_unpack_bytes_kernel<<<stream_grid(n_sym - done), kNybThreads, 0, st>>>(d_packed, done, n_sym, d_sym);
    }
    return cuda_status(cudaGetLastError());
}
