// k7_trits.cu -- K7: the digit packings of SURVEY 8f row N4: the 5-trits-per-byte payload of radix 3, and (at the end of the
// file) the base64url text form of the binary payload.
//
// The reference's default radix is 3 (n_ary_huffman.c:2529) and its author sketches the storage at :745-748: "grab 5
// trits at a time, convert into a number 1..243, and store as an 8-bit octet (which never uses byte 0 or 244..255)".
// Layout: codes as base-3 numerals, most significant trit first, concatenated in input order; every 5 trits
// t0..t4 -> the byte 1 + t0*81 + t1*27 + t2*9 + t3*3 + t4; the last group is padded with zero trits.
//
// The encode / decode kernels (K3, K4) work on bit fields, so for n = 3 they run on an intermediate "T2" stream with one
// 2-bit field per trit (4 trits per byte, MSB first; tab->packed_radix == 3).  The two streaming kernels here convert
// between the T2 stream and the payload: a thread handles 80 trits = 20 bytes of T2 = 16 bytes of payload, through a
// 1024-entry (pack) / 256-entry (unpack) shared-memory table.  Traffic 2.25 bytes per payload byte.
#include "dc_common.cuh"

namespace dc {

constexpr int kTritThreads = 256;
constexpr int kTritsPerThread = 80;

// 160 bits as five big-endian words: bits [bit, bit + 10)
__device__ __forceinline__ uint32_t ten_bits(const uint32_t (&w)[6], int bit) {
    const int a = bit >> 5, s = bit & 31;
    return __funnelshift_l(w[a + 1], w[a], s) >> 22;
}

__global__ void __launch_bounds__(kTritThreads) trit_pack_kernel(const uint8_t *__restrict__ t2, unsigned long long ntrits,
                                                                 uint8_t *__restrict__ out, int32_t *__restrict__ d_status) {
    __shared__ uint8_t s_lut[1024];  // five 2-bit fields -> byte; 0 = some field is 3 (not a trit)
    for (int x = threadIdx.x; x < 1024; x += kTritThreads) {
        uint32_t v = 1;
        bool ok = true;
        const uint32_t p3[5] = {81, 27, 9, 3, 1};
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint32_t d = (x >> (8 - 2 * k)) & 3u;
            ok &= d != 3u;
            v += d * p3[k];
        }
        s_lut[x] = ok ? (uint8_t)v : 0;
    }
    __syncthreads();
    const unsigned long long t2_bytes = (2 * ntrits + 7) / 8, out_bytes = (ntrits + 4) / 5;
    const unsigned long long nthreads_needed = (ntrits + kTritsPerThread - 1) / kTritsPerThread;
    bool bad = false;
    // the next group's five words are loaded while this one is converted (the kernel waits on memory otherwise)
    const unsigned long long stride = (unsigned long long)gridDim.x * kTritThreads;
    // (a group that holds the last trit takes the byte-wise path even if its 20 bytes exist: the bits behind the last trit are
    // padding -- whatever the caller's last byte holds there must not become a trit)
    auto whole = [&](unsigned long long t) { return t < nthreads_needed && (t + 1) * kTritsPerThread <= ntrits; };
    unsigned long long t = (unsigned long long)blockIdx.x * kTritThreads + threadIdx.x;
    uint32_t nx[5] = {0, 0, 0, 0, 0};
    bool nx_whole = whole(t);
    if (nx_whole) {
#pragma unroll
        for (int k = 0; k < 5; k++) nx[k] = __ldg((const uint32_t *)(t2 + t * 20) + k);
    }
    for (; t < nthreads_needed; t += stride) {
        const unsigned long long ib = t * 20, ob = t * 16;
        uint32_t w[6];
        const bool cur_whole = nx_whole;
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = bswap32(nx[k]);
        nx_whole = whole(t + stride);
        if (nx_whole) {
#pragma unroll
            for (int k = 0; k < 5; k++) nx[k] = __ldg((const uint32_t *)(t2 + (t + stride) * 20) + k);
        }
        if (!cur_whole) {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) x = (x << 8) | (ib + 4 * k + b < t2_bytes ? (uint32_t)t2[ib + 4 * k + b] : 0u);
                w[k] = x;
            }
            // bits behind the last trit are padding: zero trits
            const unsigned long long valid_bits = 2 * ntrits - ib * 8;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const long long keep = (long long)valid_bits - 32 * k;
                if (keep <= 0) w[k] = 0;
                else if (keep < 32) w[k] &= ~(0xFFFFFFFFu >> keep);
            }
        }
        w[5] = 0;
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < 16; g++) {
            const uint32_t b = s_lut[ten_bits(w, 10 * g)];
            bad |= b == 0;
            o[g >> 2] |= b << (8 * (g & 3));
        }
        if (ob + 16 <= out_bytes) {
            stg_stream((uint4 *)(out + ob), make_uint4(o[0], o[1], o[2], o[3]));
        } else {
            for (int g = 0; g < 16; g++)
                if (ob + g < out_bytes) out[ob + g] = (uint8_t)(o[g >> 2] >> (8 * (g & 3)));
        }
    }
    if (bad) set_status(d_status, DC_ERR_CORRUPT);
}

__global__ void __launch_bounds__(kTritThreads) trit_unpack_kernel(const uint8_t *__restrict__ packed, unsigned long long ntrits,
                                                                   uint8_t *__restrict__ t2, int32_t *__restrict__ d_status) {
    __shared__ uint16_t s_lut[256];  // byte -> five 2-bit fields; 0xFFFF = not a payload byte (0 or 244..255)
    for (int b = threadIdx.x; b < 256; b += kTritThreads) {
        uint32_t v = b >= 1 && b <= 243 ? (uint32_t)(b - 1) : 0xFFFFu, x = 0;
        if (v != 0xFFFFu) {
            const uint32_t p3[5] = {81, 27, 9, 3, 1};
#pragma unroll
            for (int k = 0; k < 5; k++) { x = (x << 2) | ((v / p3[k]) % 3u); }
            v = x;
        }
        s_lut[b] = (uint16_t)v;
    }
    __syncthreads();
    const unsigned long long in_bytes = (ntrits + 4) / 5, t2_bytes = (2 * ntrits + 7) / 8;
    const unsigned long long nthreads_needed = (ntrits + kTritsPerThread - 1) / kTritsPerThread;
    bool bad = false;
    const unsigned long long stride = (unsigned long long)gridDim.x * kTritThreads;
    auto whole = [&](unsigned long long t) { return t < nthreads_needed && t * 16 + 16 <= in_bytes && ((uintptr_t)(packed + t * 16) & 15) == 0; };
    unsigned long long t = (unsigned long long)blockIdx.x * kTritThreads + threadIdx.x;
    uint4 nx = make_uint4(0, 0, 0, 0);
    bool nx_whole = whole(t);
    if (nx_whole) nx = ldg_stream((const uint4 *)(packed + t * 16));
    for (; t < nthreads_needed; t += stride) {
        const unsigned long long ib = t * 16, ob = t * 20;
        uint32_t in[4] = {nx.x, nx.y, nx.z, nx.w};
        const bool cur_whole = nx_whole;
        nx_whole = whole(t + stride);
        if (nx_whole) nx = ldg_stream((const uint4 *)(packed + (t + stride) * 16));
        if (!cur_whole) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) x |= (ib + 4 * k + b < in_bytes ? (uint32_t)packed[ib + 4 * k + b] : 1u) << (8 * b);  // 1 = five zero trits
                in[k] = x;
            }
        }
        // 16 groups of 10 bits -> 160 bits, big-endian
        uint32_t w[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int g = 0; g < 16; g++) {
            const uint32_t f = s_lut[(in[g >> 2] >> (8 * (g & 3))) & 0xFFu];
            bad |= f == 0xFFFFu;
            const int bit = 10 * g, a = bit >> 5, s = bit & 31;  // field occupies bits [bit, bit + 10)
            const uint32_t x = f & 0x3FFu;
            if (s <= 22) w[a] |= x << (22 - s);
            else { w[a] |= x >> (s - 22); w[a + 1] |= x << (54 - s); }
        }
        if (ob + 20 <= t2_bytes) {
#pragma unroll
            for (int k = 0; k < 5; k++) ((uint32_t *)(t2 + ob))[k] = bswap32(w[k]);
        } else {
            for (int k = 0; k < 20; k++)
                if (ob + k < t2_bytes) t2[ob + k] = (uint8_t)(w[k >> 2] >> (24 - 8 * (k & 3)));
        }
    }
    if (bad) set_status(d_status, DC_ERR_CORRUPT);
}

// ---------------------------------------------------------------------------------------- base64url text form (n = 2)
// The reference's unfinished packer emits the binary code as base64url characters, 6 bits each, through int2digit()
// (n_ary_huffman.c:371-426, :1646-1671).  Layout: character k = int2digit(bits [6k, 6k + 6) of the payload, most significant
// first, zero padded) -- RFC 4648 without '=' padding.  A thread turns 12 bytes into 16 characters and back.
constexpr int kB64Threads = 256;

__device__ __forceinline__ uint32_t b64_char(uint32_t v) {  // int2digit() :371-426
    return v < 26 ? 'A' + v : v < 52 ? 'a' + (v - 26) : v < 62 ? '0' + (v - 52) : v == 62 ? '-' : '_';
}

__global__ void __launch_bounds__(kB64Threads) b64_pack_kernel(const uint8_t *__restrict__ bits, unsigned long long nbits,
                                                               uint8_t *__restrict__ chars) {
    __shared__ uint8_t s_lut[64];
    if (threadIdx.x < 64) s_lut[threadIdx.x] = (uint8_t)b64_char(threadIdx.x);
    __syncthreads();
    const unsigned long long nbytes = (nbits + 7) / 8, nchars = (nbits + 5) / 6, groups = (nchars + 15) / 16;
    for (unsigned long long g = (unsigned long long)blockIdx.x * kB64Threads + threadIdx.x; g < groups;
         g += (unsigned long long)gridDim.x * kB64Threads) {
        const unsigned long long ib = g * 12, ob = g * 16;
        uint32_t w[4] = {0, 0, 0, 0};
        if (ib + 12 <= nbytes) {
#pragma unroll
            for (int k = 0; k < 3; k++) w[k] = bswap32(__ldg((const uint32_t *)(bits + ib) + k));
        } else {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) x = (x << 8) | (ib + 4 * k + b < nbytes ? (uint32_t)bits[ib + 4 * k + b] : 0u);
                w[k] = x;
            }
        }
        // bits behind the end of the stream are padding: zero (the last byte of the caller's buffer may hold anything there)
        {
            const unsigned long long first = ib * 8;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const long long keep = (long long)nbits - (long long)(first + 32 * k);
                if (keep <= 0) w[k] = 0;
                else if (keep < 32) w[k] &= ~(0xFFFFFFFFu >> keep);
            }
        }
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const int bit = 6 * c, a = bit >> 5, sft = bit & 31;
            const uint32_t v = __funnelshift_l(w[a + 1], w[a], sft) >> 26;
            o[c >> 2] |= (uint32_t)s_lut[v] << (8 * (c & 3));
        }
        if (ob + 16 <= nchars) {
            stg_stream((uint4 *)(chars + ob), make_uint4(o[0], o[1], o[2], o[3]));
        } else {
            for (int c = 0; c < 16; c++)
                if (ob + c < nchars) chars[ob + c] = (uint8_t)(o[c >> 2] >> (8 * (c & 3)));
        }
    }
}

__global__ void __launch_bounds__(kB64Threads) b64_unpack_kernel(const uint8_t *__restrict__ chars, unsigned long long nbits,
                                                                 uint8_t *__restrict__ bits, int32_t *__restrict__ d_status) {
    __shared__ uint8_t s_lut[256];   // digit2int() :428-455: both alphabets for 62 / 63; 0xFF = not a digit
    for (int c = threadIdx.x; c < 256; c += kB64Threads) {
        uint32_t v = 0xFF;
        for (uint32_t i = 0; i < 64; i++)
            if (b64_char(i) == (uint32_t)c) v = i;
        if (c == '+') v = 62;
        if (c == '/') v = 63;
        s_lut[c] = (uint8_t)v;
    }
    __syncthreads();
    const unsigned long long nbytes = (nbits + 7) / 8, nchars = (nbits + 5) / 6, groups = (nchars + 15) / 16;
    bool bad = false;
    for (unsigned long long g = (unsigned long long)blockIdx.x * kB64Threads + threadIdx.x; g < groups;
         g += (unsigned long long)gridDim.x * kB64Threads) {
        const unsigned long long ib = g * 16, ob = g * 12;
        uint32_t in[4];
        if (ib + 16 <= nchars && ((uintptr_t)(chars + ib) & 15) == 0) {
            const uint4 v = ldg_stream((const uint4 *)(chars + ib));
            in[0] = v.x; in[1] = v.y; in[2] = v.z; in[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) x |= (ib + 4 * k + b < nchars ? (uint32_t)chars[ib + 4 * k + b] : (uint32_t)'A') << (8 * b);
                in[k] = x;
            }
        }
        uint32_t w[3] = {0, 0, 0};
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const uint32_t v = s_lut[(in[c >> 2] >> (8 * (c & 3))) & 0xFFu];
            bad |= v == 0xFFu;
            const int bit = 6 * c, a = bit >> 5, sft = bit & 31;
            const uint32_t x = v & 63u;
            if (sft <= 26) w[a] |= x << (26 - sft);
            else { w[a] |= x >> (sft - 26); w[a + 1] |= x << (58 - sft); }
        }
        if (ob + 12 <= nbytes) {
#pragma unroll
            for (int k = 0; k < 3; k++) ((uint32_t *)(bits + ob))[k] = bswap32(w[k]);
        } else {
            for (int k = 0; k < 12; k++)
                if (ob + k < nbytes) bits[ob + k] = (uint8_t)(w[k >> 2] >> (24 - 8 * (k & 3)));
        }
    }
    if (bad) set_status(d_status, DC_ERR_CORRUPT);
}

}  // namespace dc

using namespace dc;

static int trit_grid(unsigned long long ntrits) {
    const unsigned long long want = ((ntrits + kTritsPerThread - 1) / kTritsPerThread + kTritThreads - 1) / kTritThreads;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

extern "C" int dc_trit_pack(const uint8_t *d_t2, uint64_t ntrits, uint8_t *d_payload, int32_t *d_status, void *stream) {
    if (ntrits && (!d_t2 || !d_payload || ((uintptr_t)d_t2 & 3) || ((uintptr_t)d_payload & 15))) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (ntrits == 0) return DC_OK;
    LaunchScope ls(DC_K_TRIT_PACK, st);
    trit_pack_kernel<<<trit_grid(ntrits), kTritThreads, 0, st>>>(d_t2, ntrits, d_payload, d_status);
    return cuda_status(cudaGetLastError());
}

// the launch alone: *d_status is only ever set (the pipelined host decompress shares one status word over all chunks)
int dc::trit_unpack_launch(const uint8_t *d_payload, unsigned long long ntrits, uint8_t *d_t2, int32_t *d_status, cudaStream_t st) {
    if (ntrits == 0) return DC_OK;
    LaunchScope ls(DC_K_TRIT_UNPACK, st);
    trit_unpack_kernel<<<trit_grid(ntrits), kTritThreads, 0, st>>>(d_payload, ntrits, d_t2, d_status);
    return cuda_status(cudaGetLastError());
}

extern "C" int dc_trit_unpack(const uint8_t *d_payload, uint64_t ntrits, uint8_t *d_t2, int32_t *d_status, void *stream) {
    if (ntrits && (!d_t2 || !d_payload || ((uintptr_t)d_t2 & 3))) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    return trit_unpack_launch(d_payload, ntrits, d_t2, d_status, st);
}

static int b64_grid(unsigned long long nbits) {
    const unsigned long long groups = ((nbits + 5) / 6 + 15) / 16, want = (groups + kB64Threads - 1) / kB64Threads;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

extern "C" int dc_base64url_pack(const uint8_t *d_bits, uint64_t nbits, uint8_t *d_chars, void *stream) {
    if (nbits && (!d_bits || !d_chars || ((uintptr_t)d_bits & 3) || ((uintptr_t)d_chars & 15))) return DC_ERR_ARG;
    if (nbits == 0) return DC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    LaunchScope ls(DC_K_B64_PACK, st);
    b64_pack_kernel<<<b64_grid(nbits), kB64Threads, 0, st>>>(d_bits, nbits, d_chars);
    return cuda_status(cudaGetLastError());
}

extern "C" int dc_base64url_unpack(const uint8_t *d_chars, uint64_t nbits, uint8_t *d_bits, int32_t *d_status, void *stream) {
    if (nbits && (!d_bits || !d_chars || ((uintptr_t)d_bits & 3))) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (nbits == 0) return DC_OK;
    LaunchScope ls(DC_K_B64_UNPACK, st);
    b64_unpack_kernel<<<b64_grid(nbits), kB64Threads, 0, st>>>(d_chars, nbits, d_bits, d_status);
    return cuda_status(cudaGetLastError());
}
