// k1_histogram.cu -- K1: byte histogram with warp-privatised shared-memory counters.
//
// Replaces histogram() (n_ary_huffman.c:461-493): h[b]++ for every input byte, after zeroing h[0..258]
// (:474-476).  Counts are 64-bit (the reference's int overflows at 2^31).
//
// The kernel is bound by the shared-memory update rate, not by HBM, so three counter layouts are kept
// and selected at run time (DC_HIST_VARIANT, default chosen from measurements in profiles/):
//   A<R>  per-warp u32 histograms, R lane-interleaved copies, shared-memory atomics
//   B     per-LANE byte counters (lane == bank, so updates never conflict), plain load/add/store,
//         folded into registers every 240 bytes per lane
//   C     same layout as B, updated with one shared-memory atomic add of 1 << 8*(b&3)
#include <stdlib.h>

#include "dc_common.cuh"

namespace dc {

// ------------------------------------------------------------------------------------------ variant A

template <int R>
__global__ void __launch_bounds__(256) hist_warp_atomic_kernel(const uint8_t *__restrict__ in, size_t n,
                                                               unsigned long long *__restrict__ hist) {
    constexpr int kWarps = 8;
    extern __shared__ uint32_t smem[];  // [kWarps][256 * R]
    for (int i = threadIdx.x; i < kWarps * 256 * R; i += 256) smem[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wh = smem + warp * 256 * R + (lane & (R - 1));

    // unaligned head / ragged tail: a handful of bytes, counted by CTA 0
    const size_t head = min((size_t)((16 - ((uintptr_t)in & 15)) & 15), n);
    const size_t nvec = (n - head) / 16;
    const uint4 *vin = (const uint4 *)(in + head);
    if (blockIdx.x == 0) {
        for (size_t i = threadIdx.x; i < head; i += 256) atomicAdd(&wh[in[i] * R], 1u);
        for (size_t i = head + nvec * 16 + threadIdx.x; i < n; i += 256) atomicAdd(&wh[in[i] * R], 1u);
    }

    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    constexpr int U = 4;
#define DC_COUNT_WORD(w)                            \
    atomicAdd(&wh[((w) & 0xFFu) * R], 1u);          \
    atomicAdd(&wh[(((w) >> 8) & 0xFFu) * R], 1u);   \
    atomicAdd(&wh[(((w) >> 16) & 0xFFu) * R], 1u);  \
    atomicAdd(&wh[((w) >> 24) * R], 1u);
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ldg_stream(vin + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) {
            DC_COUNT_WORD(v[u].x) DC_COUNT_WORD(v[u].y) DC_COUNT_WORD(v[u].z) DC_COUNT_WORD(v[u].w)
        }
    }
    for (; i < nvec; i += stride) {
        const uint4 v = ldg_stream(vin + i);
        DC_COUNT_WORD(v.x) DC_COUNT_WORD(v.y) DC_COUNT_WORD(v.z) DC_COUNT_WORD(v.w)
    }
#undef DC_COUNT_WORD
    __syncthreads();
    // one global atomic per bin per CTA
    for (int b = threadIdx.x; b < 256; b += 256) {
        unsigned long long s = 0;
        for (int w = 0; w < kWarps; w++)
#pragma unroll
            for (int r = 0; r < R; r++) s += smem[w * 256 * R + b * R + r];
        if (s) atomicAdd(&hist[b], s);
    }
}

// ------------------------------------------------------------------------------------------ variants B, C

constexpr int kLpWarps = 4;                 // warps per CTA
constexpr int kLpThreads = kLpWarps * 32;
constexpr int kLpVecPerPeriod = 15;         // 15 x 16 B = 240 bytes per lane between folds (< 256)

// fold the warp's 32 lane-private byte-counter columns into per-lane u32 registers and clear them.
// lane L owns rows L and L+32 (bins 4L..4L+3 and 128+4L..128+4L+3); reads are rotated so that every
// lane hits a different bank.
__device__ __forceinline__ void fold_counters(uint32_t *wc, int lane, uint32_t acc[8]) {
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t *row = wc + (lane + 32 * half) * 32;
        uint32_t even = 0, odd = 0;  // 2 x 16-bit fields each; 32 copies x 255 < 65536
#pragma unroll 8
        for (int k = 0; k < 32; k++) {
            const int c = (lane + k) & 31;
            const uint32_t w = row[c];
            row[c] = 0;
            even += w & 0x00FF00FFu;
            odd += (w >> 8) & 0x00FF00FFu;
        }
        acc[half * 4 + 0] += even & 0xFFFFu;
        acc[half * 4 + 1] += odd & 0xFFFFu;
        acc[half * 4 + 2] += even >> 16;
        acc[half * 4 + 3] += odd >> 16;
    }
    __syncwarp();
}

template <bool ATOMIC>
__device__ __forceinline__ void count_word(uint32_t *col, uint8_t *colb, uint32_t w) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t b = (w >> (8 * k)) & 0xFFu;
        if (ATOMIC) {
            atomicAdd(col + (b >> 2) * 32, 1u << ((b & 3u) * 8));
        } else {
            uint8_t *p = colb + (b >> 2) * 128 + (b & 3u);
            *p = (uint8_t)(*p + 1);
        }
    }
}

template <bool ATOMIC>
__global__ void __launch_bounds__(kLpThreads) hist_lane_private_kernel(const uint8_t *__restrict__ in, size_t n,
                                                                       unsigned long long *__restrict__ hist) {
    __shared__ uint32_t cnt[kLpWarps][64 * 32];
    __shared__ unsigned long long total[256];
    for (int i = threadIdx.x; i < kLpWarps * 64 * 32; i += kLpThreads) (&cnt[0][0])[i] = 0;
    for (int i = threadIdx.x; i < 256; i += kLpThreads) total[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wc = cnt[warp];
    uint32_t *col = wc + lane;                 // word (row, lane): bank == lane
    uint8_t *colb = (uint8_t *)(wc + lane);
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    const size_t head = min((size_t)((16 - ((uintptr_t)in & 15)) & 15), n);
    const size_t nvec = (n - head) / 16;
    const uint4 *vin = (const uint4 *)(in + head);
    if (blockIdx.x == 0 && warp == 0) {
        for (size_t i = lane; i < head; i += 32) atomicAdd(&total[in[i]], 1ull);
        for (size_t i = head + nvec * 16 + lane; i < n; i += 32) atomicAdd(&total[in[i]], 1ull);
    }

    // a period = kLpVecPerPeriod vectors per lane = 480 consecutive vectors per warp
    constexpr size_t kPeriodVec = (size_t)kLpVecPerPeriod * 32;
    const size_t nperiods = (nvec + kPeriodVec - 1) / kPeriodVec;
    const size_t gwarp = (size_t)blockIdx.x * kLpWarps + warp, nwarps = (size_t)gridDim.x * kLpWarps;
    for (size_t p = gwarp; p < nperiods; p += nwarps) {
        const size_t base = p * kPeriodVec + lane;
#pragma unroll
        for (int g = 0; g < kLpVecPerPeriod; g += 5) {
            uint4 v[5];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const size_t idx = base + (size_t)(g + j) * 32;
                v[j] = idx < nvec ? ldg_stream(vin + idx) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const size_t idx = base + (size_t)(g + j) * 32;
                if (idx < nvec) {
                    count_word<ATOMIC>(col, colb, v[j].x);
                    count_word<ATOMIC>(col, colb, v[j].y);
                    count_word<ATOMIC>(col, colb, v[j].z);
                    count_word<ATOMIC>(col, colb, v[j].w);
                }
            }
        }
        fold_counters(wc, lane, acc);
    }
    // per-lane registers -> CTA totals -> one global atomic per bin per CTA
#pragma unroll
    for (int half = 0; half < 2; half++)
#pragma unroll
        for (int f = 0; f < 4; f++)
            if (acc[half * 4 + f]) atomicAdd(&total[(lane + 32 * half) * 4 + f], (unsigned long long)acc[half * 4 + f]);
    __syncthreads();
    for (int b = threadIdx.x; b < 256; b += kLpThreads)
        if (total[b]) atomicAdd(&hist[b], total[b]);
}

// ------------------------------------------------------------------------------------------ histogram + run histograms
// The encoder wants the bit offset of every 32 KB run before it starts (k3_encode.cu, "planned single pass"); that is
// sum(count[s] * length[s]) over the runs in front -- known as soon as the code lengths are, if the histogram pass keeps
// one small histogram per run.  Same counting as variant A<1>; a CTA walks whole runs, and after each run its 256
// threads fold the eight warp tables into 256 u16 counts (512 bytes per 32 KB of input: 1.6 % extra traffic).
constexpr int kHistRunBytes = 32768;   // == kRunBytes of k3_encode.cu
__global__ void __launch_bounds__(256) hist_runs_kernel(const uint8_t *__restrict__ in, size_t n, unsigned long long *__restrict__ hist,
                                                         uint16_t *__restrict__ run_hist, unsigned int nruns) {
    constexpr int kWarps = 8;
    __shared__ uint32_t smem[kWarps * 256];
    for (int i = threadIdx.x; i < kWarps * 256; i += 256) smem[i] = 0;
    __syncthreads();
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t *wh = smem + warp * 256;
    unsigned long long total = 0;
#define DC_COUNT_WORD(w)                      \
    atomicAdd(&wh[(w) & 0xFFu], 1u);          \
    atomicAdd(&wh[((w) >> 8) & 0xFFu], 1u);   \
    atomicAdd(&wh[((w) >> 16) & 0xFFu], 1u);  \
    atomicAdd(&wh[(w) >> 24], 1u);
    for (unsigned int run = blockIdx.x; run < nruns; run += gridDim.x) {
        const size_t base = (size_t)run * kHistRunBytes;
        const uint4 *vin = (const uint4 *)(in + base);
        if (base + kHistRunBytes <= n) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = ldg_stream(vin + tid + u * 256);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                DC_COUNT_WORD(v[u].x) DC_COUNT_WORD(v[u].y) DC_COUNT_WORD(v[u].z) DC_COUNT_WORD(v[u].w)
            }
        } else {  // the ragged last run
            const size_t len = n - base, nvec = len / 16;
            for (size_t i = tid; i < nvec; i += 256) {
                const uint4 v = ldg_stream(vin + i);
                DC_COUNT_WORD(v.x) DC_COUNT_WORD(v.y) DC_COUNT_WORD(v.z) DC_COUNT_WORD(v.w)
            }
            for (size_t i = nvec * 16 + tid; i < len; i += 256) atomicAdd(&wh[in[base + i]], 1u);
        }
        __syncthreads();
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            c += smem[w * 256 + tid];
            smem[w * 256 + tid] = 0;
        }
        run_hist[(size_t)run * 256 + tid] = (uint16_t)c;   // <= 32768
        total += c;
        __syncthreads();
    }
#undef DC_COUNT_WORD
    if (total) atomicAdd(&hist[tid], total);
}

int hist_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("DC_HIST_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

int launch_histogram(const uint8_t *d_in, size_t n, unsigned long long *d_hist, int variant, cudaStream_t st) {
    DC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, DC_NSLOTS * sizeof(unsigned long long), st));
    if (n == 0) return DC_OK;
    const int sms = sm_count();
    const size_t nvec = n / 16 + 1;
    LaunchScope ls(DC_K_HISTOGRAM, st);
    switch (variant) {
        case 1: case 2: case 3: {
            const int R = variant == 1 ? 2 : variant == 2 ? 4 : 8;
            const size_t smem = (size_t)8 * 256 * R * 4;
            const int per_sm = R == 8 ? 3 : 8;
            const int grid = (int)min((size_t)sms * per_sm, (nvec + 255) / 256);
            if (R == 2) hist_warp_atomic_kernel<2><<<grid, 256, smem, st>>>(d_in, n, d_hist);
            else if (R == 4) hist_warp_atomic_kernel<4><<<grid, 256, smem, st>>>(d_in, n, d_hist);
            else {
                DC_CUDA_TRY(ensure_dynamic_smem((const void *)hist_warp_atomic_kernel<8>, smem));
                hist_warp_atomic_kernel<8><<<grid, 256, smem, st>>>(d_in, n, d_hist);
            }
            break;
        }
        case 4: case 5: {
            const size_t nper = (nvec + kLpVecPerPeriod * 32 - 1) / (kLpVecPerPeriod * 32);
            const int grid = (int)min((size_t)sms * 6, (nper + kLpWarps - 1) / kLpWarps);
            if (variant == 4) hist_lane_private_kernel<false><<<grid, kLpThreads, 0, st>>>(d_in, n, d_hist);
            else hist_lane_private_kernel<true><<<grid, kLpThreads, 0, st>>>(d_in, n, d_hist);
            break;
        }
        default: {
            const int grid = (int)min((size_t)sms * 8, (nvec + 255) / 256);
            hist_warp_atomic_kernel<1><<<grid, 256, 8 * 256 * 4, st>>>(d_in, n, d_hist);
        }
    }
    return cuda_status(cudaGetLastError());
}

}  // namespace dc

extern "C" int dc_histogram_u8(const uint8_t *d_in, size_t n, uint64_t *d_hist, void *stream) {
    if (!d_hist || (!d_in && n)) return DC_ERR_ARG;
    return dc::launch_histogram(d_in, n, (unsigned long long *)d_hist, dc::hist_variant(), (cudaStream_t)stream);
}

namespace dc {
int launch_histogram_runs(const uint8_t *d_in, size_t n, unsigned long long *d_hist, uint16_t *d_run_hist, cudaStream_t st) {
    DC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, DC_NSLOTS * sizeof(unsigned long long), st));
    if (n == 0) return DC_OK;
    const size_t nruns = (n + kHistRunBytes - 1) / kHistRunBytes;
    if (nruns > 0x0FFFFFF0ull) return DC_ERR_ARG;
    LaunchScope ls(DC_K_HISTOGRAM, st);
    const unsigned int grid = (unsigned int)min(nruns, (size_t)sm_count() * 8);
    hist_runs_kernel<<<grid, 256, 0, st>>>(d_in, n, d_hist, d_run_hist, (unsigned int)nruns);
    return cuda_status(cudaGetLastError());
}
}  // namespace dc

// bench/test hook: run a specific counter layout (see the table at the top of this file)
extern "C" int dc_histogram_u8_variant(const uint8_t *d_in, size_t n, uint64_t *d_hist, int variant, void *stream) {
    if (!d_hist || (!d_in && n)) return DC_ERR_ARG;
    return dc::launch_histogram(d_in, n, (unsigned long long *)d_hist, variant, (cudaStream_t)stream);
}
