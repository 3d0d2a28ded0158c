// k1_histogram.cu -- K1: byte histogram with warp-privatised shared-memory counters.
//
// Replaces histogram() (n_ary_huffman.c:461-493): h[b]++ for every input byte, after zeroing h[0..258]
// (:474-476).  Counts are 64-bit (the reference's int overflows at 2^31).
//
// The kernel is bound by the shared-memory atomic rate, not by HBM.  Two counter layouts:
//   lane-private  [256 symbols][32 lanes] u32 per CTA: lane L only touches bank L, an atomic never has a bank conflict
//                 (hist_runs_kernel: the default for 16-byte aligned input; also keeps one histogram per 32 KB run for the
//                 planned encoder).  0.25 ms per GiB on Zipf and on uniform bytes.
//   per-warp      256 u32 per warp, 2.3 wavefronts per atomic (hist_warp_atomic_kernel: round 1's winner among six layouts,
//                 profiles/r1_v1_kernel_timings_1GiB.jsonl; kept for unaligned input).  0.32 / 0.42 ms per GiB.
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

// ------------------------------------------------------------------------------------------ histogram + run histograms
// The encoder wants the bit offset of every 32 KB run before it starts (k3_encode.cu, "planned single pass"); that is
// sum(count[s] * length[s]) over the runs in front -- known as soon as the code lengths are, if the histogram pass keeps
// one small histogram per run.  Same counting as variant A<1>; a CTA walks whole runs, and after each run its 256
// threads fold the eight warp tables into 256 u16 counts (512 bytes per 32 KB of input: 1.6 % extra traffic).
constexpr int kHistRunBytes = 32768;   // == kRunBytes of k3_encode.cu
// Counters: [256 symbols][32 lanes] u32 per CTA -- lane L of every warp only ever touches column L, i.e. bank L, so a
// shared-memory atomic never has a bank conflict (ncu, round 1: 2.3 wavefronts per ATOMS with per-warp 256-entry tables;
// the kernel is bound by exactly those).  The counters are never cleared: after each run, warp w reads its 16 rows (one
// conflict-free load + one REDUX per row) and the run's histogram is the difference to the totals it read the time before.
constexpr int kHistRunThreads = 512;
// (Two tables used alternately -- the snapshot of run r overlapping the counting of run r + 1, one barrier per run -- were
// measured and dropped: 64 KB per CTA allow three CTAs per SM instead of four, 0.29 ms against 0.25.)
template <bool RUNS>
__global__ void __launch_bounds__(kHistRunThreads, 4) hist_runs_kernel(const uint8_t *__restrict__ in, size_t n,
                                                                        unsigned long long *__restrict__ hist,
                                                                        uint16_t *__restrict__ run_hist, unsigned int nruns,
                                                                        unsigned long long *__restrict__ edge) {
    __shared__ uint32_t s_cnt[256 * 32];
    if (edge && blockIdx.x == 0 && threadIdx.x == 0) {
        // what a neighbouring shard needs to complete the byte it shares with this one (shard_nccl.cu): the symbol count,
        // the first and the last eight symbols (byte j of a word = symbol j of the eight)
        const size_t m = n < 8 ? n : 8;
        unsigned long long head = 0, tail = 0;
        for (size_t j = 0; j < m; j++) {
            head |= (unsigned long long)in[j] << (8 * j);
            tail |= (unsigned long long)in[n - m + j] << (8 * j);
        }
        edge[0] = n;
        edge[1] = head;
        edge[2] = tail;
    }
    for (int i = threadIdx.x; i < 256 * 32; i += kHistRunThreads) s_cnt[i] = 0;
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t col = (uint32_t)__cvta_generic_to_shared(s_cnt) + 4u * lane;
    uint32_t seen = 0;              // lane r < 16: the total of row 16 * warp + r at the previous snapshot
    unsigned long long total = 0;   // ... and over all runs of this CTA
#define DC_COUNT_BYTE(w, k)                                                                                          \
    asm volatile("{\n\t.reg .u32 b, a;\n\tprmt.b32 b, %0, 0, 0x444" #k ";\n\tmad.lo.u32 a, b, 128, %1;\n\t"       \
                 "red.shared.add.u32 [a], 1;\n\t}" ::"r"(w), "r"(col) : "memory");
#define DC_COUNT_WORD(w) DC_COUNT_BYTE(w, 0) DC_COUNT_BYTE(w, 1) DC_COUNT_BYTE(w, 2) DC_COUNT_BYTE(w, 3)
    auto snapshot = [&](unsigned int run) {   // the histogram of `run`: the totals now minus the totals at the snapshot before
        uint32_t mine = 0;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, s_cnt[(warp * 16 + r) * 32 + lane]);
            if (lane == r) mine = t;
        }
        if (lane < 16) {
            const uint32_t d = mine - seen;
            if (run_hist) run_hist[(size_t)run * 256 + warp * 16 + lane] = (uint16_t)d;   // <= 32768
            total += d;
            seen = mine;
        }
    };
    for (unsigned int run = blockIdx.x; run < nruns; run += gridDim.x) {
        const size_t base = (size_t)run * kHistRunBytes;
        const uint4 *vin = (const uint4 *)(in + base);
        if (base + kHistRunBytes <= n) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = ldg_stream(vin + tid + u * kHistRunThreads);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                DC_COUNT_WORD(v[u].x) DC_COUNT_WORD(v[u].y) DC_COUNT_WORD(v[u].z) DC_COUNT_WORD(v[u].w)
            }
        } else {  // the ragged last run
            const size_t len = n - base, nvec = len / 16;
            for (size_t i = tid; i < nvec; i += kHistRunThreads) {
                const uint4 x = ldg_stream(vin + i);
                DC_COUNT_WORD(x.x) DC_COUNT_WORD(x.y) DC_COUNT_WORD(x.z) DC_COUNT_WORD(x.w)
            }
            for (size_t i = nvec * 16 + tid; i < len; i += kHistRunThreads) {
                const uint32_t w = in[base + i];
                DC_COUNT_BYTE(w, 0)
            }
        }
        if (!RUNS) continue;   // plain histogram: one snapshot at the end
        __syncthreads();       // every count of this run has landed
        snapshot(run);
        __syncthreads();       // the snapshot is taken: the next run may count
    }
#undef DC_COUNT_WORD
#undef DC_COUNT_BYTE
    if (!RUNS) {
        __syncthreads();
        snapshot(0);
    }
    if (lane < 16 && total) atomicAdd(&hist[warp * 16 + lane], total);
}

int hist_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("DC_HIST_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

// variant 0: lane-private counters when the input is 16-byte aligned, else per-warp tables; variant 1: per-warp tables
int launch_histogram(const uint8_t *d_in, size_t n, unsigned long long *d_hist, int variant, cudaStream_t st) {
    DC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, DC_NSLOTS * sizeof(unsigned long long), st));
    if (n == 0) return DC_OK;
    const int sms = sm_count();
    LaunchScope ls(DC_K_HISTOGRAM, st);
    if (variant == 0 && ((uintptr_t)d_in & 15) == 0 && n < ((size_t)1 << 40)) {
        const size_t nruns = (n + kHistRunBytes - 1) / kHistRunBytes;
        // (a CTA's u32 counters hold its share of the input: at most 2^32 bytes per column needs n / grid < 2^32)
        const unsigned int grid = (unsigned int)min(nruns, (size_t)sms * 4);
        hist_runs_kernel<false><<<grid, kHistRunThreads, 0, st>>>(d_in, n, d_hist, nullptr, (unsigned int)nruns, nullptr);
    } else {
        const size_t nvec = n / 16 + 1;
        const int grid = (int)min((size_t)sms * 8, (nvec + 255) / 256);
        hist_warp_atomic_kernel<1><<<grid, 256, 8 * 256 * 4, st>>>(d_in, n, d_hist);
    }
    return cuda_status(cudaGetLastError());
}

}  // namespace dc

extern "C" int dc_histogram_u8(const uint8_t *d_in, size_t n, uint64_t *d_hist, void *stream) {
    if (!d_hist || (!d_in && n)) return DC_ERR_ARG;
    return dc::launch_histogram(d_in, n, (unsigned long long *)d_hist, dc::hist_variant(), (cudaStream_t)stream);
}

namespace dc {
int launch_histogram_runs(const uint8_t *d_in, size_t n, unsigned long long *d_hist, uint16_t *d_run_hist, cudaStream_t st,
                          unsigned long long *d_edge) {
    DC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, DC_NSLOTS * sizeof(unsigned long long), st));
    if (n == 0) {
        if (d_edge) DC_CUDA_TRY(cudaMemsetAsync(d_edge, 0, 3 * sizeof(unsigned long long), st));
        return DC_OK;
    }
    const size_t nruns = (n + kHistRunBytes - 1) / kHistRunBytes;
    if (nruns > 0x0FFFFFF0ull) return DC_ERR_ARG;
    LaunchScope ls(DC_K_HISTOGRAM, st);
    const unsigned int grid = (unsigned int)min(nruns, (size_t)sm_count() * 4);
    hist_runs_kernel<true><<<grid, kHistRunThreads, 0, st>>>(d_in, n, d_hist, d_run_hist, (unsigned int)nruns, d_edge);
    return cuda_status(cudaGetLastError());
}
}  // namespace dc

// bench/test hook: run a specific counter layout (see the table at the top of this file)
extern "C" int dc_histogram_u8_variant(const uint8_t *d_in, size_t n, uint64_t *d_hist, int variant, void *stream) {
    if (!d_hist || (!d_in && n)) return DC_ERR_ARG;
    return dc::launch_histogram(d_in, n, (unsigned long long *)d_hist, variant, (cudaStream_t)stream);
}
