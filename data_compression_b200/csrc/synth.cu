// synth.cu -- counter-based synthetic input generator shared by bench.py and the tests.
// byte i = value_base + #{k : thresholds[k] <= (splitmix64(seed + i) >> 32)}; the thresholds are the
// inverse CDF of the wanted distribution (Zipf(1.1) over ranks), computed once on the host in float64 and
// handed to both this kernel and the numpy twin in tests/, so host and device streams are identical.
#include "dc_common.cuh"

namespace dc {

__global__ void __launch_bounds__(256) synth_kernel(uint8_t *__restrict__ out, size_t n, unsigned long long seed,
                                                    const uint32_t *__restrict__ thresholds, int nthresh, int value_base) {
    __shared__ uint32_t s_thr[256];
    for (int i = threadIdx.x; i < 256; i += 256) s_thr[i] = i < nthresh ? thresholds[i] : 0xFFFFFFFFu;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * 256 * 16;
    for (size_t base = ((size_t)blockIdx.x * 256 + threadIdx.x) * 16; base < n; base += stride) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t r = (uint32_t)(splitmix64(seed + base + k) >> 32);
            // rank = number of thresholds <= r (thresholds ascending; entries >= nthresh are +inf unless r is too)
            int lo = 0, hi = nthresh;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_thr[mid] <= r) lo = mid + 1; else hi = mid;
            }
            w[k >> 2] |= (uint32_t)((value_base + lo) & 0xFF) << (8 * (k & 3));
        }
        if (base + 16 <= n && (((uintptr_t)out) & 15) == 0) {
            *(uint4 *)(out + base) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int k = 0; k < 16 && base + k < n; k++) out[base + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
        }
    }
}

}  // namespace dc

extern "C" int dc_synth_fill(uint8_t *d_out, size_t n, uint64_t seed, const uint32_t *d_thresholds, int nthresh,
                             int value_base, void *stream) {
    if ((!d_out && n) || nthresh < 0 || nthresh > 256 || (nthresh && !d_thresholds)) return DC_ERR_ARG;
    if (n == 0) return DC_OK;
    const size_t want = (n / 16 + 256) / 256;
    const size_t cap = (size_t)dc::sm_count() * 8;
    dc::LaunchScope ls(DC_K_SYNTH, (cudaStream_t)stream);
    dc::synth_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(d_out, n, seed, d_thresholds, nthresh,
                                                                                      value_base);
    return dc::cuda_status(cudaGetLastError());
}
