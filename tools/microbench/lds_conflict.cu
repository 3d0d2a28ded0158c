// Micro-benchmark: cost of a shared-memory gather as a function of its bank-conflict degree (B200).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_conflict lds_conflict.cu && ./lds_conflict
// Every warp issues `iters` LDS.32 whose lane addresses give an exact conflict degree N (N lanes per bank, different
// rows), or pseudo-random words of a 44 KB table (the decoder's pattern).  Independent loads (ILP 4), 32 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) k_lds(int mode, int iters, unsigned long long *out_cycles, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    const int words = 11 * 1024;   // 44 KB
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = (i * 2654435761u) >> 7;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t a[4];
    uint32_t rnd = (threadIdx.x + 1) * 2246822519u + blockIdx.x * 3266489917u;
    for (int j = 0; j < 4; j++) {
        if (mode >= 1 && mode <= 32) {         // exact N-way conflict
            const int N = mode, per = 32 / N;  // lanes [0, per) distinct banks; groups of `per` lanes stack on the same banks
            a[j] = (lane % per) + 32 * (lane / per) + 1024 * j;
        } else {                                // random
            rnd = rnd * 1664525u + 1013904223u;
            a[j] = (rnd >> 8) % words;
        }
    }
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (mode == 100) {   // dependent random chain (like the decoder: next address from the loaded value)
        uint32_t x = a[0];
        for (int i = 0; i < iters * 4; i++) { x = sm[x % words]; }
        acc = x;
    } else if (mode == 101) {   // random, independent, addresses re-randomised from loaded data (mix, ILP 4)
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) { const uint32_t v = sm[a[j]]; a[j] = (v + a[j] * 5u + j) % words; acc += v; }
        }
    } else {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) { acc += sm[a[j]]; a[j] ^= 0; }
            asm volatile("" : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]));
        }
    }
    __syncthreads();
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    unsigned long long *d_c; uint32_t *d_s;
    cudaMalloc(&d_c, 148 * 8); cudaMalloc(&d_s, 4);
    cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 4096;
    int modes[] = {1, 2, 4, 8, 16, 32, 0, 101, 100};
    for (int threads : {1024, 512, 256}) {
        for (int m : modes) {
            k_lds<<<148, threads, 48 * 1024>>>(m, iters, d_c, d_s);
            k_lds<<<148, threads, 48 * 1024>>>(m, iters, d_c, d_s);
            unsigned long long h[148];
            cudaMemcpy(h, d_c, sizeof h, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; i++) avg += (double)h[i]; avg /= 148;
            const double lds = (double)iters * 4 * (threads / 32);
            printf("threads %4d mode %3d: %.2f cycles per warp-LDS (SM-wide)\n", threads, m, avg / lds);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
