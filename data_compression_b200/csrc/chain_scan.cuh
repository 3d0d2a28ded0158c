// chain_scan.cuh -- exclusive scan ACROSS the CTAs of one small launch (grid <= SM count, so every CTA is resident and a CTA
// only ever waits for CTAs with smaller block indices, which are dispatched no later than itself).
// Used by the two control kernels that sit between the big ones (encode plan, decode F2): a single CTA scanning 32 K .. 48 K
// values costs 50 .. 70 us of an otherwise idle GPU; 148 CTAs with a few hundred values each cost under 10.
#pragma once

#include "dc_common.cuh"

namespace dc {

struct ChainSlots {
    unsigned long long *vals;   // [grid]
    unsigned int *flags;        // [grid], zeroed before the launch
};
constexpr size_t kChainSlotsBytes = 256 * (sizeof(unsigned long long) + sizeof(unsigned int));   // up to 256 CTAs

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Every thread of the CTA calls this once.  Returns the sum of `mine` over the CTAs in front of this one.
__device__ __forceinline__ unsigned long long chain_exclusive(ChainSlots c, unsigned long long mine) {
    __shared__ unsigned long long s_part[32];
    __shared__ unsigned long long s_res;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    const unsigned int b = blockIdx.x;
    if (tid == 0) {
        c.vals[b] = mine;
        st_release_u32(&c.flags[b], 1u);
    }
    unsigned long long v = 0;
    for (unsigned int t = tid; t < b; t += blockDim.x) {
        while (ld_acquire_u32(&c.flags[t]) == 0u) {}
        v += *(volatile unsigned long long *)&c.vals[t];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) s_part[warp] = v;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < nwarps; w++) t += s_part[w];
        s_res = t;
    }
    __syncthreads();
    return s_res;
}

// exclusive scan of one value per thread over the CTA; *total = the CTA's sum (every thread gets both)
__device__ __forceinline__ unsigned long long block_exclusive(unsigned long long mine, unsigned long long *total) {
    __shared__ unsigned long long s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();   // s_w may still be read from a previous call
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned long long off = 0, tot = 0;
    for (int w = 0; w < nwarps; w++) {
        const unsigned long long x = s_w[w];
        if (w < warp) off += x;
        tot += x;
    }
    *total = tot;
    return off + incl - mine;
}

}  // namespace dc
