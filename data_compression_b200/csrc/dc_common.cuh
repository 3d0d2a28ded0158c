// dc_common.cuh -- shared device/host helpers for libdc_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "dc_b200.h"

namespace dc {

// every kernel launch in the library goes through LaunchScope: it feeds the launch counter (bench
// "gpu_launches") and, when dc_profile_enable(1) was called, brackets the launch with CUDA events on the
// launching stream so per-kernel durations can be read back (dc_profile_kernel).
extern std::atomic<unsigned long long> g_launches;
void *prof_begin(int kernel_id, cudaStream_t st);            // returns the opening event (nullptr when profiling is off)
void prof_end(int kernel_id, cudaStream_t st, void *open);
struct LaunchScope {
    int id;
    cudaStream_t st;
    void *open;
    LaunchScope(int kernel_id, cudaStream_t stream) : id(kernel_id), st(stream) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        open = prof_begin(id, st);
    }
    ~LaunchScope() { prof_end(id, st, open); }
};

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? DC_OK : DC_ERR_CUDA; }

#define DC_CUDA_TRY(expr)                          \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return DC_ERR_CUDA; \
    } while (0)

// SM count of the current device (cached per device); the grids below are sized in multiples of it
int sm_count();
// raise a kernel's opt-in dynamic shared memory limit on the current device if `bytes` needs it (per device, thread-safe)
cudaError_t ensure_dynamic_smem(const void *func, size_t bytes);

// K2 internals shared with the host-pointer entry points (generic alphabets, raw arrays)
struct TableRaw {
    int32_t *lengths;   // [nsym] code lengths in digits
    uint32_t *values;   // [nsym] canonical code values
    int32_t *assigned;  // [nsym] 1 if convert_lengths_to_encode_table() assigns the slot
    int32_t *status;    // [1]
};
// nhist > 1: the counts are the sum of nhist histograms, hist_stride u64 apart (the gathered histograms of the shard layer)
int launch_table(const unsigned long long *d_hist, const int32_t *d_lengths, int nsym, int n_ary, dc_huff_table *tab,
                 TableRaw raw, cudaStream_t st, int nhist = 1, int hist_stride = 0);

// pipelined form of dc_host_huff_decompress (k4_decode.cu): DC_OK / negative dc_status / +1 = use the one-shot path.
// d_packed != nullptr: radix 3, h_payload is the 5-trits-per-byte payload (total_bits = 2 * trits) and every chunk is
// unpacked into d_bits on the device before it is decoded
int host_decompress_pipelined(const uint8_t *h_payload, uint64_t total_bits, const dc_huff_table *d_table, uint8_t *d_bits,
                              uint8_t *d_out, void *d_workspace, size_t workspace_bytes, uint8_t *h_out, size_t n_out,
                              int32_t *d_status, uint8_t *d_packed);
// K7 unpack without the status reset (k7_trits.cu)
int trit_unpack_launch(const uint8_t *d_payload, unsigned long long ntrits, uint8_t *d_t2, int32_t *d_status, cudaStream_t st);

// K1 with one 256 x u16 histogram per 32 KB run (k1_histogram.cu), for the planned encoder
int launch_histogram_runs(const uint8_t *d_in, size_t n, unsigned long long *d_hist, uint16_t *d_run_hist, cudaStream_t st,
                          unsigned long long *d_edge = nullptr);
// dc_histogram_u8_runs that also leaves {n, first eight symbols, last eight symbols} in d_edge[3] (k3_encode.cu knows the workspace layout)
int histogram_runs_edges(const uint8_t *d_in, size_t n, unsigned long long *d_hist, void *d_encode_workspace, size_t workspace_bytes,
                         unsigned long long *d_edge, cudaStream_t st);

int encode_planned_device_phase(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity,
                                const uint32_t *d_phase, uint64_t *d_total_bits, int32_t *d_status, void *d_workspace,
                                size_t workspace_bytes, cudaStream_t st);

// K8 (k8_mtf.cu): the move-to-front contexts of the adaptive nybble compressor, used by K6
size_t mtf_workspace_bytes(size_t n);
// d_pos[i] = position of d_src[i] in its context's list just before it is touched (8 = absent; d_pos[0] = 8)
int mtf_positions(const uint8_t *d_src, size_t n, uint8_t *d_pos, void *d_ws, cudaStream_t st);
// d_buf[1..*d_len): bytes with bit 7 are 0x80 | position and are replaced by the letter they mean (only if *d_mode == 0)
int mtf_resolve(uint8_t *d_buf, const unsigned long long *d_len, const int32_t *d_mode, cudaStream_t st);

// ---------------------------------------------------------------- what the host knows about a device table (capi.cu)
// K2 leaves the table's first ten words, lut2_used and fsm_states in mapped host memory and the build records an event, so
// the entry points that pick a kernel by the table's shape need neither a device-to-host copy nor a stream synchronise.
constexpr int kTableMetaWords = 12;
struct TableMetaTicket { int32_t *dev; int dev_index, slot; unsigned long long serial; };
// before the K2 launch: a mapped slot for `d_table` (t.dev == nullptr if none could be had: K2 then publishes nothing)
void table_meta_begin(const dc_huff_table *d_table, TableMetaTicket *t);
// after the K2 launch on `st`
void table_meta_end(const dc_huff_table *d_table, const TableMetaTicket &t, cudaStream_t st);
// wait = true: DC_OK, blocks until the build of `d_table` has finished (a table this library did not build, or one whose slot
// has been recycled, is read from the device after synchronising `st`).  wait = false: DC_OK only if the facts are known
// already, 1 otherwise (never blocks).
int table_meta_fetch(const dc_huff_table *d_table, cudaStream_t st, int32_t out[kTableMetaWords], bool wait);
void table_meta_forget(const dc_huff_table *d_table);
// a mapped int32 + an event: a kernel stores one word, the host waits for the event (not for the stream) and reads it
struct HostFlag { volatile int32_t *host; int32_t *dev; cudaEvent_t ev; int dev_index, slot; };
int host_flag_acquire(HostFlag *f);
void host_flag_release(const HostFlag &f);

// radices whose payload is one nibble per digit (k2_table.cu): no window LUTs, decoded by the byte-stepped state machine only
__host__ __device__ inline bool nibble_radix(int n) { return n >= 5 && n < 16; }

// ---------------------------------------------------------------- byte-stepped decoder geometry (k4_fsm.cuh; K2 records the state count)
constexpr int kFsmMaxStates = 255;       // internal nodes of the code tree; F3 adds the DEAD sink with id nstates: 8 bits in all
constexpr int kFsmMaxSyncStates = 256;   // F1 alone has no sink: a binary code of all 256 byte values + the dummy leaf has exactly 256
// Internal nodes per depth from the canonical arrays; returns their number, 0 = not eligible.  `radix` is the arity of the
// code tree (values are base-`radix` numerals), `bpd` the bits a digit takes in the stream (radix 3: 2-bit fields).
__host__ __device__ inline int fsm_geometry(const uint32_t *first, const uint32_t *count, int min_len, int max_len, int bpd, int radix,
                                            uint32_t *ilo, uint32_t *ihi, uint32_t *base) {
    if (!(bpd == 1 || bpd == 2 || bpd == 4) || radix < 2 || radix > (1 << bpd)) return 0;
    if (min_len < 1 || max_len < min_len || max_len >= 16) return 0;
    if (min_len * bpd < 2) return 0;   // a 1-bit code: up to 8 symbols per byte
    if (count[max_len] == 0) return 0;
    const unsigned long long last = (unsigned long long)first[max_len] + count[max_len] - 1;
    // hi(d) = last / radix ^ (max_len - d): shifts for the power-of-two radices (no 64-bit division on the device: K2 runs this
    // on one thread), 32-bit divisions otherwise (radix 3: 3 ^ 15 < 2 ^ 32)
    const bool pow2 = (radix & (radix - 1)) == 0;
    unsigned long long span = 1;       // radix ^ max_len
    for (int d = 0; d < max_len; d++) span *= (unsigned long long)radix;
    if (last >= span) return 0;        // over-subscribed lengths: values do not fit their digits
    if (!pow2 && span > 0xFFFFFFFFull) return 0;
    unsigned total = 0;
    uint32_t div32 = pow2 ? 0u : (uint32_t)span;     // radix ^ (max_len - d)
    for (int d = 0; d < max_len; d++) {
        const unsigned long long lo = d < min_len ? 0ull : (unsigned long long)first[d] + count[d];
        unsigned long long hi;
        if (pow2) {
            hi = last >> (bpd * (max_len - d));
        } else {
            hi = (unsigned long long)((uint32_t)last / div32);
            div32 /= (uint32_t)radix;
        }
        if (hi < lo) return 0;
        if (d >= min_len && d + 1 <= max_len && (unsigned long long)first[d + 1] != lo * (unsigned long long)radix) return 0;   // not the canonical chain
        ilo[d] = (uint32_t)lo;
        ihi[d] = (uint32_t)hi;
        base[d] = total;
        total += (unsigned)(hi - lo + 1);
        if (total > (unsigned)kFsmMaxSyncStates) return 0;
    }
    if (min_len <= max_len && first[min_len] != 0) return 0;
    return (int)total;
}

// ---------------------------------------------------------------- device helpers

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// streaming 128-bit global accesses: data is touched once, keep it out of L1
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint2 ldg_stream8(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream8(uint2 *p, const uint2 &v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// a 16-byte load that bypasses L1 (the boundary slots another CTA has just written)
__device__ __forceinline__ uint4 ld_cg_u128(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

// set a sticky error code once (first error wins)
__device__ __forceinline__ void set_status(int32_t *d_status, int code) {
    if (d_status) atomicCAS(d_status, DC_OK, code);
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

}  // namespace dc
