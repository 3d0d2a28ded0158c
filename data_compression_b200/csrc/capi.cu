// capi.cu -- library globals and the HOST-pointer entry points of include/dc_b200.h.
//
// The dc_host_* functions are what a C caller of the reference's functions links against (refapi.h maps
// the reference's verbatim names onto them): they stage the caller's host buffers into a grow-only device
// arena, run the same device path as the dc_* entry points, copy the results back and synchronise.
// There is no CPU implementation behind any of them: without a usable device they return DC_ERR_CUDA.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "dc_common.cuh"

namespace dc {

std::atomic<unsigned long long> g_launches{0};

// ---- per-kernel event timing (dc_profile_*).  One mutex guards all of it: launches may come from several host threads.
struct ProfPair { cudaEvent_t a, b; int id; };
static std::mutex g_prof_mu;
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfPair> g_prof_pending;
static std::vector<cudaEvent_t> g_prof_pool;
static double g_prof_ms[DC_K_COUNT];
static unsigned long long g_prof_n[DC_K_COUNT];

static cudaEvent_t prof_event() {   // caller holds g_prof_mu
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
// the opening event travels in the LaunchScope that made it, so concurrent launches of one kernel group do not share state
void *prof_begin(int id, cudaStream_t st) {
    if (!g_prof_on.load(std::memory_order_relaxed) || id < 0 || id >= DC_K_COUNT) return nullptr;
    cudaEvent_t e;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        e = prof_event();
    }
    cudaEventRecord(e, st);
    return (void *)e;
}
void prof_end(int id, cudaStream_t st, void *open) {
    if (!open) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfPair p = {(cudaEvent_t)open, prof_event(), id};
    cudaEventRecord(p.b, st);
    g_prof_pending.push_back(p);
}
static void prof_collect() {   // caller holds g_prof_mu
    for (const ProfPair &p : g_prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_prof_ms[p.id] += (double)ms;
            g_prof_n[p.id]++;
        }
        g_prof_pool.push_back(p.a);
        g_prof_pool.push_back(p.b);
    }
    g_prof_pending.clear();
}

// per-device facts, looked up by the CURRENT device of the calling thread
static std::mutex g_dev_mu;
static int g_sm_count[64];
int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;  // B200
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (g_sm_count[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        g_sm_count[dev] = sms;
    }
    return g_sm_count[dev];
}
// opt-in dynamic shared memory of a kernel: the attribute is per device and only ever grows (it is a limit, not a request)
static std::map<std::pair<int, const void *>, size_t> g_smem_limit;
cudaError_t ensure_dynamic_smem(const void *func, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    size_t &limit = g_smem_limit[std::make_pair(dev, func)];
    if (bytes <= limit) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) limit = bytes;
    return e;
}

// ---- host-visible table facts and flags: rings of mapped pinned words + events, one ring per device
constexpr int kMetaSlots = 256, kFlagSlots = 64, kMaxDevices = 64;
struct MetaSlot { cudaEvent_t ev; const dc_huff_table *tab; unsigned long long serial; };
struct HostRing {
    bool ready = false;
    int32_t *pinned = nullptr;          // kMetaSlots x 16 words, then kFlagSlots x 16 words
    MetaSlot meta[kMetaSlots];
    cudaEvent_t flag_ev[kFlagSlots];
    bool flag_busy[kFlagSlots];
    unsigned next_meta = 0, next_flag = 0;
};
struct MetaRef { int dev, slot; unsigned long long serial; };
static std::mutex g_meta_mu;
static HostRing g_ring[kMaxDevices];
static std::map<const dc_huff_table *, MetaRef> g_meta_of;
static unsigned long long g_meta_serial = 0;

static HostRing *host_ring(int *dev_out) {   // caller holds g_meta_mu
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return nullptr; }
    HostRing &r = g_ring[dev];
    if (!r.ready) {
        if (cudaHostAlloc((void **)&r.pinned, (size_t)(kMetaSlots + kFlagSlots) * 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        memset(r.pinned, 0, (size_t)(kMetaSlots + kFlagSlots) * 64);
        for (int i = 0; i < kMetaSlots; i++) {
            r.meta[i].tab = nullptr;
            r.meta[i].serial = 0;
            if (cudaEventCreateWithFlags(&r.meta[i].ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        }
        for (int i = 0; i < kFlagSlots; i++) {
            r.flag_busy[i] = false;
            if (cudaEventCreateWithFlags(&r.flag_ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        }
        r.ready = true;
    }
    *dev_out = dev;
    return &r;
}

void table_meta_begin(const dc_huff_table *d_table, TableMetaTicket *t) {
    t->dev = nullptr;
    std::lock_guard<std::mutex> lk(g_meta_mu);
    g_meta_of.erase(d_table);   // whatever was known about this address is about to be overwritten
    int dev = 0;
    HostRing *r = host_ring(&dev);
    if (!r) return;
    const int slot = (int)(r->next_meta++ % kMetaSlots);
    MetaSlot &m = r->meta[slot];
    if (m.tab) {   // recycle: the table that held this slot falls back to a device read
        auto it = g_meta_of.find(m.tab);
        if (it != g_meta_of.end() && it->second.dev == dev && it->second.slot == slot) g_meta_of.erase(it);
    }
    m.tab = d_table;
    m.serial = ++g_meta_serial;
    int32_t *host = r->pinned + (size_t)slot * 16, *devp = nullptr;
    if (cudaHostGetDevicePointer((void **)&devp, host, 0) != cudaSuccess) { cudaGetLastError(); m.tab = nullptr; return; }
    t->dev = devp;
    t->dev_index = dev;
    t->slot = slot;
    t->serial = m.serial;
}

void table_meta_end(const dc_huff_table *d_table, const TableMetaTicket &t, cudaStream_t st) {
    if (!t.dev) return;
    std::lock_guard<std::mutex> lk(g_meta_mu);
    MetaSlot &m = g_ring[t.dev_index].meta[t.slot];
    if (m.serial != t.serial) return;   // recycled in between (256 builds raced this one)
    if (cudaEventRecord(m.ev, st) != cudaSuccess) { cudaGetLastError(); m.tab = nullptr; return; }
    g_meta_of[d_table] = MetaRef{t.dev_index, t.slot, t.serial};
}

static int table_meta_read_device(const dc_huff_table *d_table, cudaStream_t st, int32_t out[kTableMetaWords]) {
    DC_CUDA_TRY(cudaMemcpyAsync(out, d_table, 10 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DC_CUDA_TRY(cudaMemcpyAsync(out + 10, (const char *)d_table + offsetof(dc_huff_table, lut2_used), 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DC_CUDA_TRY(cudaStreamSynchronize(st));
    return DC_OK;
}

int table_meta_fetch(const dc_huff_table *d_table, cudaStream_t st, int32_t out[kTableMetaWords], bool wait) {
    MetaRef ref;
    cudaEvent_t ev = nullptr;
    bool known = false;
    {
        std::lock_guard<std::mutex> lk(g_meta_mu);
        auto it = g_meta_of.find(d_table);
        if (it != g_meta_of.end()) {
            ref = it->second;
            ev = g_ring[ref.dev].meta[ref.slot].ev;
            known = g_ring[ref.dev].meta[ref.slot].serial == ref.serial;
        }
    }
    if (known) {
        cudaError_t e = wait ? cudaEventSynchronize(ev) : cudaEventQuery(ev);
        if (e == cudaErrorNotReady) { cudaGetLastError(); return 1; }
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> lk(g_meta_mu);
            const HostRing &r = g_ring[ref.dev];
            if (r.meta[ref.slot].serial == ref.serial) {
                const volatile int32_t *src = r.pinned + (size_t)ref.slot * 16;
                for (int i = 0; i < kTableMetaWords; i++) out[i] = src[i];
                return DC_OK;
            }
        } else {
            cudaGetLastError();
        }
    }
    if (!wait) return 1;
    return table_meta_read_device(d_table, st, out);
}

void table_meta_forget(const dc_huff_table *d_table) {
    std::lock_guard<std::mutex> lk(g_meta_mu);
    g_meta_of.erase(d_table);
}

int host_flag_acquire(HostFlag *f) {
    std::lock_guard<std::mutex> lk(g_meta_mu);
    int dev = 0;
    HostRing *r = host_ring(&dev);
    if (!r) return DC_ERR_CUDA;
    for (int k = 0; k < kFlagSlots; k++) {
        const int slot = (int)(r->next_flag++ % kFlagSlots);
        if (r->flag_busy[slot]) continue;
        int32_t *host = r->pinned + (size_t)(kMetaSlots + slot) * 16, *devp = nullptr;
        if (cudaHostGetDevicePointer((void **)&devp, host, 0) != cudaSuccess) { cudaGetLastError(); return DC_ERR_CUDA; }
        r->flag_busy[slot] = true;
        f->host = host;
        f->dev = devp;
        f->ev = r->flag_ev[slot];
        f->dev_index = dev;
        f->slot = slot;
        return DC_OK;
    }
    return DC_ERR_CUDA;   // 64 calls in flight on one device
}

void host_flag_release(const HostFlag &f) {
    std::lock_guard<std::mutex> lk(g_meta_mu);
    g_ring[f.dev_index].flag_busy[f.slot] = false;
}

// grow-only device scratch for the host-pointer entry points (single-threaded use, like the reference)
struct Arena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes) {
        used = 0;
        if (bytes <= cap) return DC_OK;
        if (base) cudaFree(base);
        base = nullptr;
        cap = 0;
        if (cudaMalloc((void **)&base, bytes) != cudaSuccess) { cudaGetLastError(); return DC_ERR_CUDA; }
        cap = bytes;
        return DC_OK;
    }
    void *take(size_t bytes) {
        void *p = base + used;
        used += (bytes + 255) & ~(size_t)255;
        return p;
    }
    static size_t pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};
static Arena g_arena;

}  // namespace dc

using namespace dc;

extern "C" const char *dc_version(void) { return "dc_b200 0.1 (sm_100a)"; }

extern "C" const char *dc_status_string(int s) {
    switch (s) {
        case DC_OK: return "ok";
        case DC_ERR_ARG: return "bad argument";
        case DC_ERR_CUDA: return "CUDA error or no device (there is no CPU fallback)";
        case DC_ERR_CODE_TOO_LONG: return "code too long (reference limits: length < 16 digits, value fits int)";
        case DC_ERR_CAPACITY: return "output capacity too small";
        case DC_ERR_CORRUPT: return "corrupt bitstream";
        case DC_ERR_SYMBOL: return "symbol without a code / nibble symbol >= 16";
        case DC_ERR_RADIX: return "payload packing needs a radix n <= 16";
        case DC_ERR_NCCL: return "NCCL unavailable (libnccl.so.2 / $DC_NCCL_LIB) or a collective failed";
        default: return "unknown status";
    }
}

extern "C" int dc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return DC_ERR_CUDA; }
    return n;
}

extern "C" uint64_t dc_launch_count(void) { return g_launches.load(); }

extern "C" int dc_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!on) prof_collect();
    g_prof_on = on != 0;
    return DC_OK;
}
extern "C" int dc_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_collect();
    for (int i = 0; i < DC_K_COUNT; i++) { g_prof_ms[i] = 0.0; g_prof_n[i] = 0; }
    return DC_OK;
}
extern "C" int dc_profile_kernel(int id, double *total_ms, uint64_t *launches) {
    if (id < 0 || id >= DC_K_COUNT) return DC_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_collect();
    if (total_ms) *total_ms = g_prof_ms[id];
    if (launches) *launches = g_prof_n[id];
    return DC_OK;
}
extern "C" const char *dc_profile_kernel_name(int id) {
    static const char *names[DC_K_COUNT] = {"histogram", "table", "bits_for_hist", "encode_count", "encode_scan", "encode", "encode_mid", "encode_wide", "decode_sync", "decode_handoff",
                                            "decode_scan", "decode_write", "decode_fast_sync", "decode_fast_scan", "decode_fast_write", "nybble_pack", "nybble_unpack", "nybble_tail", "text_summary", "text_scan", "text_emit", "trit_pack", "trit_unpack", "b64_pack", "b64_unpack", "mtf_walk", "mtf_scan", "mtf_resolve", "text_batch", "synth", "decode_fsm_build", "decode_fsm_sync", "decode_fsm_write", "encode_plan", "encode_fast", "shard_exchange", "shard_plan"};
    return id >= 0 && id < DC_K_COUNT ? names[id] : "?";
}

// ------------------------------------------------------------------------------------------ histogram

extern "C" int dc_host_histogram_u8(const uint8_t *in, size_t n, uint64_t h[DC_NSLOTS]) {
    if (!h || (!in && n)) return DC_ERR_ARG;
    int rc = g_arena.reserve(Arena::pad(n) + Arena::pad(DC_NSLOTS * 8));
    if (rc != DC_OK) return rc;
    uint8_t *d_in = (uint8_t *)g_arena.take(n);
    uint64_t *d_hist = (uint64_t *)g_arena.take(DC_NSLOTS * 8);
    if (n) DC_CUDA_TRY(cudaMemcpyAsync(d_in, in, n, cudaMemcpyHostToDevice, 0));
    rc = dc_histogram_u8(d_in, n, d_hist, nullptr);
    if (rc != DC_OK) return rc;
    DC_CUDA_TRY(cudaMemcpyAsync(h, d_hist, DC_NSLOTS * 8, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    return DC_OK;
}

// histogram(text, max_symbol_value, h) n_ary_huffman.c:461-493: counts until NUL, zeroes h[0..max] first
extern "C" int dc_host_histogram(const char *text, int max_symbol_value, int h[]) {
    if (!text || !h || max_symbol_value < 0) return DC_ERR_ARG;
    uint64_t hist[DC_NSLOTS];
    const int rc = dc_host_histogram_u8((const uint8_t *)text, strlen(text), hist);
    if (rc != DC_OK) return rc;
    for (int i = 0; i <= max_symbol_value; i++) h[i] = i < 256 ? (int)hist[i] : 0;
    return DC_OK;
}

// ------------------------------------------------------------------------------------------ tables

extern "C" int dc_host_huffman_u64(int max_leaf_value, const uint64_t freqs[], int compressed_symbols, int lengths[]) {
    if (!freqs || !lengths || max_leaf_value < 0 || max_leaf_value + 1 > DC_MAX_LEAVES) return DC_ERR_ARG;
    const int nsym = max_leaf_value + 1;
    int rc = g_arena.reserve(Arena::pad((size_t)nsym * 8) + Arena::pad((size_t)nsym * 4));
    if (rc != DC_OK) return rc;
    unsigned long long *d_hist = (unsigned long long *)g_arena.take((size_t)nsym * 8);
    int32_t *d_len = (int32_t *)g_arena.take((size_t)nsym * 4);
    DC_CUDA_TRY(cudaMemcpyAsync(d_hist, freqs, (size_t)nsym * 8, cudaMemcpyHostToDevice, 0));
    TableRaw raw = {d_len, nullptr, nullptr, nullptr};
    rc = launch_table(d_hist, nullptr, nsym, compressed_symbols, nullptr, raw, 0);
    if (rc != DC_OK) return rc;
    DC_CUDA_TRY(cudaMemcpyAsync(lengths, d_len, (size_t)nsym * 4, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    return DC_OK;
}

// huffman(max_leaf_value, symbol_frequencies, compressed_symbols, lengths) n_ary_huffman.c:1161-1208
extern "C" int dc_host_huffman(int max_leaf_value, const int freqs[], int compressed_symbols, int lengths[]) {
    if (!freqs || max_leaf_value < 0 || max_leaf_value + 1 > DC_MAX_LEAVES) return DC_ERR_ARG;
    uint64_t f64[DC_MAX_LEAVES];
    for (int i = 0; i <= max_leaf_value; i++) {
        if (freqs[i] < 0) return DC_ERR_ARG;  // assert( 0 <= symbol_frequencies[i] ) :793
        f64[i] = (uint64_t)freqs[i];
    }
    return dc_host_huffman_u64(max_leaf_value, f64, compressed_symbols, lengths);
}

// convert_lengths_to_encode_table(...) n_ary_huffman.c:1382-1612
extern "C" int dc_host_convert_lengths_to_encode_table(int max_symbol_value, const int lens[], int compressed_symbols,
                                                       int elen[], unsigned int evalue[]) {
    if (!lens || !elen || !evalue || max_symbol_value < 1 || max_symbol_value + 1 > DC_MAX_LEAVES) return DC_ERR_ARG;
    const int nsym = max_symbol_value + 1;
    int rc = g_arena.reserve(4 * Arena::pad((size_t)nsym * 4) + 256);
    if (rc != DC_OK) return rc;
    int32_t *d_len = (int32_t *)g_arena.take((size_t)nsym * 4);
    uint32_t *d_val = (uint32_t *)g_arena.take((size_t)nsym * 4);
    int32_t *d_asg = (int32_t *)g_arena.take((size_t)nsym * 4);
    int32_t *d_st = (int32_t *)g_arena.take(4);
    DC_CUDA_TRY(cudaMemcpyAsync(d_len, lens, (size_t)nsym * 4, cudaMemcpyHostToDevice, 0));
    TableRaw raw = {nullptr, d_val, d_asg, d_st};
    rc = launch_table(nullptr, d_len, nsym, compressed_symbols, nullptr, raw, 0);
    if (rc != DC_OK) return rc;
    uint32_t val[DC_MAX_LEAVES];
    int32_t asg[DC_MAX_LEAVES], st = 0;
    DC_CUDA_TRY(cudaMemcpyAsync(val, d_val, (size_t)nsym * 4, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaMemcpyAsync(asg, d_asg, (size_t)nsym * 4, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaMemcpyAsync(&st, d_st, 4, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    // the reference clears slots i < max_symbol_value only (:1421) and assigns i <= max_symbol_value (:1547)
    for (int i = 0; i < max_symbol_value; i++) {
        elen[i] = asg[i] ? lens[i] : 0;
        evalue[i] = asg[i] ? val[i] : 0u;
    }
    if (asg[max_symbol_value]) {
        elen[max_symbol_value] = lens[max_symbol_value];
        evalue[max_symbol_value] = val[max_symbol_value];
    }
    return st;
}

// ------------------------------------------------------------------------------------------ payload

static int encode_from_table(const uint8_t *h_in, size_t n, dc_huff_table *d_tab, uint8_t *h_out, size_t out_capacity,
                             uint64_t *total_bits, size_t *bytes_written, uint8_t *d_in, uint8_t *d_out, size_t d_cap,
                             void *d_ws, size_t ws_bytes, uint64_t *d_bits, int32_t *d_status, bool input_resident, bool trits,
                             uint8_t *d_packed, bool planned = false) {
    if (!input_resident && n) DC_CUDA_TRY(cudaMemcpyAsync(d_in, h_in, n, cudaMemcpyHostToDevice, 0));
    int rc = planned ? dc_huff_encode_planned(d_in, n, d_tab, d_out, d_cap, 0, d_bits, d_status, d_ws, ws_bytes, nullptr)
                     : dc_huff_encode(d_in, n, d_tab, d_out, d_cap, 0, d_bits, d_status, d_ws, ws_bytes, nullptr);
    if (rc != DC_OK) return rc;
    uint64_t bits = 0;
    int32_t st = 0;
    DC_CUDA_TRY(cudaMemcpyAsync(&bits, d_bits, 8, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    if (st != DC_OK) return st;
    // radix 3: the kernels wrote one 2-bit field per trit; the payload holds 5 trits per byte (K7)
    const uint8_t *d_payload = d_out;
    size_t nbytes = (size_t)((bits + 7) / 8);
    if (trits) {
        const uint64_t ntrits = bits / 2;
        rc = dc_trit_pack(d_out, ntrits, d_packed, d_status, nullptr);
        if (rc != DC_OK) return rc;
        d_payload = d_packed;
        nbytes = (size_t)((ntrits + 4) / 5);
    }
    if (nbytes > out_capacity) return DC_ERR_CAPACITY;
    if (nbytes) DC_CUDA_TRY(cudaMemcpyAsync(h_out, d_payload, nbytes, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    if (total_bits) *total_bits = bits;
    *bytes_written = nbytes;
    return DC_OK;
}

static size_t device_out_capacity(size_t n, size_t host_capacity, bool trits) {
    const size_t worst = n * 4 + 64;  // 32 bits per symbol
    // (radix 3: the intermediate 2-bit-per-trit stream is 1.25 x the payload the host buffer is sized for)
    const size_t want = trits ? host_capacity + host_capacity / 4 + 64 : host_capacity;
    return ((want < worst ? want : worst) + 15) & ~(size_t)15;
}

// represent_items_with_codes(...) n_ary_huffman.c:1621-1678
extern "C" int dc_host_represent_items_with_codes(int max_symbol_value, const int lens[], int compressed_symbols, int bufsize,
                                                  int original_length, const char text[], int start, char out[],
                                                  uint64_t *total_bits) {
    if (max_symbol_value != DC_MAX_SYMBOL_VALUE || !lens || !text || !out || original_length < 0 || start < 0 ||
        start > bufsize || original_length > bufsize)
        return DC_ERR_ARG;
    const size_t n = (size_t)original_length, host_cap = (size_t)(bufsize + 1 - start);
    const bool trits = compressed_symbols == 3;
    const size_t d_cap = device_out_capacity(n, host_cap, trits), ws_bytes = dc_huff_encode_workspace_bytes(n);
    int rc = g_arena.reserve(Arena::pad(n) + 2 * Arena::pad(d_cap) + Arena::pad(ws_bytes) + Arena::pad(sizeof(dc_huff_table)) +
                             Arena::pad(DC_NSLOTS * 4) + 512);
    if (rc != DC_OK) return rc;
    uint8_t *d_in = (uint8_t *)g_arena.take(n);
    uint8_t *d_out = (uint8_t *)g_arena.take(d_cap);
    uint8_t *d_packed = (uint8_t *)g_arena.take(d_cap);
    void *d_ws = g_arena.take(ws_bytes);
    dc_huff_table *d_tab = (dc_huff_table *)g_arena.take(sizeof(dc_huff_table));
    int32_t *d_len = (int32_t *)g_arena.take(DC_NSLOTS * 4);
    uint64_t *d_bits = (uint64_t *)g_arena.take(8);
    int32_t *d_status = (int32_t *)g_arena.take(4);
    DC_CUDA_TRY(cudaMemcpyAsync(d_len, lens, DC_NSLOTS * 4, cudaMemcpyHostToDevice, 0));
    rc = dc_huff_table_from_lengths(d_len, compressed_symbols, d_tab, nullptr);
    if (rc != DC_OK) return rc;
    size_t written = 0;
    rc = encode_from_table((const uint8_t *)text, n, d_tab, (uint8_t *)out + start, host_cap, total_bits, &written, d_in, d_out,
                           d_cap, d_ws, ws_bytes, d_bits, d_status, false, trits, d_packed);
    return rc != DC_OK ? rc : (int)written;
}

extern "C" long long dc_host_huff_compress(const uint8_t *in, size_t n, int compressed_symbols, uint8_t *out, size_t out_capacity,
                                           int lengths_out[DC_NSLOTS], uint64_t *total_bits) {
    if ((!in && n) || (!out && out_capacity) || !lengths_out) return DC_ERR_ARG;
    const bool trits = compressed_symbols == 3;
    const size_t d_cap = device_out_capacity(n, out_capacity, trits), ws_bytes = dc_huff_encode_workspace_bytes(n);
    int rc = g_arena.reserve(Arena::pad(n) + (trits ? 2 : 1) * Arena::pad(d_cap) + Arena::pad(ws_bytes) +
                             Arena::pad(sizeof(dc_huff_table)) + Arena::pad(DC_NSLOTS * 8) + 512);
    if (rc != DC_OK) return rc;
    uint8_t *d_in = (uint8_t *)g_arena.take(n);
    uint8_t *d_out = (uint8_t *)g_arena.take(d_cap);
    uint8_t *d_packed = trits ? (uint8_t *)g_arena.take(d_cap) : nullptr;
    void *d_ws = g_arena.take(ws_bytes);
    dc_huff_table *d_tab = (dc_huff_table *)g_arena.take(sizeof(dc_huff_table));
    uint64_t *d_hist = (uint64_t *)g_arena.take(DC_NSLOTS * 8);
    uint64_t *d_bits = (uint64_t *)g_arena.take(8);
    int32_t *d_status = (int32_t *)g_arena.take(4);
    if (n) DC_CUDA_TRY(cudaMemcpyAsync(d_in, in, n, cudaMemcpyHostToDevice, 0));
    // the histogram pass leaves one small histogram per 32 KB run in the workspace: the encoder then knows every run's bit
    // offset without reading the input a second time
    rc = dc_histogram_u8_runs(d_in, n, d_hist, d_ws, ws_bytes, nullptr);
    if (rc != DC_OK) return rc;
    rc = dc_huff_build(d_hist, compressed_symbols, d_tab, nullptr);
    if (rc != DC_OK) return rc;
    size_t written = 0;
    rc = encode_from_table(in, n, d_tab, out, out_capacity, total_bits, &written, d_in, d_out, d_cap, d_ws, ws_bytes, d_bits,
                           d_status, true, trits, d_packed, true);
    if (rc != DC_OK) return rc;
    DC_CUDA_TRY(cudaMemcpy(lengths_out, (const char *)d_tab + offsetof(dc_huff_table, lengths), DC_NSLOTS * 4,
                           cudaMemcpyDeviceToHost));
    return (long long)written;
}

extern "C" int dc_host_huff_decompress(const uint8_t *payload, uint64_t total_bits, const int lens[DC_NSLOTS],
                                       int compressed_symbols, uint8_t *out, size_t n_out) {
    if ((!payload && total_bits) || !lens || (!out && n_out)) return DC_ERR_ARG;
    const bool trits = compressed_symbols == 3;   // payload = 5 trits per byte; total_bits = 2 * trits (K7)
    const uint64_t ntrits = total_bits / 2;
    const size_t nbytes = (size_t)((total_bits + 7) / 8), ws_bytes = dc_huff_decode_workspace_bytes(0, total_bits);
    const size_t pbytes = trits ? (size_t)((ntrits + 4) / 5) : 0;
    int rc = g_arena.reserve(Arena::pad(nbytes + 64) + Arena::pad(pbytes + 64) + Arena::pad(n_out + 16) + Arena::pad(ws_bytes) +
                             Arena::pad(sizeof(dc_huff_table)) + Arena::pad(DC_NSLOTS * 4) + 512);
    if (rc != DC_OK) return rc;
    uint8_t *d_bits = (uint8_t *)g_arena.take(nbytes + 64);
    uint8_t *d_packed = (uint8_t *)g_arena.take(pbytes + 64);
    uint8_t *d_out = (uint8_t *)g_arena.take(n_out + 16);
    void *d_ws = g_arena.take(ws_bytes);
    dc_huff_table *d_tab = (dc_huff_table *)g_arena.take(sizeof(dc_huff_table));
    int32_t *d_len = (int32_t *)g_arena.take(DC_NSLOTS * 4);
    int32_t *d_status = (int32_t *)g_arena.take(4);
    DC_CUDA_TRY(cudaMemcpyAsync(d_len, lens, DC_NSLOTS * 4, cudaMemcpyHostToDevice, 0));
    rc = dc_huff_table_from_lengths(d_len, compressed_symbols, d_tab, nullptr);
    if (rc != DC_OK) return rc;
    if (trits) {
        if (total_bits & 1) return DC_ERR_ARG;
        // large streams: chunked like the other radices, every chunk unpacked on the device before it is decoded
        rc = host_decompress_pipelined(payload, total_bits, d_tab, d_bits, d_out, d_ws, ws_bytes, out, n_out, d_status, d_packed);
        if (rc <= 0) return rc;
        if (rc != 1) return DC_ERR_CUDA;
        if (pbytes) DC_CUDA_TRY(cudaMemcpyAsync(d_packed, payload, pbytes, cudaMemcpyHostToDevice, 0));
        rc = dc_trit_unpack(d_packed, ntrits, d_bits, d_status, nullptr);
        if (rc != DC_OK) return rc;
        int32_t st0 = 0;
        DC_CUDA_TRY(cudaMemcpy(&st0, d_status, 4, cudaMemcpyDeviceToHost));
        if (st0 != DC_OK) return st0;
    } else {
        // large streams: chunked, with the upload, the decode and the download overlapping
        rc = host_decompress_pipelined(payload, total_bits, d_tab, d_bits, d_out, d_ws, ws_bytes, out, n_out, d_status, nullptr);
        if (rc <= 0) return rc;
        if (rc != 1 && nbytes) return DC_ERR_CUDA;
        // one-shot path.  (After a pipelined attempt that fell back the payload is on the device already; copying it
        // again keeps this path independent of that.)
        if (nbytes) DC_CUDA_TRY(cudaMemcpyAsync(d_bits, payload, nbytes, cudaMemcpyHostToDevice, 0));
    }
    rc = dc_huff_decode(d_bits, 0, total_bits, d_tab, d_out, n_out, d_status, d_ws, ws_bytes, nullptr);
    if (rc != DC_OK) return rc;
    int32_t st = 0;
    DC_CUDA_TRY(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, 0));
    if (n_out) DC_CUDA_TRY(cudaMemcpyAsync(out, d_out, n_out, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    return st;
}

// ------------------------------------------------------------------------------------------ nybble

extern "C" int dc_host_nybble_pack(const uint8_t *sym, size_t n_sym, uint8_t *packed) {
    if ((!sym || !packed) && n_sym) return DC_ERR_ARG;
    const size_t nb = (n_sym + 1) / 2;
    int rc = g_arena.reserve(Arena::pad(n_sym) + Arena::pad(nb) + 256);
    if (rc != DC_OK) return rc;
    uint8_t *d_sym = (uint8_t *)g_arena.take(n_sym), *d_packed = (uint8_t *)g_arena.take(nb);
    int32_t *d_status = (int32_t *)g_arena.take(4);
    if (n_sym) DC_CUDA_TRY(cudaMemcpyAsync(d_sym, sym, n_sym, cudaMemcpyHostToDevice, 0));
    rc = dc_nybble_pack(d_sym, n_sym, d_packed, d_status, nullptr);
    if (rc != DC_OK) return rc;
    int32_t st = 0;
    DC_CUDA_TRY(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, 0));
    if (nb) DC_CUDA_TRY(cudaMemcpyAsync(packed, d_packed, nb, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    return st;
}

extern "C" int dc_host_nybble_unpack(const uint8_t *packed, size_t n_sym, uint8_t *sym) {
    if ((!sym || !packed) && n_sym) return DC_ERR_ARG;
    const size_t nb = (n_sym + 1) / 2;
    int rc = g_arena.reserve(Arena::pad(n_sym) + Arena::pad(nb));
    if (rc != DC_OK) return rc;
    uint8_t *d_sym = (uint8_t *)g_arena.take(n_sym), *d_packed = (uint8_t *)g_arena.take(nb);
    if (nb) DC_CUDA_TRY(cudaMemcpyAsync(d_packed, packed, nb, cudaMemcpyHostToDevice, 0));
    rc = dc_nybble_unpack(d_packed, n_sym, d_sym, nullptr);
    if (rc != DC_OK) return rc;
    if (n_sym) DC_CUDA_TRY(cudaMemcpyAsync(sym, d_sym, n_sym, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    return DC_OK;
}

// ------------------------------------------------------------------------------------------ nybble compressor (static table / adaptive contexts)

static long long host_text(const char *source, char *dest, int modify, bool compress) {
    if (!source || !dest) return DC_ERR_ARG;
    const size_t n = strlen(source);
    const size_t cap = compress ? n + 2 : 2 * n + 2;
    const size_t ws_bytes = modify ? dc_nybble_adaptive_workspace_bytes(n) : dc_nybble_text_workspace_bytes(n);
    int rc = g_arena.reserve(Arena::pad(n + 16) + Arena::pad(cap + 16) + Arena::pad(ws_bytes) + 512);
    if (rc != DC_OK) return rc;
    uint8_t *d_src = (uint8_t *)g_arena.take(n + 16), *d_dst = (uint8_t *)g_arena.take(cap + 16);
    void *d_ws = g_arena.take(ws_bytes);
    uint64_t *d_len = (uint64_t *)g_arena.take(8);
    int32_t *d_status = (int32_t *)g_arena.take(4);
    if (n) DC_CUDA_TRY(cudaMemcpyAsync(d_src, source, n, cudaMemcpyHostToDevice, 0));
    if (modify)
        rc = compress ? dc_nybble_adaptive_compress(d_src, n, d_dst, cap, d_len, d_status, d_ws, ws_bytes, nullptr)
                      : dc_nybble_adaptive_decompress(d_src, n, d_dst, cap, d_len, d_status, d_ws, ws_bytes, nullptr);
    else
        rc = compress ? dc_nybble_text_compress(d_src, n, d_dst, cap, d_len, d_status, d_ws, ws_bytes, nullptr)
                      : dc_nybble_text_decompress(d_src, n, d_dst, cap, d_len, d_status, d_ws, ws_bytes, nullptr);
    if (rc != DC_OK) return rc;
    uint64_t len = 0;
    int32_t st = 0;
    DC_CUDA_TRY(cudaMemcpyAsync(&len, d_len, 8, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, 0));
    DC_CUDA_TRY(cudaStreamSynchronize(0));
    if (st != DC_OK) return st;
    if (len) DC_CUDA_TRY(cudaMemcpy(dest, d_dst, (size_t)len, cudaMemcpyDeviceToHost));
    dest[len] = '\0';
    return (long long)len;
}

extern "C" long long dc_host_compress_bytestring(const char *source, char *dest, int modify) {
    return host_text(source, dest, modify, true);
}
extern "C" long long dc_host_decompress_bytestring(const char *source, char *dest, int modify) {
    return host_text(source, dest, modify, false);
}
