// shard_nccl.cu -- the multi-GPU form of the hot path behind the C-ABI (SURVEY 8b last cell, 8e): one process per GPU,
// one NCCL communicator per process supplied by (or created for) the caller.
//
// Nothing here is a compute step followed by a bulk collective: the ranks exchange a few hundred bytes.
//   encode   all-gather of the LOCAL histograms (259 x u64 per rank).  Every rank sums them (one global code table,
//            built redundantly: K2 is deterministic) and computes EVERY rank's bit total = sum hist_r[s] * length[s],
//            so the exclusive scan of bit totals needs no second collective.  Rank r encodes at bit phase O_r mod 8
//            (read from device memory: nothing blocks).  The all-gather also carries every rank's symbol count and its
//            first and last eight symbols: with the global table a rank re-creates the few bits its neighbours put into
//            the byte it shares with them, so the shared bytes are completed without a second collective.  The
//            concatenation of the shard buffers at byte offsets O_r / 8 is the single-stream payload bit for bit;
//            dc_shard_huff_gather() places them in one buffer.
//   decode   of ONE stream cut blindly into byte ranges (BASELINE config 5): neighbours exchange 1 KB halos
//            (ncclSend/ncclRecv), every rank finds its first code by synchronising over its left neighbour's tail, the
//            24-byte summaries are all-gathered, assumed starts are checked against real exits, symbol counts become
//            output offsets.
//   nybble   pack / unpack shards need no exchange at all: shard starts are kept on even symbol indices.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, or $DC_NCCL_LIB): libdc_b200.so has no link-time dependency on it,
// and a process that already loaded NCCL (torch) shares that copy.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "dc_common.cuh"

namespace dc {

// ---- the few NCCL entry points used, by their public C signatures (nccl.h)
typedef void *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclUint8 = 1, kNcclUint64 = 5 };
struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    bool ok;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static const NcclApi *nccl() {
    std::call_once(g_nccl_once, [] {
        const char *names[] = {getenv("DC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        void *h = nullptr;
        for (const char *nm : names)
            if (nm && !h) h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
        g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
        g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
        g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
        g_nccl.Send = (int (*)(const void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclSend");
        g_nccl.Recv = (int (*)(void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclRecv");
        g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
        g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
        g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather && g_nccl.Send && g_nccl.Recv &&
                    g_nccl.GroupStart && g_nccl.GroupEnd;
    });
    return g_nccl.ok ? &g_nccl : nullptr;
}

#define DC_NCCL_TRY(expr)                 \
    do {                                  \
        if ((expr) != 0) return DC_ERR_NCCL; \
    } while (0)

}  // namespace dc

namespace dc {
// ---- peer-memory exchange: every rank owns one small buffer that all the others map (CUDA IPC); a rank STORES its 2 KB
// contribution straight into its peers' buffers over NVLink and raises a flag there, and spins on the flags in its own.
// One single-CTA kernel, a few microseconds, against 25-30 us for the same all-gather through NCCL (whose cost at this size
// is all latency).  Two slots used alternately: a rank can be at most one call ahead of its slowest peer (to finish call e
// it needs every peer's flag of call e), and it overwrites slot (e & 1) of its peers next in call e + 2.
constexpr int kMaxPeers = 16;
struct PeerPtrs {
    unsigned long long *buf[kMaxPeers];   // buf[r] = rank r's exchange buffer as mapped in this process (buf[rank] = my own)
};
struct PeerXchg {
    int state = 0;                        // 0 = not tried yet, 1 = in use, -1 = unavailable (NCCL does the exchange)
    PeerPtrs ptrs;
    unsigned long long epoch = 0;
    size_t flags_off = 0;                 // in u64: [2][world][kShardSlots] data, then [2][world] flags
};
}  // namespace dc

struct dc_shard_comm {
    dc::ncclComm_t comm;
    int rank, world;
    bool owned;
    dc::PeerXchg peer;
};

namespace dc {

// ---- device side of the encode plan.  all_hist: [world][259] u64 (what the all-gather left).
constexpr int kShardSlots = DC_NSLOTS + 3;   // what a rank contributes to the all-gather: histogram, symbol count, first 8 and last 8 symbols
struct ShardPlan {                        // one per call, in the workspace; read back by dc_shard_huff_encode_info
    unsigned long long bit_offset, bits, total_bits, reserved;
    uint32_t phase, pad[3];
};
// bits of every rank under the global table, their exclusive scan, this rank's phase
__global__ void __launch_bounds__(256) shard_plan_kernel(const unsigned long long *__restrict__ all_hist, int world, int rank,
                                                        const dc_huff_table *__restrict__ tab, unsigned long long *__restrict__ rank_bits,
                                                        unsigned long long *__restrict__ rank_off, ShardPlan *__restrict__ plan) {
    // one warp per rank (8 at a time): a lane takes 8 of the 256 symbols, the warp adds up; then one thread scans the totals
    __shared__ unsigned long long s_bits[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long len[8];
#pragma unroll
    for (int k = 0; k < 8; k++) len[k] = (unsigned long long)(tab->lengths[lane + 32 * k] * tab->bits_per_digit);
    for (int r = warp; r < world; r += 8) {
        unsigned long long sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) sum += all_hist[(size_t)r * kShardSlots + lane + 32 * k] * len[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if (lane == 0) {
            if (r < 64) s_bits[r] = sum;
            rank_bits[r] = sum;
        }
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int r = 0; r < world; r++) {
            const unsigned long long bits_r = r < 64 ? s_bits[r] : rank_bits[r];
            rank_off[r] = run;
            run += bits_r;
        }
        rank_off[world] = run;
        plan->bit_offset = rank_off[rank];
        plan->bits = world <= 64 ? s_bits[rank] : rank_bits[rank];
        plan->total_bits = run;
        plan->phase = (uint32_t)(rank_off[rank] & 7ull);
    }
}
// The bytes of the stream that several shards touch are the OR of what each of them wrote there (every shard writes zeros
// outside its own bits).  A rank completes its first and its last byte itself: the bits in front of its first bit are the
// tail of the codes of the symbols before its shard, the bits behind its last bit the head of the codes of the symbols
// after it -- at most 7 bits each way, i.e. at most 7 symbols, and the all-gather brought 8 from every rank.
__global__ void shard_fix_edges_kernel(uint8_t *__restrict__ out, const unsigned long long *__restrict__ all, int world, int rank,
                                       const dc_huff_table *__restrict__ tab, const unsigned long long *__restrict__ rank_bits,
                                       const unsigned long long *__restrict__ rank_off, const ShardPlan *__restrict__ plan,
                                       int32_t *__restrict__ d_status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (plan && plan->reserved && d_status) atomicCAS(d_status, 0, DC_ERR_NCCL);   // the peer-memory exchange gave up waiting (plan == nullptr: NCCL did the exchange)
    const unsigned long long bits = rank_bits[rank];
    if (bits == 0 || tab->status != DC_OK) return;
    const unsigned long long off = rank_off[rank], e = off + bits, lo = off >> 3, hi1 = (e - 1) >> 3;
    const uint32_t k = (uint32_t)(off & 7ull), m = (uint32_t)((8ull - (e & 7ull)) & 7ull);
    uint32_t first = 0, last = 0;
    if (k) {   // stream bits [off - k, off): walk the symbols in front of this shard backwards
        unsigned long long acc = 0;
        uint32_t have = 0;
        for (int q = rank - 1; q >= 0 && have < k; q--) {
            const unsigned long long nq = all[(size_t)q * kShardSlots + DC_NSLOTS], tail = all[(size_t)q * kShardSlots + DC_NSLOTS + 2];
            const int cnt = nq < 8 ? (int)nq : 8;
            for (int j = cnt - 1; j >= 0 && have < k; j--) {
                const unsigned long long ent = tab->enc64[(tail >> (8 * j)) & 0xFFull];
                acc |= (ent & 0xFFFFFFFFull) << have;   // bit 0 of acc = the bit just in front of `off`
                have += (uint32_t)(ent >> 32);
            }
        }
        first = (uint32_t)(acc & ((1ull << k) - 1ull)) << (8u - k);
    }
    if (m) {   // stream bits [e, e + m): the symbols behind this shard, forwards; behind the end of the stream: zero padding
        unsigned long long acc = 0;
        uint32_t have = 0;
        for (int q = rank + 1; q < world && have < m; q++) {
            const unsigned long long nq = all[(size_t)q * kShardSlots + DC_NSLOTS], head = all[(size_t)q * kShardSlots + DC_NSLOTS + 1];
            const int cnt = nq < 8 ? (int)nq : 8;
            for (int j = 0; j < cnt && have < m; j++) {
                const unsigned long long ent = tab->enc64[(head >> (8 * j)) & 0xFFull];
                const uint32_t len = (uint32_t)(ent >> 32);
                acc = (acc << len) | (ent & 0xFFFFFFFFull);
                have += len;
            }
        }
        last = (uint32_t)(have >= m ? acc >> (have - m) : acc << (m - have)) & ((1u << m) - 1u);
    }
    if (first) out[0] |= (uint8_t)first;
    if (last) out[hi1 - lo] |= (uint8_t)last;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one CTA: push my kShardSlots words into slot (epoch & 1), row `rank`, of every rank's buffer, flag them, wait for everybody's
__global__ void __launch_bounds__(320) shard_exchange_kernel(PeerPtrs p, const unsigned long long *__restrict__ local, int rank, int world,
                                                             unsigned long long epoch, size_t flags_off, unsigned long long *__restrict__ timed_out) {
    const int t = threadIdx.x;
    const size_t slot = (size_t)(epoch & 1ull);
    if (t == 0) *timed_out = 0ull;
    __syncthreads();
    if (t < kShardSlots) {
        const unsigned long long v = local[t];
        for (int r = 0; r < world; r++) p.buf[r][(slot * world + rank) * kShardSlots + t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < world) st_release_sys_u64(p.buf[t] + flags_off + slot * world + rank, epoch);
    if (t < world) {
        const unsigned long long *flag = p.buf[rank] + flags_off + slot * world + t;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys_u64(flag) != epoch) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 60000000000ull) {   // 60 s: a peer never made this call; report instead of hanging the device
                *timed_out = 1ull;   // (shard_fix_edges_kernel turns it into DC_ERR_NCCL behind the encoder's own status)
                break;
            }
        }
    }
    __syncthreads();
}

struct ShardEncLayout {
    size_t local_hist, all_hist, rank_bits, rank_off, plan, enc_ws, total;
};
static ShardEncLayout shard_enc_layout(size_t n_local, int world) {
    ShardEncLayout L;
    size_t p = 0;
    auto take = [&](size_t bytes) { size_t o = p; p += (bytes + 255) & ~(size_t)255; return o; };
    L.local_hist = take(kShardSlots * 8);
    L.all_hist = take((size_t)world * kShardSlots * 8);
    L.rank_bits = take((size_t)world * 8);
    L.rank_off = take((size_t)(world + 1) * 8);
    L.plan = take(sizeof(ShardPlan));
    L.enc_ws = take(dc_huff_encode_workspace_bytes(n_local));
    L.total = p;
    return L;
}

}  // namespace dc

using namespace dc;

// ------------------------------------------------------------------------------------------ communicator

extern "C" int dc_shard_unique_id(void *id128) {
    const NcclApi *N = nccl();
    if (!N || !id128) return N ? DC_ERR_ARG : DC_ERR_NCCL;
    ncclUniqueId id;
    DC_NCCL_TRY(N->GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return DC_OK;
}

extern "C" int dc_shard_comm_create(const void *id128, int rank, int world, dc_shard_comm **out) {
    const NcclApi *N = nccl();
    if (!N) return DC_ERR_NCCL;
    if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return DC_ERR_ARG;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    DC_NCCL_TRY(N->CommInitRank(&c, world, id, rank));
    *out = new dc_shard_comm{c, rank, world, true, PeerXchg()};
    return DC_OK;
}

extern "C" int dc_shard_comm_from_nccl(void *nccl_comm, int rank, int world, dc_shard_comm **out) {
    if (!nccl()) return DC_ERR_NCCL;
    if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world) return DC_ERR_ARG;
    *out = new dc_shard_comm{(ncclComm_t)nccl_comm, rank, world, false, PeerXchg()};
    return DC_OK;
}

// Maps every rank's exchange buffer into this process.  Collective (two small NCCL all-gathers, blocking); every rank
// reaches the same verdict.  Returns true if the peer path is in use.
static bool peer_setup(dc_shard_comm *c, cudaStream_t st) {
    PeerXchg &x = c->peer;
    if (x.state != 0) return x.state > 0;
    x.state = -1;
    const NcclApi *N = nccl();
    const char *env = getenv("DC_SHARD_PEER");
    // (the decision below must be the same on every rank: it only depends on the world size, the environment -- assumed
    // equal across ranks, like NCCL's own -- and the all-gathered outcome)
    if (!N || c->world < 2 || c->world > kMaxPeers || (env && atoi(env) == 0)) return false;
    const size_t data_words = (size_t)2 * c->world * kShardSlots, words = data_words + (size_t)2 * c->world;
    unsigned long long *mine = nullptr;
    unsigned char *d_h = nullptr;
    int okay = 1;
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof h);
    const bool dbg = getenv("DC_SHARD_DEBUG") != nullptr;
    cudaError_t ce;
    if ((ce = cudaMalloc((void **)&mine, words * 8)) != cudaSuccess) okay = 0;
    if (okay && (ce = cudaMemsetAsync(mine, 0, words * 8, st)) != cudaSuccess) okay = 0;
    if (okay && (ce = cudaIpcGetMemHandle(&h, mine)) != cudaSuccess) okay = 0;
    if (dbg && !okay) fprintf(stderr, "[dc_shard] rank %d: exchange buffer / IPC handle: %s\n", c->rank, cudaGetErrorString(ce));
    cudaGetLastError();
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;   // handle + "okay so far"
    std::vector<unsigned char> mine_rec(rec, 0), all((size_t)c->world * rec, 0);
    memcpy(mine_rec.data(), &h, sizeof h);
    mine_rec[sizeof h] = (unsigned char)okay;
    if (cudaMalloc((void **)&d_h, (size_t)(c->world + 1) * rec) != cudaSuccess) { cudaGetLastError(); if (mine) cudaFree(mine); return false; }
    bool comm_ok = cudaMemcpyAsync(d_h, mine_rec.data(), rec, cudaMemcpyHostToDevice, st) == cudaSuccess &&
                   N->AllGather(d_h, d_h + rec, rec, kNcclUint8, c->comm, st) == 0 &&
                   cudaMemcpyAsync(all.data(), d_h + rec, (size_t)c->world * rec, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                   cudaStreamSynchronize(st) == cudaSuccess;
    for (int r = 0; comm_ok && r < c->world; r++) okay &= all[(size_t)r * rec + sizeof h];
    // second round: did every rank manage to map every buffer?
    int mapped = comm_ok && okay;
    for (int r = 0; r < kMaxPeers; r++) x.ptrs.buf[r] = nullptr;
    if (mapped) {
        x.ptrs.buf[c->rank] = mine;
        for (int r = 0; r < c->world && mapped; r++) {
            if (r == c->rank) continue;
            cudaIpcMemHandle_t hr;
            memcpy(&hr, all.data() + (size_t)r * rec, sizeof hr);
            void *ptr = nullptr;
            if ((ce = cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) {
                if (dbg) fprintf(stderr, "[dc_shard] rank %d: cudaIpcOpenMemHandle(rank %d): %s\n", c->rank, r, cudaGetErrorString(ce));
                cudaGetLastError();
                mapped = 0;
            }
            else x.ptrs.buf[r] = (unsigned long long *)ptr;
        }
    }
    if (comm_ok) {
        mine_rec[0] = (unsigned char)mapped;
        comm_ok = cudaMemcpyAsync(d_h, mine_rec.data(), 1, cudaMemcpyHostToDevice, st) == cudaSuccess &&
                  N->AllGather(d_h, d_h + rec, 1, kNcclUint8, c->comm, st) == 0 &&
                  cudaMemcpyAsync(all.data(), d_h + rec, (size_t)c->world, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                  cudaStreamSynchronize(st) == cudaSuccess;
        for (int r = 0; comm_ok && r < c->world; r++) mapped &= all[r];
    }
    cudaFree(d_h);
    if (dbg) fprintf(stderr, "[dc_shard] rank %d: peer-memory exchange %s (collectives ok %d, buffers ok %d, mapped everywhere %d)\n", c->rank,
                     comm_ok && mapped ? "in use" : "unavailable, using ncclAllGather", (int)comm_ok, okay, mapped);
    if (!comm_ok || !mapped) {
        for (int r = 0; r < c->world; r++)
            if (r != c->rank && x.ptrs.buf[r]) cudaIpcCloseMemHandle(x.ptrs.buf[r]);
        if (mine) cudaFree(mine);
        for (int r = 0; r < kMaxPeers; r++) x.ptrs.buf[r] = nullptr;
        cudaGetLastError();
        return false;
    }
    x.flags_off = data_words;
    x.epoch = 0;
    x.state = 1;
    return true;
}
static void peer_teardown(dc_shard_comm *c) {
    PeerXchg &x = c->peer;
    if (x.state <= 0) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; r++)
        if (r != c->rank && x.ptrs.buf[r]) cudaIpcCloseMemHandle(x.ptrs.buf[r]);
    if (x.ptrs.buf[c->rank]) cudaFree(x.ptrs.buf[c->rank]);
    x.state = -1;
    cudaGetLastError();
}

// test / bench hook (not in the public header): 1 if this communicator exchanges through peer memory
extern "C" int dc_debug_shard_peer_active(const dc_shard_comm *c) { return c && c->peer.state > 0 ? 1 : 0; }

extern "C" int dc_shard_comm_destroy(dc_shard_comm *c) {
    if (!c) return DC_OK;
    peer_teardown(c);
    int rc = DC_OK;
    if (c->owned && nccl() && nccl()->CommDestroy(c->comm) != 0) rc = DC_ERR_NCCL;
    delete c;
    return rc;
}

extern "C" int dc_shard_comm_rank(const dc_shard_comm *c) { return c ? c->rank : DC_ERR_ARG; }
extern "C" int dc_shard_comm_world(const dc_shard_comm *c) { return c ? c->world : DC_ERR_ARG; }

// ------------------------------------------------------------------------------------------ encode

extern "C" size_t dc_shard_huff_encode_workspace_bytes(size_t n_local, int world) {
    return world < 1 ? 0 : shard_enc_layout(n_local, world).total;
}

extern "C" int dc_shard_huff_encode(dc_shard_comm *c, const uint8_t *d_in, size_t n_local, int n_ary, dc_huff_table *d_table,
                                    uint8_t *d_out, size_t out_capacity, uint64_t *d_total_bits, int32_t *d_status, void *d_workspace,
                                    size_t workspace_bytes, void *stream) {
    const NcclApi *N = nccl();
    if (!N) return DC_ERR_NCCL;
    if (!c || !d_table || !d_out || !d_workspace || (n_local && !d_in)) return DC_ERR_ARG;
    if ((((uintptr_t)d_in | (uintptr_t)d_out | (uintptr_t)d_workspace) & 15) != 0) return DC_ERR_ARG;
    const ShardEncLayout L = shard_enc_layout(n_local, c->world);
    if (workspace_bytes < L.total || out_capacity < 16) return DC_ERR_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    char *w = (char *)d_workspace;
    unsigned long long *local_hist = (unsigned long long *)(w + L.local_hist), *all_hist = (unsigned long long *)(w + L.all_hist);
    unsigned long long *rank_bits = (unsigned long long *)(w + L.rank_bits);
    unsigned long long *rank_off = (unsigned long long *)(w + L.rank_off);
    ShardPlan *plan = (ShardPlan *)(w + L.plan);
    void *enc_ws = w + L.enc_ws;
    const size_t enc_ws_bytes = L.total - L.enc_ws;

    // 1. local histogram (+ one small histogram per 32 KB run for the planned encoder, + the shard's edge symbols)
    int rc = histogram_runs_edges(d_in, n_local, local_hist, enc_ws, enc_ws_bytes, local_hist + DC_NSLOTS, st);
    if (rc != DC_OK) return rc;
    // 2. every rank gets every rank's histogram and edge symbols: the one exchange of the encode path -- through peer memory
    //    (stores over NVLink + flags, one single-CTA kernel) when the ranks could map each other's buffers, else ncclAllGather
    const bool peer = peer_setup(c, st);
    {
        LaunchScope ls(DC_K_SHARD_EXCHANGE, st);
        if (peer) {
            PeerXchg &x = c->peer;
            x.epoch++;
            shard_exchange_kernel<<<1, 320, 0, st>>>(x.ptrs, local_hist, c->rank, c->world, x.epoch, x.flags_off, &plan->reserved);
            all_hist = x.ptrs.buf[c->rank] + (size_t)(x.epoch & 1ull) * c->world * kShardSlots;
        } else {
            DC_NCCL_TRY(N->AllGather(local_hist, all_hist, kShardSlots, kNcclUint64, c->comm, st));
        }
    }
    // 3. global table from the sum of the histograms (redundantly on every rank), every rank's bit total, the exclusive
    //    scan, my phase
    {
        TableRaw raw = {nullptr, nullptr, nullptr, nullptr};
        rc = launch_table(all_hist, nullptr, DC_NSLOTS, n_ary, d_table, raw, st, c->world, kShardSlots);
        if (rc != DC_OK) return rc;
    }
    {
        LaunchScope ls(DC_K_SHARD_PLAN, st);
        shard_plan_kernel<<<1, 256, 0, st>>>(all_hist, c->world, c->rank, d_table, rank_bits, rank_off, plan);
    }
    // 4. encode at my phase (in device memory: the host never waits for the plan)
    if (n_local) {
        rc = encode_planned_device_phase(d_in, n_local, d_table, d_out, out_capacity, &plan->phase, d_total_bits, d_status, enc_ws, enc_ws_bytes, st);
        if (rc != DC_OK) return rc;
    } else {
        DC_CUDA_TRY(cudaMemsetAsync(d_out, 0, 16, st));   // a shard without bits still shows a defined first byte
        if (d_total_bits) DC_CUDA_TRY(cudaMemsetAsync(d_total_bits, 0, 8, st));
        if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, 4, st));
    }
    // 5. the bytes that neighbouring shards share, completed locally
    if (c->world > 1 && n_local) {
        LaunchScope ls(DC_K_SHARD_PLAN, st);
        shard_fix_edges_kernel<<<1, 32, 0, st>>>(d_out, all_hist, c->world, c->rank, d_table, rank_bits, rank_off, c->peer.state > 0 ? plan : nullptr, d_status);
    }
    return cuda_status(cudaGetLastError());
}

// blocking: where this rank's shard sits in the logical stream (valid after dc_shard_huff_encode on the same workspace)
extern "C" int dc_shard_huff_encode_info(const void *d_workspace, size_t n_local, int world, uint64_t *bit_offset, uint64_t *bits,
                                         uint64_t *total_bits, void *stream) {
    if (!d_workspace || world < 1) return DC_ERR_ARG;
    const ShardEncLayout L = shard_enc_layout(n_local, world);
    ShardPlan h;
    cudaStream_t st = (cudaStream_t)stream;
    DC_CUDA_TRY(cudaMemcpyAsync(&h, (const char *)d_workspace + L.plan, sizeof h, cudaMemcpyDeviceToHost, st));
    DC_CUDA_TRY(cudaStreamSynchronize(st));
    if (bit_offset) *bit_offset = h.bit_offset;
    if (bits) *bits = h.bits;
    if (total_bits) *total_bits = h.total_bits;
    return DC_OK;
}

// Places the shards in ONE contiguous buffer on `root` (BASELINE config 4): rank r's bytes go to d_stream + O_r / 8.  Blocking
// (the byte counts are read back first).  d_stream: only used on root, ceil(total_bits / 8) bytes.
extern "C" int dc_shard_huff_gather(dc_shard_comm *c, int root, const uint8_t *d_shard, const void *d_workspace, size_t n_local,
                                    uint8_t *d_stream, size_t stream_capacity, void *stream) {
    const NcclApi *N = nccl();
    if (!N) return DC_ERR_NCCL;
    if (!c || !d_workspace || root < 0 || root >= c->world) return DC_ERR_ARG;
    const ShardEncLayout L = shard_enc_layout(n_local, c->world);
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<unsigned long long> bits(c->world), off(c->world + 1);
    DC_CUDA_TRY(cudaMemcpyAsync(bits.data(), (const char *)d_workspace + L.rank_bits, (size_t)c->world * 8, cudaMemcpyDeviceToHost, st));
    DC_CUDA_TRY(cudaMemcpyAsync(off.data(), (const char *)d_workspace + L.rank_off, (size_t)(c->world + 1) * 8, cudaMemcpyDeviceToHost, st));
    DC_CUDA_TRY(cudaStreamSynchronize(st));
    if (c->rank == root && (!d_stream || stream_capacity < (size_t)((off[c->world] + 7) / 8))) return DC_ERR_CAPACITY;
    auto lo = [&](int r) { return (size_t)(off[r] >> 3); };
    auto nb = [&](int r) { return bits[r] ? (size_t)(((off[r] + bits[r] + 7) >> 3) - (off[r] >> 3)) : (size_t)0; };
    DC_NCCL_TRY(N->GroupStart());
    if (c->rank == root) {
        for (int r = 0; r < c->world; r++) {
            if (r == root || nb(r) == 0) continue;
            DC_NCCL_TRY(N->Recv(d_stream + lo(r), nb(r), kNcclUint8, r, c->comm, st));
        }
    } else if (nb(c->rank)) {
        if (!d_shard) { N->GroupEnd(); return DC_ERR_ARG; }
        DC_NCCL_TRY(N->Send(d_shard, nb(c->rank), kNcclUint8, root, c->comm, st));
    }
    DC_NCCL_TRY(N->GroupEnd());
    // (the shared bytes already hold the merged value on both sides, so the order of arrival does not matter)
    if (c->rank == root && nb(root)) DC_CUDA_TRY(cudaMemcpyAsync(d_stream + lo(root), d_shard, nb(root), cudaMemcpyDeviceToDevice, st));
    return DC_OK;
}

// ------------------------------------------------------------------------------------------ decode of one blindly cut stream

namespace dc {
constexpr size_t kHalo = 1024;
struct ShardDecLayout { size_t summaries, dec_ws, total; };
static ShardDecLayout shard_dec_layout(size_t part_bytes, int world) {
    ShardDecLayout L;
    size_t p = 0;
    auto take = [&](size_t bytes) { size_t o = p; p += (bytes + 255) & ~(size_t)255; return o; };
    L.summaries = take((size_t)(world + 1) * sizeof(dc_shard_summary));
    L.dec_ws = take(dc_huff_decode_workspace_bytes(0, (uint64_t)(part_bytes + kHalo) * 8));
    L.total = p;
    return L;
}
}  // namespace dc

extern "C" size_t dc_shard_huff_decode_workspace_bytes(size_t part_bytes, int world) {
    return world < 1 ? 0 : shard_dec_layout(part_bytes, world).total;
}

/*
 * d_buf: [1024 bytes headroom | this rank's bytes of the stream | >= 1024 bytes tailroom], 16-byte aligned; rank r holds
 * stream bytes [r * part_bytes, ...) -- part_bytes a multiple of 1024, the same on every rank; the last ranks may hold
 * less or nothing.  The halos are filled here.  Blocking (the ranks' summaries decide what happens next).
 * On return: *n_symbols symbols in d_out, which belong at *symbol_offset of the whole output.
 */
extern "C" int dc_shard_huff_decode_stream(dc_shard_comm *c, uint8_t *d_buf, size_t part_bytes, uint64_t total_bits,
                                           const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity, uint64_t *n_symbols,
                                           uint64_t *symbol_offset, uint64_t *total_symbols, int32_t *d_status, void *d_workspace,
                                           size_t workspace_bytes, void *stream) {
    const NcclApi *N = nccl();
    if (!N) return DC_ERR_NCCL;
    if (!c || !d_buf || !d_table || !d_workspace || part_bytes == 0 || part_bytes % kHalo) return DC_ERR_ARG;
    if ((((uintptr_t)d_buf | (uintptr_t)d_workspace) & 15) != 0) return DC_ERR_ARG;
    const ShardDecLayout L = shard_dec_layout(part_bytes, c->world);
    if (workspace_bytes < L.total) return DC_ERR_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    char *w = (char *)d_workspace;
    dc_shard_summary *d_sum = (dc_shard_summary *)(w + L.summaries);   // [world] gathered, [world] = mine
    dc_shard_summary *d_mine = d_sum + c->world;
    void *dec_ws = w + L.dec_ws;
    const size_t dec_ws_bytes = L.total - L.dec_ws;
    const int G = c->world, r = c->rank;
    const uint64_t total_bytes = (total_bits + 7) / 8;
    auto range_lo = [&](int g) { return (uint64_t)g * part_bytes < total_bytes ? (uint64_t)g * part_bytes : total_bytes; };
    auto range_n = [&](int g) { const uint64_t lo = range_lo(g); return lo + part_bytes < total_bytes ? (uint64_t)part_bytes : total_bytes - lo; };
    const uint64_t mine = range_n(r), lo = range_lo(r);
    uint8_t *part = d_buf + kHalo;
    // halos: my tail to the right neighbour's headroom, my head to the left neighbour's tailroom
    DC_NCCL_TRY(N->GroupStart());
    if (r + 1 < G && range_n(r + 1) > 0 && mine >= kHalo) DC_NCCL_TRY(N->Send(part + mine - kHalo, kHalo, kNcclUint8, r + 1, c->comm, st));
    if (r > 0 && mine > 0) DC_NCCL_TRY(N->Recv(d_buf, kHalo, kNcclUint8, r - 1, c->comm, st));
    if (r > 0 && mine > 0) DC_NCCL_TRY(N->Send(part, mine < kHalo ? (size_t)mine : kHalo, kNcclUint8, r - 1, c->comm, st));
    if (r + 1 < G && range_n(r + 1) > 0) DC_NCCL_TRY(N->Recv(part + mine, range_n(r + 1) < kHalo ? (size_t)range_n(r + 1) : kHalo, kNcclUint8, r + 1, c->comm, st));
    DC_NCCL_TRY(N->GroupEnd());

    const uint64_t left = mine ? total_bits - 8 * lo : 0, my_bits = mine ? (8 * mine < left ? 8 * mine : left) : 0;
    DC_CUDA_TRY(cudaMemsetAsync(d_mine, 0, sizeof(dc_shard_summary), st));
    int has_halo = r > 0 ? 1 : 0;
    if (my_bits) {
        const int rc = dc_huff_decode_shard_sync(part, has_halo, 0, my_bits, left, d_table, d_mine, dec_ws, dec_ws_bytes, stream);
        if (rc != DC_OK) return rc;
    }
    std::vector<dc_shard_summary> h(G);
    std::vector<int> active;
    for (int g = 0; g < G; g++)
        if (range_n(g) > 0) active.push_back(g);
    for (int iter = 0;; iter++) {
        DC_NCCL_TRY(N->AllGather(d_mine, d_sum, sizeof(dc_shard_summary), kNcclUint8, c->comm, st));
        DC_CUDA_TRY(cudaMemcpyAsync(h.data(), d_sum, (size_t)G * sizeof(dc_shard_summary), cudaMemcpyDeviceToHost, st));
        DC_CUDA_TRY(cudaStreamSynchronize(st));
        bool any_wrong = false, i_am_wrong = false;
        uint32_t my_prev_exit = 0;
        for (size_t i = 0; i < active.size(); i++) {
            if (h[active[i]].resync) return DC_ERR_CORRUPT;   // does not self-synchronise within 8192 bits: decode on one device
            if (i > 0 && h[active[i]].assumed_start != h[active[i - 1]].exit) {
                any_wrong = true;
                if (active[i] == r) { i_am_wrong = true; my_prev_exit = h[active[i - 1]].exit; }
            }
        }
        if (!any_wrong) break;
        if (iter > G) return DC_ERR_CORRUPT;
        if (i_am_wrong) {   // start over from the exact place my neighbour reports
            has_halo = 0;
            const int rc = dc_huff_decode_shard_sync(part, 0, my_prev_exit, my_bits, left, d_table, d_mine, dec_ws, dec_ws_bytes, stream);
            if (rc != DC_OK) return rc;
        }
    }
    uint64_t off = 0, tot = 0;
    for (int g : active) {
        if (g < r) off += h[g].symbols;
        tot += h[g].symbols;
    }
    const uint64_t nsym = my_bits ? h[r].symbols : 0;
    if (n_symbols) *n_symbols = nsym;
    if (symbol_offset) *symbol_offset = off;
    if (total_symbols) *total_symbols = tot;
    if (d_status) DC_CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    if (nsym > out_capacity) return DC_ERR_CAPACITY;
    if (nsym) {
        const int rc = dc_huff_decode_shard_write(part, has_halo, my_bits, left, d_table, d_out, (size_t)nsym, d_status, dec_ws, dec_ws_bytes, stream);
        if (rc != DC_OK) return rc;
    }
    return DC_OK;
}

// ------------------------------------------------------------------------------------------ nybble shards

// symbols [lo, hi) of n_total that rank `rank` of `world` takes: every shard starts on an even symbol index, so that no
// packed byte is shared (write_nybble nybble_compression.c:1091-1114 puts symbols 2i and 2i + 1 into byte i)
extern "C" int dc_shard_nybble_range(uint64_t n_total, int rank, int world, uint64_t *lo, uint64_t *hi) {
    if (world < 1 || rank < 0 || rank >= world || !lo || !hi) return DC_ERR_ARG;
    const uint64_t pairs = (n_total + 1) / 2, per = (pairs + world - 1) / world;
    const uint64_t a = 2 * per * (uint64_t)rank, b = 2 * per * (uint64_t)(rank + 1);
    *lo = a < n_total ? a : n_total;
    *hi = b < n_total ? b : n_total;
    return DC_OK;
}
// rank-local calls on the shard's symbols / packed bytes (d_sym, d_packed point at the shard's own first symbol / byte)
extern "C" int dc_shard_nybble_pack(uint64_t n_total, int rank, int world, const uint8_t *d_sym, uint8_t *d_packed, int32_t *d_status,
                                    void *stream) {
    uint64_t lo, hi;
    const int rc = dc_shard_nybble_range(n_total, rank, world, &lo, &hi);
    if (rc != DC_OK) return rc;
    return dc_nybble_pack(d_sym, (size_t)(hi - lo), d_packed, d_status, stream);
}
extern "C" int dc_shard_nybble_unpack(uint64_t n_total, int rank, int world, const uint8_t *d_packed, uint8_t *d_sym, void *stream) {
    uint64_t lo, hi;
    const int rc = dc_shard_nybble_range(n_total, rank, world, &lo, &hi);
    if (rc != DC_OK) return rc;
    return dc_nybble_unpack(d_packed, (size_t)(hi - lo), d_sym, stream);
}
