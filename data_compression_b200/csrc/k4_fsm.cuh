// k4_fsm.cuh -- K4, byte-stepped decoder: the code as a finite-state machine that eats one BYTE per table look-up.
// (Included by k4_decode.cu, which owns the workspace, the scan kernel F2 and the entry points.)
//
// Why.  The window decoders in k4_decode.cu spend ~8 instructions per look-up (funnel shift, index, address, LDS, position
// update, count update, compare, branch), most of them on the half-rate integer ALU pipe, and every 32-bit word of a
// lane's subsequence is its own data-dependent loop, so a warp runs the slowest lane's trip count (ncu, round 1: 555 M +
// 604 M warp instructions per GiB, ALU pipe 55-60 %, a third of the lane slots idle).  A prefix code read from a byte
// boundary is fully described by WHERE INSIDE A CODE the boundary falls -- one of the internal nodes of the code tree,
// at most 256 of them for byte alphabets.  So the tables here are indexed by (state, next byte):
//   F1  u16  number of codes that END in this byte | next state << 8
//   F3  u32  symbol 0 | symbol 1 << 8 | next state << 16 | k << 24 | count << 27   (a random 8-byte gather costs 5.9
//            shared-memory wavefronts, a 4-byte one 3.1, and the walks are bound by exactly those bank-conflict replays, so
//            the third and fourth symbol of a byte -- rare -- are not in the entry: k is the digit at which the third code
//            starts, and what a byte decodes to FROM THE ROOT from digit k on is a table of 8 x 256 u16 for the whole code,
//            read only by the few lanes whose byte completes more than two codes)
// and a walk is 32 fixed, fully unrolled steps per 256-bit subsequence: no positions, no windows, no look-ahead word,
// no loop condition, no divergence, codes of any length at the same speed.  F1: PRMT (index = state, byte) + IMAD +
// LDS + IADD per byte.  F3: the symbols of a step are appended to a 4-byte sliding window with one PRMT whose selector
// comes from the table, completed 32-bit words go to the staging tile with a predicated STS (so the shared-memory
// stores are words, not bytes), and the fill state advances with one IMAD: 4 ALU + 3 FMA + 2 LSU slots per byte.
//
// States.  Canonical n-ary code (n = 2, 3, 4, 16; a digit is bpd bits -- radix 3 runs on the stream with one 2-bit field per
// trit, where a field of 3 is simply an unused slot): at depth d (digits) the values
// [first[d], first[d] + count[d]) are leaves, [ilo[d], ihi[d]] = [first[d] + count[d], last / n ^ (max_len - d)] are
// internal nodes, everything above is an unused slot (the reference's dummy leaves).  State id = base[d] + (v -
// ilo[d]): breadth first, the root is 0.  Streams whose codes start on digit boundaries relative to the BYTE grid only
// ever stop on such nodes at byte boundaries (bit_start % bpd == 0; anything else takes the window kernels).
// An unused slot sends F1 back to the root (legal on speculative paths, as in the window kernels) and F3, which only
// walks true paths, into an absorbing DEAD state that is reported as DC_ERR_CORRUPT.
//
// Eligibility (fsm_geometry): bpd in {1, 2, 4}, shortest code >= 2 bits (at most 4 symbols per byte, 128 per
// subsequence), <= 255 internal nodes, lengths < 16 digits.  Everything else (1-bit codes, binary codes of full byte
// alphabets) keeps the window kernels.  The first tile of a stream that does not start on a byte boundary and the ragged last tile are walked digit
// by digit by a generic routine (two tiles per call).
//
// Binary codes of full byte alphabets (257 leaves with the reference's dummy: 256 internal nodes) have an F1 table of 128 KB,
// which fits, and F3 rows of 256 KB, which do not.  They take F1 here and the window kernel for F3 (COMPAT): F1 then stores
// what the window kernel expects -- the bit offset of the first code that STARTS in the subsequence and the number of codes
// that start in it -- which it gets from its own results: a lane that begins inside a code finds the end of that code
// with at most two more look-ups (codes are shorter than 16 bits; the entry also holds the bit at which the first code of
// the byte ends), and codes that start = codes that end - (began inside a code) + (ends inside a code).
#pragma once

namespace dc {

constexpr int kFsmMaxDepth = 32;
constexpr uint32_t kFsmToken = 0x80000000u;   // start / exit tokens of the FSM path: kFsmToken | state (else: a bit offset)

struct FsmHeader {
    int32_t nstates, bpd, n_ary, min_len, max_len, reserved[3];
    uint32_t first[kFsmMaxDepth], count[kFsmMaxDepth], off[kFsmMaxDepth];
    uint32_t ilo[kFsmMaxDepth], ihi[kFsmMaxDepth], base[kFsmMaxDepth];
    uint16_t sorted[DC_NSLOTS + 1];
};
constexpr size_t kFsmHeaderBytes = (sizeof(FsmHeader) + 255) & ~(size_t)255;
constexpr size_t kFsmSyncRowBytes = 256 * sizeof(uint16_t);            // F1: one state
constexpr size_t kFsmWriteRowBytes = 256 * sizeof(uint32_t);           // F3: one state, symbols 0 and 1
constexpr int kFsmEntryRows = 7;                                       // rows nstates + k: a stream that starts at digit k of its first byte
constexpr int kFsmSuffixRows = 16;                                     // F3: (count - 3) * 8 + k, k = first digit of the third code
constexpr size_t kFsmWriteXRowBytes = 256 * sizeof(uint16_t);          // F3: one suffix row, symbols 2 and 3
constexpr size_t kFsmSyncTableBytes = kFsmMaxSyncStates * kFsmSyncRowBytes;
constexpr size_t kFsmWriteTableBytes = (kFsmMaxStates + 1) * kFsmWriteRowBytes;
constexpr size_t kFsmWriteXTableBytes = kFsmSuffixRows * kFsmWriteXRowBytes;
constexpr size_t kFsmWorkspaceBytes = kFsmHeaderBytes + kFsmSyncTableBytes + kFsmWriteTableBytes + kFsmWriteXTableBytes;

struct FsmTables {   // where the three pieces live in the decode workspace
    FsmHeader *hdr;
    uint16_t *sync;
    uint32_t *write;
    uint16_t *writex;
};
static inline FsmTables fsm_tables_at(void *p) {
    char *c = (char *)p;
    FsmTables t;
    t.hdr = (FsmHeader *)c;
    t.sync = (uint16_t *)(c + kFsmHeaderBytes);
    t.write = (uint32_t *)(c + kFsmHeaderBytes + kFsmSyncTableBytes);
    t.writex = (uint16_t *)(c + kFsmHeaderBytes + kFsmSyncTableBytes + kFsmWriteTableBytes);
    return t;
}

// one digit x from node (d, v): a symbol (>= 0), nothing (-1, inside a code) or an unused slot (-2); (d, v) = next node
__device__ __forceinline__ int fsm_digit(const FsmHeader *h, int &d, uint32_t &v, uint32_t x) {
    if (x >= (uint32_t)h->n_ary) {   // not a digit of this radix (a 2-bit field of 3 in a radix-3 stream)
        d = 0;
        v = 0;
        return -2;
    }
    d++;
    v = v * (uint32_t)h->n_ary + x;
    if (d >= h->min_len) {
        const uint32_t c = h->count[d], f = h->first[d];
        if (c && v >= f && v - f < c) {
            const int sym = (int)(h->sorted[h->off[d] + (v - f)] & 0xFFu);
            d = 0;
            v = 0;
            return sym;
        }
    }
    if (d < h->max_len && v >= h->ilo[d] && v <= h->ihi[d]) return -1;
    d = 0;
    v = 0;
    return -2;
}
__device__ __forceinline__ uint32_t fsm_state_id(const FsmHeader *h, int d, uint32_t v) { return h->base[d] + (v - h->ilo[d]); }
__device__ __forceinline__ void fsm_state_node(const FsmHeader *h, uint32_t id, int &d, uint32_t &v) {
    int dd = 0;
    while (dd + 1 < h->max_len && h->base[dd + 1] <= id) dd++;
    d = dd;
    v = h->ilo[dd] + (id - h->base[dd]);
}

// ------------------------------------------------------------------------------------------ table build
// grid = nstates + 1 + kFsmEntryRows + kFsmSuffixRows CTAs of 256 threads: CTA s fills row s of both tables (thread = byte
// value); row nstates is DEAD; rows nstates + k (k = 1 .. 7, tables with at most 248 states) are ENTRY rows -- the root, but
// the first k digits of the byte belong to somebody else (a shard that starts at bit phase k * bpd of its first byte begins
// in that state and needs no digit-by-digit walk of its first tile); the CTAs behind them fill the suffix rows.
__host__ __device__ inline bool fsm_has_entry_rows(int nstates) { return nstates > 0 && nstates + kFsmEntryRows <= 255; }
__global__ void __launch_bounds__(256) fsm_build_kernel(const dc_huff_table *__restrict__ tab, FsmTables t, int sync_only, int expected_states) {
    __shared__ FsmHeader h;
    const int tid = threadIdx.x;
    for (int i = tid; i < kFsmMaxDepth; i += 256) {
        h.first[i] = tab->first_code[i];
        h.count[i] = tab->len_count[i];
        h.off[i] = tab->len_offset[i];
        h.ilo[i] = h.ihi[i] = h.base[i] = 0;
    }
    for (int i = tid; i <= DC_NSLOTS; i += 256) h.sorted[i] = tab->sorted[i];
    __syncthreads();
    if (tid == 0) {
        h.bpd = tab->bits_per_digit;
        h.n_ary = tab->n_ary;
        h.min_len = tab->min_len;
        h.max_len = tab->max_len;
        // lengths outside [min_len, max_len] have no codes (K2 scans skip the last slot, as the reference does)
        for (int d = 0; d < kFsmMaxDepth; d++)
            if (d < h.min_len || d > h.max_len) h.count[d] = 0;
        h.nstates = (tab->status == DC_OK && h.bpd != 0)
                        ? fsm_geometry(h.first, h.count, h.min_len, h.max_len, h.bpd, h.n_ary, h.ilo, h.ihi, h.base) : 0;
        // the host sized this launch and the walks' shared memory from the table header it knows; a header that is not this
        // table's (dc_huff_table_forget was owed) must not be walked with: no states = the walks return, F2 flags the stream,
        // the window kernels -- which only trust the table itself -- redo it
        if (h.nstates != expected_states) h.nstates = 0;
        h.reserved[0] = h.reserved[1] = h.reserved[2] = 0;
    }
    __syncthreads();
    const int ns = h.nstates;
    if (blockIdx.x == 0) {
        uint32_t *dst = (uint32_t *)t.hdr;
        const uint32_t *src = (const uint32_t *)&h;
        for (int i = tid; i < (int)(sizeof(FsmHeader) / 4); i += 256) dst[i] = src[i];
    }
    const int s = blockIdx.x;
    if (ns == 0 || s > ns + kFsmEntryRows + kFsmSuffixRows) return;
    if (sync_only && s >= ns) return;   // F1 rows only (the F3 rows of 256 states would not fit anyway)
    const int bpd = h.bpd, steps = 8 / bpd;
    const uint32_t mask = (1u << bpd) - 1u;
    if (s == ns) {  // DEAD: absorbs everything, emits nothing
        t.write[(size_t)s * 256 + tid] = (uint32_t)ns << 16;
        return;
    }
    if (s > ns + kFsmEntryRows) {   // suffix row: what byte `tid` decodes to from the root, from digit k on (the third and fourth code of a byte)
        const int row = s - ns - kFsmEntryRows - 1, k0 = row & 7;
        int d = 0;
        uint32_t v = 0, cnt = 0, syms = 0;
        for (int k = k0; k < steps; k++) {
            const int r = fsm_digit(&h, d, v, ((uint32_t)tid >> (8 - bpd * (k + 1))) & mask);
            if (r == -2) break;
            if (r >= 0 && cnt < 2) { syms |= (uint32_t)r << (8 * cnt); cnt++; }
        }
        t.writex[(size_t)row * 256 + tid] = (uint16_t)syms;
        return;
    }
    int d0 = 0, kstart = 0;
    uint32_t v0 = 0;
    if (s > ns) {   // entry row: from the root, from digit s - ns on
        kstart = s - ns;
        if (!fsm_has_entry_rows(ns) || kstart >= steps) return;
    } else {
        fsm_state_node(&h, (uint32_t)s, d0, v0);
    }
    {   // F1: an unused slot sends the walk back to the root
        int d = d0;
        uint32_t v = v0, cnt = 0, first_end = 0;
        for (int k = kstart; k < steps; k++) {
            const int r = fsm_digit(&h, d, v, ((uint32_t)tid >> (8 - bpd * (k + 1))) & mask);
            if (r >= 0 && cnt == 0) first_end = (uint32_t)((k + 1) * bpd - 1);   // last bit of the first code that ends in this byte
            cnt += r >= 0;
        }
        // (COMPAT reads bits 5..7; the plain walk adds whole entries up, so they are only set for it)
        t.sync[(size_t)s * 256 + tid] = (uint16_t)(cnt | (sync_only ? first_end << 5 : 0u) | (fsm_state_id(&h, d, v) << 8));
    }
    if (sync_only) return;
    {   // F3: an unused slot is the end of the true path
        int d = d0;
        uint32_t v = v0, cnt = 0, syms = 0, next = 0, third = 0;
        bool dead = false;
        for (int k = kstart; k < steps && !dead; k++) {
            const int r = fsm_digit(&h, d, v, ((uint32_t)tid >> (8 - bpd * (k + 1))) & mask);
            if (r >= 0) {
                if (cnt < 2) syms |= (uint32_t)r << (8 * cnt);
                cnt++;
                if (cnt == 2) third = (uint32_t)(k + 1);   // the third code, if any, starts at the next digit
            }
            dead = r == -2;
        }
        next = dead ? (uint32_t)ns : fsm_state_id(&h, d, v);
        t.write[(size_t)s * 256 + tid] = (syms & 0xFFFFu) | (next << 16) | ((cnt >= 3 ? third & 7u : 0u) << 24) | (cnt << 27);
    }
}

// ------------------------------------------------------------------------------------------ helpers

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ uint32_t lds_u16_at(uint32_t base, uint32_t idx) {
    uint32_t v;
    asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %1, 2, %2;\n\tld.shared.u16 %0, [a];\n\t}" : "=r"(v) : "r"(idx), "r"(base));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32_at(uint32_t base, uint32_t idx) {
    uint32_t v;
    asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %1, 4, %2;\n\tld.shared.u32 %0, [a];\n\t}" : "=r"(v) : "r"(idx), "r"(base));
    return v;
}

// bit i (0 = first bit of the stream buffer, MSB first within a byte) .. as `bpd` bits; the caller keeps i + bpd <= limit
__device__ __forceinline__ uint32_t stream_digit(const uint8_t *__restrict__ bits, unsigned long long i, int bpd) {
    const uint32_t b = __ldg(bits + (i >> 3));
    return (b >> (8 - bpd - (int)(i & 7))) & ((1u << bpd) - 1u);
}

// generic walk of bits [p, lim) of the stream from node (d, v); symbols go to `sink(sym)`.  Returns the symbol count;
// `dead` is set when an unused slot was met.  (The slow tiles: the first one of a stream that starts inside a byte, and
// the last one.)
template <typename Sink>
__device__ __forceinline__ uint32_t fsm_walk_digits(const FsmHeader *h, const uint8_t *__restrict__ bits, unsigned long long p,
                                                    unsigned long long lim, int &d, uint32_t &v, bool &dead, Sink sink) {
    uint32_t cnt = 0;
    const int bpd = h->bpd;
    while (p + bpd <= lim) {
        const int r = fsm_digit(h, d, v, stream_digit(bits, p, bpd));
        if (r >= 0) { sink((uint32_t)r, cnt); cnt++; }
        if (r == -2) dead = true;
        p += bpd;
    }
    return cnt;
}

// a warp's view of its segment: one 32-byte load per lane, one tile ahead (little-endian words as loaded: byte 0 of a
// word is the first byte of the stream, which is the order the byte steps want)
struct FsmCursor {
    const uint4 *src;
    unsigned long long vec_left;   // 16-byte vectors of the stream from the segment's first tile on
    uint4 n0, n1;
    uint32_t fetched;
    int lane;
    __device__ __forceinline__ void init(const uint8_t *d_bits, unsigned long long first_tile, unsigned long long nvec, int lane_) {
        lane = lane_;
        const unsigned long long vec0 = first_tile * kF_TileVecs;
        src = (const uint4 *)d_bits + vec0 + 2 * lane;
        vec_left = nvec > vec0 ? nvec - vec0 : 0;
        fetched = 0;
        fetch();
    }
    __device__ __forceinline__ void fetch() {
        n0 = make_uint4(0, 0, 0, 0);
        n1 = make_uint4(0, 0, 0, 0);
        const unsigned long long i = (unsigned long long)fetched * kF_TileVecs + 2 * lane;
        if (i + 1 < vec_left) ldg_256(src, n0, n1);
        else if (i < vec_left) n0 = ldg_stream(src);
        src += kF_TileVecs;
        fetched++;
    }
    // the lane's eight words of the current tile; prefetches the next one only if the warp will walk it
    __device__ __forceinline__ void take(uint32_t (&w)[8], bool more) {
        w[0] = n0.x; w[1] = n0.y; w[2] = n0.z; w[3] = n0.w;
        w[4] = n1.x; w[5] = n1.y; w[6] = n1.z; w[7] = n1.w;
        if (more) fetch();
    }
};


// The same view with the tile staged in shared memory by the bulk-copy engine (TMA, cp.async.bulk + mbarrier) -- the
// staging north_star (4) names.  One 1 KB buffer and one mbarrier per warp: lane 0 arms the barrier with the byte count and
// issues the copy of the NEXT tile as soon as every lane has moved the current one into registers (two 16-byte LDS per
// lane; the lanes of the odd groups of four read their halves in the other order, which makes the 32-byte stride
// conflict-free), so the copy runs under the current tile's walk and no registers hold data in flight.
struct FsmCursorTma {
    const uint8_t *src;            // next tile to fetch
    unsigned long long bytes_left; // readable bytes (whole 16-byte vectors) from `src` on
    uint32_t buf, bar;             // shared-memory addresses: this warp's tile buffer and its mbarrier
    uint32_t phase;
    bool pending;
    int lane;
    __device__ __forceinline__ void init(const uint8_t *d_bits, unsigned long long first_tile, unsigned long long nvec, int lane_,
                                         uint32_t buf_, uint32_t bar_, uint32_t phase_) {
        lane = lane_;
        buf = buf_;
        bar = bar_;
        phase = phase_;
        const unsigned long long vec0 = first_tile * kF_TileVecs;
        src = d_bits + vec0 * 16;
        bytes_left = nvec > vec0 ? (nvec - vec0) * 16 : 0;
        pending = false;
        fetch();
    }
    __device__ __forceinline__ void fetch() {
        const uint32_t bytes = (uint32_t)min(bytes_left, (unsigned long long)(kF_TileVecs * 16));
        pending = bytes != 0;
        if (pending && lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lanes' reads of the buffer precede the engine's writes
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(buf), "l"(src), "r"(bytes), "r"(bar) : "memory");
        }
        src += kF_TileVecs * 16;
        bytes_left -= bytes;
    }
    __device__ __forceinline__ void take(uint32_t (&w)[8], bool more) {
        if (pending) {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(phase) : "memory");
            phase ^= 1u;
        }
        const uint32_t swap = (lane >> 2) & 1u, a0 = buf + 32u * lane + 16u * swap;
        uint4 r0, r1;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(a0) : "memory");
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(a0 ^ 16u) : "memory");
        const uint4 lo = swap ? r1 : r0, hi = swap ? r0 : r1;
        w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w;
        w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
        __syncwarp();
        if (more) fetch();
    }
};

// ------------------------------------------------------------------------------------------ F1 (FSM)

// replace byte k of `word` by byte `src_byte` of `from`
template <int K, int SRC>
__device__ __forceinline__ uint32_t put_byte_from(uint32_t word, uint32_t from) {
    constexpr uint32_t sel = (0x3210u & ~(0xFu << (4 * K))) | ((uint32_t)(4 + SRC) << (4 * K));
    return prmt(word, from, sel);
}

// Walk the lane's 32 bytes from state `st`.  Record: chk byte k = state after word k, wc byte k = codes that ended in word k.
// FIRST: whole subsequence.  !FIRST: re-walk until the state at a word end equals the recorded one (same path from there on).
template <bool FIRST>
__device__ __forceinline__ void fsm_sync_walk(uint32_t tab, const uint32_t (&w)[8], uint32_t st, bool merged, uint32_t (&chk)[2],
                                              uint32_t (&wc)[2]) {
    uint32_t e = st << 8;   // the previous entry: state in byte 1
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (!FIRST && !__any_sync(0xFFFFFFFFu, !merged)) return;
        if (FIRST || !merged) {
            uint32_t cs;
            e = lds_u16_at(tab, prmt(w[k], e, 0x7750u));
            cs = e;
            e = lds_u16_at(tab, prmt(w[k], e, 0x7751u));
            cs += e;
            e = lds_u16_at(tab, prmt(w[k], e, 0x7752u));
            cs += e;
            e = lds_u16_at(tab, prmt(w[k], e, 0x7753u));
            cs += e;
            const int h = k >> 2;
            const uint32_t old = chk[h];
            uint32_t neu, nwc;
            switch (k & 3) {
                case 0: neu = put_byte_from<0, 1>(old, e); nwc = put_byte_from<0, 0>(wc[h], cs); break;
                case 1: neu = put_byte_from<1, 1>(old, e); nwc = put_byte_from<1, 0>(wc[h], cs); break;
                case 2: neu = put_byte_from<2, 1>(old, e); nwc = put_byte_from<2, 0>(wc[h], cs); break;
                default: neu = put_byte_from<3, 1>(old, e); nwc = put_byte_from<3, 0>(wc[h], cs); break;
            }
            chk[h] = neu;
            wc[h] = nwc;
            if (!FIRST && k < 7 && neu == old) merged = true;
        }
    }
}

struct FsmSyncArgs {
    const uint8_t *d_bits;
    unsigned long long end;        // bits of the stream from d_bits on
    unsigned long long nsub, ntiles, nseg;
    uint32_t start_token;          // first segment (lead == 0): kFsmToken | state, or a bit offset (< 256) of the first code
    int lead;
    const DecodeChain *chain;
    uint32_t tma_table_bytes;      // TMA staging only: the table's share of shared memory, rounded up to 128 bytes
};

// slow tile: every lane walks its part digit by digit.  start: kFsmToken | state for a lane that begins at bit 0 of its
// subsequence, else the bit offset (lane 0 only) of the first code.
// first_end: bits from the lane's first bit to the end of the first code that ends in its part (lim if none does)
__device__ __forceinline__ void fsm_slow_sync_lane(const FsmHeader *h, const uint8_t *__restrict__ d_bits, unsigned long long sub_bit0,
                                                   uint32_t lim, uint32_t start, uint32_t &cnt, uint32_t &exit_state, uint32_t &first_end) {
    int d = 0;
    uint32_t v = 0;
    unsigned long long p = sub_bit0;
    if (!(start & kFsmToken)) p += start;
    else if ((start & 0x1FFu) > (uint32_t)h->nstates) p += ((start & 0x1FFu) - (uint32_t)h->nstates) * (uint32_t)h->bpd;   // an entry row: the root, that many digits in
    else fsm_state_node(h, start & 0x1FFu, d, v);
    cnt = 0;
    first_end = lim;
    const int bpd = h->bpd;
    const unsigned long long stop = sub_bit0 + lim;
    while (p + bpd <= stop) {
        const int r = fsm_digit(h, d, v, stream_digit(d_bits, p, bpd));
        p += bpd;
        if (r >= 0) {
            if (cnt == 0) first_end = (uint32_t)(p - sub_bit0);
            cnt++;
        }
    }
    exit_state = fsm_state_id(h, d, v);
}

template <bool TMA, bool COMPAT>
__global__ void __launch_bounds__(1024, 1) fsm_sync_kernel(FsmSyncArgs a, FsmTables t, FastWorkspace ws) {
    extern __shared__ __align__(16) uint8_t fsm_smem[];
    FsmHeader *s_h = (FsmHeader *)fsm_smem;
    uint16_t *s_tab = (uint16_t *)(fsm_smem + kFsmHeaderBytes);
    {
        const uint32_t *src = (const uint32_t *)t.hdr;
        uint32_t *dst = (uint32_t *)s_h;
        for (int i = threadIdx.x; i < (int)(sizeof(FsmHeader) / 4); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        if (s_h->nstates == 0) {   // fsm_build_kernel refused the table (stale header): F2 sends the stream to the robust path
            if (blockIdx.x == 0 && threadIdx.x == 0) *ws.bad_input() = 1;
            return;
        }
        const int words = (s_h->nstates + (!COMPAT && fsm_has_entry_rows(s_h->nstates) ? 1 + kFsmEntryRows : 0)) * 128;
        const uint4 *s4 = (const uint4 *)t.sync;
        uint4 *d4 = (uint4 *)s_tab;
        for (int i = threadIdx.x; i < words / 4; i += blockDim.x) d4[i] = s4[i];
        __syncthreads();
    }
    uint32_t tab = (uint32_t)__cvta_generic_to_shared(s_tab);
    asm volatile("" : "+r"(tab));
    uint32_t start_token = a.start_token;
    if (a.chain) start_token = a.chain->next_start;
    if (start_token == 0) start_token = kFsmToken;   // bit 0 of the first byte = the root at a byte boundary
    else if (!COMPAT && start_token < 8u && start_token % (uint32_t)s_h->bpd == 0u && fsm_has_entry_rows(s_h->nstates))
        start_token = kFsmToken | ((uint32_t)s_h->nstates + start_token / (uint32_t)s_h->bpd);   // inside the first byte: an entry row
    if (blockIdx.x == 0 && threadIdx.x == 0) *ws.start_slot() = start_token;   // F3 walks the first lane the same way
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const unsigned long long nvec = ((a.end + 7) / 8 + 15) / 16;
    const unsigned long long tile_bits = 32ull * kF_SubBits;
    // TMA: [header | table (a.tma_table_bytes, a multiple of 128) | one 1 KB tile buffer per warp | one mbarrier per warp]
    uint32_t tma_buf = 0, tma_bar = 0, tma_phase = 0;
    if (TMA) {
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(fsm_smem + kFsmHeaderBytes + a.tma_table_bytes);
        tma_buf = base + 1024u * warp;
        tma_bar = base + 1024u * warps + 8u * warp;
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tma_bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    for (unsigned long long seg = (unsigned long long)blockIdx.x * warps + warp; seg < a.nseg; seg += (unsigned long long)gridDim.x * warps) {
        const bool exact = seg == 0 && a.lead == 0;
        const int warm = exact ? 0 : 1;
        const unsigned long long tile0 = (unsigned long long)a.lead + seg * kF_SegTiles - warm;
        typename std::conditional<TMA, FsmCursorTma, FsmCursor>::type cur;
        if constexpr (TMA) cur.init(a.d_bits, tile0, nvec, lane, tma_buf, tma_bar, tma_phase);
        else cur.init(a.d_bits, tile0, nvec, lane);
        const uint32_t ntile = (uint32_t)min((unsigned long long)(kF_SegTiles + warm), a.ntiles - tile0);
        uint16_t *info = ws.sub_info + tile0 * 32 + lane;
        const unsigned long long sub_left = a.nsub - tile0 * 32;
        uint32_t carry = exact ? start_token : kFsmToken;   // token of the state in front of lane 0
        uint32_t assumed = carry, total = 0;
        for (uint32_t tt = 0; tt < ntile; tt++, info += 32) {
            uint32_t w[8];
            cur.take(w, tt + 1 < ntile);
            const unsigned long long tbit0 = (tile0 + tt) * tile_bits;
            const bool full = tbit0 + tile_bits <= a.end && (carry & kFsmToken);   // warp-uniform
            uint32_t start, cnt, exit_state;
            if (full) {
                start = lane == 0 ? (carry & 0xFFu) : 0u;
                uint32_t chk[2] = {0, 0}, wc[2] = {0, 0};
                fsm_sync_walk<true>(tab, w, start, false, chk, wc);
                while (true) {
                    uint32_t ns = __shfl_up_sync(0xFFFFFFFFu, chk[1] >> 24, 1);
                    if (lane == 0) ns = start;
                    const bool redo = ns != start;
                    if (!__any_sync(0xFFFFFFFFu, redo)) break;
                    start = ns;
                    fsm_sync_walk<false>(tab, w, start, !redo, chk, wc);
                }
                exit_state = chk[1] >> 24;
                if (COMPAT) { wc[0] &= 0x1F1F1F1Fu; wc[1] &= 0x1F1F1F1Fu; }   // bits 5..7 of an entry's low byte: where the first code ends
                cnt = __dp4a(wc[0], 0x01010101u, __dp4a(wc[1], 0x01010101u, 0u));
                if (COMPAT) {
                    // what the window kernel wants: where the first code that starts here starts, how many start here
                    uint32_t off = 0;
                    if (start != 0u) {
                        const uint32_t e0 = lds_u16_at(tab, prmt(w[0], start << 8, 0x7750u));
                        const uint32_t e1 = lds_u16_at(tab, prmt(w[0], e0, 0x7751u));
                        off = (e0 & 7u) ? ((e0 >> 5) & 7u) + 1u : 8u + ((e1 >> 5) & 7u) + 1u;
                    }
                    cnt = cnt - (start != 0u ? 1u : 0u) + (exit_state != 0u ? 1u : 0u);
                    start = off;
                }
            } else {
                const unsigned long long sub_bit0 = tbit0 + (unsigned long long)lane * kF_SubBits;
                const bool active = sub_bit0 < a.end;
                const uint32_t lim = active ? (uint32_t)min((unsigned long long)kF_SubBits, a.end - sub_bit0) : 0u;
                uint32_t st = lane == 0 ? carry : kFsmToken, first_end = 0;
                cnt = 0;
                exit_state = st & 0x1FFu;
                if (active) fsm_slow_sync_lane(s_h, a.d_bits, sub_bit0, lim, st, cnt, exit_state, first_end);
                while (true) {
                    uint32_t ns = __shfl_up_sync(0xFFFFFFFFu, exit_state, 1) | kFsmToken;
                    if (lane == 0) ns = st;
                    const bool redo = active && ns != st;
                    if (!__any_sync(0xFFFFFFFFu, redo)) break;
                    if (redo) {
                        st = ns;
                        fsm_slow_sync_lane(s_h, a.d_bits, sub_bit0, lim, st, cnt, exit_state, first_end);
                    }
                }
                // lanes behind the end pass the last active lane's exit on
                const unsigned act = __ballot_sync(0xFFFFFFFFu, active);
                const int last = act ? 31 - __clz(act) : 0;
                const uint32_t ex = __shfl_sync(0xFFFFFFFFu, exit_state, last);
                const uint32_t own_exit = exit_state;
                if (!active) exit_state = ex;
                start = st & kFsmToken ? (st & 0xFFu) : 0u;   // (a bit-offset start: F3 gets the offset from its own arguments)
                if (COMPAT) {   // the first code that STARTS in the lane's part, the codes that start there
                    const uint32_t more = own_exit != 0u ? 1u : 0u;   // a code begins here and ends further on
                    if (!active) { start = 0u; }
                    else if (!(st & kFsmToken)) { start = st; cnt += more; }                  // the stream's first code, at its given bit
                    else if ((st & 0x1FFu) == 0u) { start = 0u; cnt += more; }
                    else if (cnt == 0u) { start = 0u; }                                        // all of it inside one code: nothing starts here
                    else { start = first_end; cnt = cnt - 1u + more; }
                }
            }
            carry = __shfl_sync(0xFFFFFFFFu, exit_state, 31) | kFsmToken;
            if (tt < (uint32_t)warm) {
                assumed = carry;
            } else {
                if ((unsigned long long)tt * 32 + lane < sub_left) *info = (uint16_t)(COMPAT ? (start | (cnt << 7)) : (start | (cnt << 8)));   // COMPAT: the window kernels' form
                total += cnt;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
        if (lane == 0) {
            ws.seg_cnt[seg] = total;
            ws.seg_assumed[seg] = assumed;
            ws.seg_exit[seg] = carry;
        }
        if constexpr (TMA) tma_phase = cur.phase;   // every copy that was issued has been waited for (take(.., more) fetches only what is walked)
    }
}

// ------------------------------------------------------------------------------------------ F3 (FSM)

struct FsmWriteArgs {
    const uint8_t *d_bits;
    unsigned long long end, nsub, ntiles, nseg;
    uint8_t *out;
    unsigned long long n_out;
    int lead;
    uint32_t stage_bytes;      // per warp
    uint32_t rows;             // rows of the F3 table (states + DEAD), all resident in shared memory
    int32_t *d_status;
};

__device__ __forceinline__ void sts_u32_if(uint32_t addr, uint32_t v, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
}

struct FsmWriteTabs {
    uint32_t tab, xtab;            // shared-memory addresses: the rows, and the suffix rows moved back by 24 rows (count 3, k 0 = row 24)
};

// One byte step of the write walk.  acc = the lane's last four symbols (newest in the top byte), G = 8 x pending bytes in its
// low 5 bits (bit 5 toggles when a word completes), wptr = shared address of the word being filled, e = the previous entry.
// XTAB = false: a byte holds at most two digits (4-bit digits: n = 16 and the nibble-per-digit radices), so no byte completes a third
// code and the suffix look-up is not compiled in (its predicated-off instructions alone cost 7 % of the kernel at n = 16).
template <int J, bool XTAB>
__device__ __forceinline__ void fsm_write_step(const FsmWriteTabs &T, uint32_t w, uint32_t &e, uint32_t &acc, uint32_t &G, uint32_t &wptr) {
    const uint32_t idx = prmt(w, e, 0xFF60u | (uint32_t)J);   // byte J | state << 8; bytes 2, 3 = sign of the count byte = 0
    e = lds_u32_at(T.tab, idx);
    const uint32_t c = e >> 27;
    uint32_t syms = e;   // symbols 0 and 1 (the state and the count above them are never taken: a window slides by c bytes)
    if (XTAB && c >= 3u) {       // symbols 2 and 3: what the byte decodes to from digit k on, row = count * 8 + k
        const uint32_t xi = prmt(w, e, 0xFF70u | (uint32_t)J);
        uint32_t x;
        asm volatile("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %1, 2, %2;\n\tld.shared.u16 %0, [a];\n\t}" : "=r"(x) : "r"(xi), "r"(T.xtab));
        syms = prmt(e, x, 0x5410u);
    }
    const uint32_t sw = __funnelshift_l(acc, syms, G);   // the pending bytes, then this step's symbols
    uint32_t sel;
    asm("mad.lo.u32 %0, %1, 0x1111, 0x3210;" : "=r"(sel) : "r"(c));
    acc = prmt(acc, syms, sel);                           // slide the window by c bytes
    // G += 8 c; bit 5 of G toggles exactly when a word has been completed: store it and advance
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .u32 g2, x;\n\t"
        "mad.lo.u32 g2, %3, 8, %1;\n\t"
        "xor.b32 x, g2, %1;\n\t"
        "and.b32 x, x, 32;\n\t"
        "setp.ne.u32 q, x, 0;\n\t"
        "@q st.shared.u32 [%0], %2;\n\t"
        "@q add.u32 %0, %0, 4;\n\t"
        "mov.u32 %1, g2;\n\t}"
        : "+r"(wptr), "+r"(G)
        : "r"(sw), "r"(c)
        : "memory");
}

template <bool XTAB>
__device__ __forceinline__ void fsm_write_walk(const FsmWriteTabs &T, const uint32_t (&w)[8], uint32_t &e, uint32_t &acc, uint32_t &G, uint32_t &wptr) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        fsm_write_step<0, XTAB>(T, w[k], e, acc, G, wptr);
        fsm_write_step<1, XTAB>(T, w[k], e, acc, G, wptr);
        fsm_write_step<2, XTAB>(T, w[k], e, acc, G, wptr);
        fsm_write_step<3, XTAB>(T, w[k], e, acc, G, wptr);
    }
}

template <bool XTAB>
__global__ void __launch_bounds__(768, 1) fsm_write_kernel(FsmWriteArgs a, FsmTables t, FastWorkspace ws) {
    extern __shared__ __align__(16) uint8_t fsm_smem[];
    FsmHeader *s_h = (FsmHeader *)fsm_smem;
    uint32_t *s_tab = (uint32_t *)(fsm_smem + kFsmHeaderBytes);
    uint16_t *s_xtab = (uint16_t *)(fsm_smem + kFsmHeaderBytes + (size_t)a.rows * kFsmWriteRowBytes);
    if (*ws.mismatch) return;  // the robust path redoes the stream
    {
        const uint32_t *src = (const uint32_t *)t.hdr;
        uint32_t *dst = (uint32_t *)s_h;
        for (int i = threadIdx.x; i < (int)(sizeof(FsmHeader) / 4); i += blockDim.x) dst[i] = src[i];
        const uint4 *s4 = (const uint4 *)t.write, *x4 = (const uint4 *)t.writex;
        uint4 *d4 = (uint4 *)s_tab, *dx4 = (uint4 *)s_xtab;
        for (int i = threadIdx.x; i < (int)a.rows * 64; i += blockDim.x) d4[i] = s4[i];            // 1 KB per row
        for (int i = threadIdx.x; i < kFsmSuffixRows * 32; i += blockDim.x) dx4[i] = x4[i];        // 512 bytes per suffix row
        __syncthreads();
    }
    FsmWriteTabs T;
    T.tab = (uint32_t)__cvta_generic_to_shared(s_tab);
    T.xtab = (uint32_t)__cvta_generic_to_shared(s_xtab) - 24u * (uint32_t)kFsmWriteXRowBytes;
    asm volatile("" : "+r"(T.tab), "+r"(T.xtab));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    uint8_t *stage = fsm_smem + kFsmHeaderBytes + (size_t)a.rows * kFsmWriteRowBytes + kFsmWriteXTableBytes + (size_t)warp * a.stage_bytes;
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(stage);
    const unsigned long long nvec = ((a.end + 7) / 8 + 15) / 16;
    const unsigned long long tile_bits = 32ull * kF_SubBits;
    const uint32_t dead_state = (uint32_t)s_h->nstates;
    const uint32_t start_token = *ws.start_slot();   // written by F1 (a bit offset only for a stream that starts inside a byte)
    bool corrupt = false;
    for (unsigned long long seg = (unsigned long long)blockIdx.x * warps + warp; seg < a.nseg; seg += (unsigned long long)gridDim.x * warps) {
        const unsigned long long tile0 = (unsigned long long)a.lead + seg * kF_SegTiles;
        FsmCursor cur;
        cur.init(a.d_bits, tile0, nvec, lane);
        const uint32_t ntile = (uint32_t)min((unsigned long long)kF_SegTiles, a.ntiles - tile0);
        const uint16_t *info = ws.sub_info + tile0 * 32 + lane;
        const unsigned long long sub_left = a.nsub - tile0 * 32;
        unsigned long long ob = ws.seg_off[seg];
        uint32_t next_info = lane < sub_left ? *info : 0u;
        for (uint32_t tt = 0; tt < ntile; tt++) {
            uint32_t w[8];
            cur.take(w, tt + 1 < ntile);
            const uint32_t my_info = next_info;
            info += 32;
            next_info = (tt + 1 < ntile && (unsigned long long)(tt + 1) * 32 + lane < sub_left) ? *info : 0u;
            const uint32_t my_cnt = my_info >> 8;
            uint32_t incl = my_cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += x;
            }
            const uint32_t tile_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const uint32_t excl = incl - my_cnt;
            const unsigned long long tbit0 = (tile0 + tt) * tile_bits;
            const bool first_lane_offset = seg == 0 && a.lead == 0 && tt == 0 && !(start_token & kFsmToken);
            const bool full = tbit0 + tile_bits <= a.end && (unsigned long long)(tt + 1) * 32 <= sub_left && !first_lane_offset;   // warp-uniform
            if (!full) {
                // slow tile: digit by digit, bytes straight to the output
                const unsigned long long sub_bit0 = tbit0 + (unsigned long long)lane * kF_SubBits;
                if (sub_bit0 < a.end && (unsigned long long)tt * 32 + lane < sub_left) {
                    const uint32_t lim = (uint32_t)min((unsigned long long)kF_SubBits, a.end - sub_bit0);
                    int d = 0;
                    uint32_t v = 0;
                    unsigned long long p = sub_bit0;
                    if (first_lane_offset && lane == 0) p += start_token;
                    else if ((my_info & 0xFFu) > dead_state) p += ((my_info & 0xFFu) - dead_state) * (uint32_t)s_h->bpd;   // an entry row
                    else fsm_state_node(s_h, my_info & 0xFFu, d, v);
                    bool dead = false;
                    uint8_t *dst = a.out + ob + excl;
                    const unsigned long long room = a.n_out > ob + excl ? a.n_out - (ob + excl) : 0ull;
                    const uint32_t c = fsm_walk_digits(s_h, a.d_bits, p, sub_bit0 + lim, d, v, dead, [&](uint32_t sym, uint32_t i) {
                        if (i < room && i < my_cnt) dst[i] = (uint8_t)sym;
                    });
                    // a code that is cut off by the end of the STREAM (not by the end of the lane's subsequence)
                    const bool cut = sub_bit0 + lim == a.end && (d != 0 || ((a.end - p) % (unsigned)s_h->bpd) != 0);
                    if (dead || c != my_cnt || cut) corrupt = true;
                }
                ob += tile_total;
                continue;
            }
            // fast tile, in one pass or (more symbols than the staging tile holds) lanes 0..15, then 16..31
            const uint32_t half_total = __shfl_sync(0xFFFFFFFFu, incl, 15);
            const uint32_t a0 = (uint32_t)(((uintptr_t)a.out + ob) & 15);
            const bool one_pass = a0 + tile_total + 8 <= a.stage_bytes;
            const int npass = one_pass ? 1 : 2;
            for (int pass = 0; pass < npass; pass++) {
                const bool mine = one_pass || (lane >> 4) == pass;
                const uint32_t pass_excl0 = (!one_pass && pass == 1) ? half_total : 0u;
                const uint32_t pass_total = one_pass ? tile_total : (pass == 0 ? half_total : tile_total - half_total);
                const unsigned long long pob = ob + pass_excl0;
                const uint32_t al = (uint32_t)(((uintptr_t)a.out + pob) & 15);
                const uint32_t pos0 = al + (excl - pass_excl0);
                if (al + pass_total + 8u > a.stage_bytes) {   // counts that no walk of F1 can have produced (a damaged index): do not stage them
                    corrupt = true;
                    continue;
                }
                __syncwarp();  // the previous copy-out has read the staging tile
                if (lane == 0) *(uint32_t *)(stage + ((al + pass_total) & ~3u)) = 0u;   // the one boundary word nobody's first store initialises
                __syncwarp();
                uint32_t wptr0 = stage_addr + (pos0 & ~3u), wptr = wptr0, G = 8u * (pos0 & 3u), acc = 0u;
                uint32_t meta = (my_info & 0xFFu) << 16;
                if (mine && my_cnt) fsm_write_walk<XTAB>(T, w, meta, acc, G, wptr);
                __syncwarp();
                if (mine && my_cnt) {
                    const uint32_t pend = (G >> 3) & 3u;
                    if (pend) atomicOr((uint32_t *)(stage + (wptr - stage_addr)), __funnelshift_l(acc, 0u, G));
                    const uint32_t written = (wptr - wptr0) + pend - (pos0 & 3u);
                    if (written != my_cnt || ((meta >> 16) & 0xFFu) == dead_state) corrupt = true;
                }
                __syncwarp();
                // copy-out: staging byte i <-> out[pob - al + i]; 16-byte words are aligned on both sides
                const uint32_t span = al + pass_total;
                if (pob + pass_total <= a.n_out) {
                    const uint32_t jfull = span >> 4;
                    const uint32_t head = al ? 1u : 0u;
                    for (uint32_t j = head + lane; j < jfull; j += 32)
                        stg_stream((uint4 *)(a.out + (pob - al)) + j, *(const uint4 *)(stage + j * 16));
                    const uint32_t k = lane < 16 ? (uint32_t)lane : jfull * 16 + (lane - 16);
                    const bool edge = lane < 16 ? (al != 0 || jfull == 0) : (jfull != 0);
                    if (edge && k >= al && k < span) a.out[pob - al + k] = stage[k];
                } else {  // more symbols than the caller expects (corrupt stream): clamp every byte
                    for (uint32_t k = al + lane; k < span; k += 32)
                        if (pob - al + k < a.n_out) a.out[pob - al + k] = stage[k];
                }
            }
            ob += tile_total;
        }
    }
    if (corrupt) set_status(a.d_status, DC_ERR_CORRUPT);
}

// ------------------------------------------------------------------------------------------ last resort: one thread, whole stream
// For the nibble-per-digit radices (5 .. 15), whose tables have no window LUTs: a stream that does not self-synchronise
// (codes that all have the same odd number of digits, say) cannot go to the window kernels' robust path.  One thread walks it
// digit by digit from its first bit.  Slow (tens of MB/s) and exact; such streams are degenerate.
__global__ void fsm_serial_kernel(const uint8_t *__restrict__ d_bits, unsigned long long bit_start, unsigned long long end, FsmTables t,
                                  uint8_t *__restrict__ out, unsigned long long n_out, int32_t *__restrict__ d_status) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const FsmHeader *h = t.hdr;
    if (h->nstates == 0) { set_status(d_status, DC_ERR_ARG); return; }
    const int bpd = h->bpd;
    int d = 0;
    uint32_t v = 0;
    unsigned long long p = bit_start, o = 0;
    bool bad = false;
    while (p + bpd <= end) {
        const int r = fsm_digit(h, d, v, stream_digit(d_bits, p, bpd));
        p += bpd;
        if (r >= 0) {
            if (o < n_out) out[o] = (uint8_t)r;
            o++;
        } else if (r == -2) {
            bad = true;
            break;
        }
    }
    if (bad || d != 0 || p != end || o != n_out) set_status(d_status, o > n_out && !bad ? DC_ERR_CAPACITY : DC_ERR_CORRUPT);
}

}  // namespace dc
