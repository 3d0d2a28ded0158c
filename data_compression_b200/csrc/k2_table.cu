// k2_table.cu -- K2: n-ary code lengths and canonical code table in ONE single-CTA kernel.
//
// Replaces huffman() (n_ary_huffman.c:1161-1208 = setup_nodes :773, generate_huffman_tree :868,
// summarize_tree_with_lengths :1033) and convert_lengths_to_encode_table() (:1382-1612), as the reference
// behaves when compiled with -DNDEBUG:
//   * nz  = symbols with non-zero count                                   (:880-886)
//   * d   = (n-1) - ((nz-1) % (n-1)) dummy leaves of count 1, indexed after the real symbols (:900-929)
//   * the reference's repeated stable bubble sort (:672-731) orders nodes by (count, node index) and a
//     freshly merged node is appended on the right, so it follows every node of equal count (:962-1002).
//     That is exactly: rank-sort the leaves once by (count, index), then merge from two queues -- sorted
//     leaves and a FIFO of internal nodes -- where a leaf wins a tie.
//   * length = number of parent hops to the root                          (:1069-1076)
//   * canonical values: for len = min..max, symbols ascending, value = code++, then code *= n (:1540-1568);
//     min/max are taken over i < max_symbol_value (the last slot is skipped, :1336/:1360).
//
// O(alphabet^2) work, ~10 us: off the time-critical path but it fixes every output bit, so it is
// written for exactness, not throughput.  Counts are 64-bit.
#include "dc_common.cuh"

namespace dc {

constexpr int kTabThreads = 1024;
constexpr int kTabCap = 1040;  // >= DC_MAX_LEAVES + max dummy leaves (n_ary <= 512)

// bits a digit occupies in the stream the encode / decode kernels work on.  n = 3 (the reference's default radix) is
// handled as 2 bits per trit -- an intermediate "T2" stream that dc_trit_pack() turns into the 5-trits-per-byte payload
// (tab->packed_radix = 3 marks such a table).  The other radices up to 16 (5, 6, 7, 9, 10, ... -- SURVEY N4 names 9 and 10) take
// one NIBBLE per digit, most significant digit first, high nibble first: binary-coded digits, the layout n = 16 has anyway, and
// the stream is the payload.  Their tables carry no window LUTs (those index by the numeric value of a bit window, which a
// nibble-per-digit code is not): the decoder walks them with the byte-stepped state machine only (k4_fsm.cuh), which is
// radix-generic.  Radices above 16 have tables only.
__device__ __forceinline__ int bits_per_digit_of(int n) { return n == 2 ? 1 : n == 4 ? 2 : n == 3 ? 2 : (n >= 5 && n <= 16) ? 4 : 0; }
// a base-n numeral of `len` digits rewritten with one nibble per digit (most significant first)
__device__ __forceinline__ unsigned int digits_to_nibbles(unsigned int v, int len, unsigned int n) {
    unsigned int r = 0;
    for (int k = 0; k < len; k++) { r |= (v % n) << (4 * k); v /= n; }
    return r;
}
// a base-3 numeral of `len` digits rewritten with one 2-bit field per digit (most significant first)
__device__ __forceinline__ unsigned int trits_to_t2(unsigned int v, int len) {
    unsigned int r = 0;
    for (int k = 0; k < len; k++) { r |= (v % 3u) << (2 * k); v /= 3u; }
    return r;
}
// the first `len` 2-bit digits of a left-aligned window as a base-3 value; false if one of them is 3 (no trit)
__device__ __forceinline__ bool t2_to_trits(unsigned int window, int window_bits, int len, unsigned int *v) {
    unsigned int r = 0;
    for (int k = 0; k < len; k++) {
        const unsigned int d = (window >> (window_bits - 2 * (k + 1))) & 3u;
        if (d == 3u) return false;
        r = r * 3u + d;
    }
    *v = r;
    return true;
}

__global__ void __launch_bounds__(kTabThreads) table_kernel(const unsigned long long *__restrict__ d_hist,
                                                            const int32_t *__restrict__ d_lengths_in, int nsym,
                                                            int n_ary, dc_huff_table *__restrict__ tab, TableRaw raw,
                                                            int32_t *__restrict__ host_meta, int nhist, int hist_stride) {
    __shared__ unsigned long long s_cnt[kTabCap];     // compacted leaf counts (index order)
    __shared__ unsigned long long s_scnt[kTabCap];    // leaf counts sorted by (count, index)
    __shared__ unsigned long long s_icount[kTabCap];  // internal node counts, creation order
    __shared__ int s_idx[kTabCap];                    // compacted leaf -> symbol (>= nsym: dummy)
    __shared__ int s_ssym[kTabCap];                   // sorted leaf -> symbol
    __shared__ int s_lparent[kTabCap];                // sorted leaf -> internal node
    __shared__ int s_iparent[kTabCap];                // internal node -> internal node
    __shared__ int s_len[kTabCap];                    // symbol -> length in digits
    __shared__ unsigned int s_lencount[64], s_first[64], s_off[64];
    __shared__ int s_warp[32];
    uint16_t *s_lut = (uint16_t *)s_cnt;  // the single-symbol LUT, reusing the leaf-count array (8 KB) once the tree is done
    __shared__ int s_nint, s_minlen, s_maxlen, s_status;
    __shared__ unsigned long long s_totsym, s_totbits;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) { s_lencount[tid] = 0; s_first[tid] = 0; s_off[tid] = 0; }
    if (tid == 0) { s_nint = 0; s_minlen = 300; s_maxlen = 0; s_status = DC_OK; s_totsym = 0; s_totbits = 0; }
    for (int i = tid; i < kTabCap; i += kTabThreads) s_len[i] = 0;

    unsigned long long my_count = 0;
    int nz = 0, dummies = 0;
    if (d_hist) {
        // ---- 1. compact the non-zero leaves in index order, append the dummy leaves
        if (tid < nsym)
            for (int h = 0; h < nhist; h++) my_count += d_hist[(size_t)h * hist_stride + tid];
        const bool used = my_count != 0;
        const unsigned ball = __ballot_sync(0xFFFFFFFFu, used);
        if (lane == 0) s_warp[warp] = __popc(ball);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < 32; w++) {
            const int c = s_warp[w];
            if (w < warp) before += c;
            nz += c;
        }
        const int pos = before + __popc(ball & ((1u << lane) - 1u));
        const int k = n_ary - 1;
        dummies = k - ((nz - 1) % k);  // as written, :900-903 (C remainder semantics)
        const int nleaf = nz + dummies;
        if (used) { s_cnt[pos] = my_count; s_idx[pos] = tid; }
        for (int j = tid; j < dummies; j += kTabThreads) { s_cnt[nz + j] = 1ull; s_idx[nz + j] = nsym + j; }
        __syncthreads();

        // ---- 2. rank sort by (count, index): position order == index order, dummies last
        for (int p = tid; p < nleaf; p += kTabThreads) {
            const unsigned long long mine = s_cnt[p];
            int rank = 0;
            for (int q = 0; q < nleaf; q++) {
                const unsigned long long cq = s_cnt[q];
                rank += (cq < mine) || (cq == mine && q < p);
            }
            s_scnt[rank] = mine;
            s_ssym[rank] = s_idx[p];
        }
        __syncthreads();

        // ---- 3. two-queue n-way merge (serial: every step depends on the previous sum).  This loop is most of the kernel's
        // time, and most of its time was the shared-memory load of the next queue head sitting in the dependent chain
        // (compare -> pop -> load -> compare): the next four leaf counts and the next three internal counts are kept in
        // registers, so a pop is a few moves and the load it issues is not needed before three pops later.
        if (tid == 0) {
            int lh = 0, ih = 0, it = 0, remaining = nleaf;
            auto leaf_at = [&](int i) { return i < nleaf ? s_scnt[i] : ~0ull; };
            unsigned long long l0 = leaf_at(0), l1 = leaf_at(1), l2 = leaf_at(2), l3 = leaf_at(3);
            unsigned long long i0 = 0ull, i1 = 0ull, i2 = 0ull;   // s_icount[ih], [ih + 1], [ih + 2] where those exist (< it)
            while (remaining > 1) {
                unsigned long long sum = 0;
                for (int c = 0; c < n_ary; c++) {
                    const bool leaf_ok = lh < nleaf, int_ok = ih < it;
                    if (leaf_ok && (!int_ok || l0 <= i0)) {  // leaf wins ties
                        sum += l0;
                        s_lparent[lh] = it;
                        lh++;
                        l0 = l1; l1 = l2; l2 = l3;
                        l3 = leaf_at(lh + 3);
                    } else {
                        sum += i0;
                        s_iparent[ih] = it;
                        ih++;
                        i0 = i1; i1 = i2;
                        i2 = ih + 2 < it ? s_icount[ih + 2] : 0ull;
                    }
                }
                s_icount[it] = sum;
                const int d = it - ih;   // the new node joins the internal FIFO d places behind its head
                if (d == 0) i0 = sum; else if (d == 1) i1 = sum; else if (d == 2) i2 = sum;
                it++;
                remaining -= k;
            }
            s_nint = it;
        }
        __syncthreads();

        // ---- 4. depth of every real leaf = hops to the root (the last internal node)
        const int root = s_nint - 1;
        if (root >= 0) {
            for (int r = tid; r < nleaf; r += kTabThreads) {
                int node = s_lparent[r], depth = 1;
                while (node != root) { node = s_iparent[node]; depth++; }
                const int sym = s_ssym[r];
                if (sym < nsym) s_len[sym] = depth;
            }
        }
        __syncthreads();
    } else {
        if (tid < nsym) s_len[tid] = d_lengths_in[tid];
        __syncthreads();
    }

    // ---- 5. canonical code values
    const int my_len = tid < nsym ? s_len[tid] : 0;
    const int max_symbol_value = nsym - 1;
    if (tid < max_symbol_value) {  // the reference's scans skip the last slot (:1336, :1360)
        if (my_len > 0) { atomicMax(&s_maxlen, my_len); atomicMin(&s_minlen, my_len); }
    }
    if (tid < nsym && my_len > 0 && my_len < 64) atomicAdd(&s_lencount[my_len], 1u);
    if (tid < nsym && (my_len < 0 || my_len >= 64)) s_status = DC_ERR_CODE_TOO_LONG;
    __syncthreads();
    const int min_len = s_minlen, max_len = s_maxlen;
    if (tid == 0) {
        if (max_len >= 16) s_status = DC_ERR_CODE_TOO_LONG;  // assert :1414
        unsigned long long code = 0;
        unsigned int off = 0;
        for (int cl = min_len; cl <= max_len && cl < 64; cl++) {
            s_first[cl] = (unsigned int)code;
            s_off[cl] = off;
            const unsigned int c = s_lencount[cl];
            if (c && code + c - 1 > 0x7FFFFFFFull) s_status = DC_ERR_CODE_TOO_LONG;  // int current_code :1540
            code += c;
            off += c;
            code *= (unsigned long long)n_ary;
            if (code > (1ull << 40)) { s_status = DC_ERR_CODE_TOO_LONG; code &= (1ull << 40) - 1; }
        }
    }
    __syncthreads();
    unsigned int my_value = 0;
    int my_rank = 0, assigned = 0;
    if (tid < nsym && my_len > 0 && my_len >= min_len && my_len <= max_len && my_len < 64) {
        for (int j = 0; j < tid; j++) my_rank += (s_len[j] == my_len);
        my_value = s_first[my_len] + (unsigned int)my_rank;
        assigned = 1;
    }
    const int status = s_status;

    if (raw.lengths && tid < nsym) raw.lengths[tid] = my_len;
    if (raw.values && tid < nsym) raw.values[tid] = my_value;
    if (raw.assigned && tid < nsym) raw.assigned[tid] = assigned;
    if (raw.status && tid == 0) *raw.status = status;

    // ---- 6. device table for the encode / decode kernels (byte alphabet in 259 slots)
    if (!tab) return;
    const int bpd = bits_per_digit_of(n_ary);
    const int nbits = my_len * bpd;
    const bool t2 = n_ary == 3;
    // what the kernels emit / match for this symbol: the canonical value itself, or its trits as 2-bit fields
    const bool nib = nibble_radix(n_ary);
    const unsigned int my_stream_value = (t2 && assigned && my_len <= 16) ? trits_to_t2(my_value, my_len)
                                         : (nib && assigned && my_len <= 8) ? digits_to_nibbles(my_value, my_len, (unsigned int)n_ary) : my_value;
    if (tid <= DC_NSLOTS) {
        tab->lengths[tid] = tid < nsym ? my_len : 0;
        tab->values[tid] = tid < nsym ? my_value : 0u;
        tab->sorted[tid] = 0;
    }
    if (tid < 256) {
        tab->enc[tid] = (assigned && nbits <= 26) ? ((my_stream_value << 6) | (unsigned int)nbits) : 0u;
        tab->enc64[tid] = assigned ? ((unsigned long long)my_stream_value | ((unsigned long long)nbits << 32)) : 0ull;
    }
    if (tid < 32) {
        tab->first_code[tid] = s_first[tid];
        tab->len_count[tid] = s_lencount[tid];
        tab->len_offset[tid] = s_off[tid];
    }
    if (d_hist && my_count) {
        atomicAdd(&s_totsym, my_count);
        atomicAdd(&s_totbits, my_count * (unsigned long long)nbits);
    }
    __syncthreads();
    if (assigned && tid <= DC_NSLOTS) tab->sorted[s_off[my_len] + my_rank] = (uint16_t)tid;
    // reuse s_ssym as the canonical order for the LUT fill
    if (assigned) s_ssym[s_off[my_len] + my_rank] = tid;
    __syncthreads();
    for (int e = tid; e < (1 << DC_LUT_BITS); e += kTabThreads) {
        unsigned int entry = 0;
        if (bpd && !nib) {
            for (int l = min_len; l <= max_len && l < 32; l++) {
                const int lb = l * bpd;
                if (lb > DC_LUT_BITS) break;
                unsigned int v = (unsigned int)e >> (DC_LUT_BITS - lb);
                if (t2 && !t2_to_trits((unsigned int)e, DC_LUT_BITS, l, &v)) continue;  // a 2-bit field of 3 is no trit
                if (s_lencount[l] && v >= s_first[l] && v - s_first[l] < s_lencount[l]) {
                    entry = ((unsigned int)lb << 8) | (unsigned int)(s_ssym[s_off[l] + (v - s_first[l])] & 0xFF);
                    break;
                }
            }
        }
        tab->lut[e] = (uint16_t)entry;
        s_lut[e] = (uint16_t)entry;
    }
    __syncthreads();
    // multi-symbol tables: greedily take every code that lies completely inside the 12 index bits
    for (int e = tid; e < (1 << DC_LUT_BITS) && !t2; e += kTabThreads) {
        unsigned int used = 0, count = 0, first = 0, sym0 = 0, sym1 = 0, used2 = 0;
        while (used < DC_LUT_BITS) {
            const unsigned int one = s_lut[((unsigned int)e << used) & ((1u << DC_LUT_BITS) - 1u)];
            const unsigned int nb = one >> 8;
            if (nb == 0 || used + nb > DC_LUT_BITS) break;
            if (count == 0) { first = nb; sym0 = one & 0xFFu; }
            if (count == 1) sym1 = one & 0xFFu;
            used += nb;
            count++;
            if (count <= 2) used2 = used;
        }
        // an unused code slot (the dummy leaves): when every code fits the index it gets a real entry -- advance one
        // digit, no symbol, flagged -- so the decoders need no escape branch; with longer codes 0 = escape
        const bool total_lut = bpd != 0 && max_len * bpd <= DC_LUT_BITS;
        const unsigned int dead = (unsigned int)bpd;
        tab->lut_count[e] = count ? (used | (count << 16) | (first << 24)) : (total_lut ? (dead | (dead << 24)) : (DC_LUT_COUNT_MARK | DC_LUT_NO_SUBTABLE));
        tab->lut_pair[e] = count ? (sym0 | (sym1 << 8) | (used2 << 16) | (first << 24) | ((count < 2 ? count : 2u) << 30))
                                 : (total_lut ? ((dead << 16) | (dead << 24) | (1u << 29)) : (DC_LUT_PAIR_MARK | DC_LUT_NO_SUBTABLE));
    }
    // radix 3: the same two tables indexed by the base-3 value of the next DC_TRIT_WINDOW trits (the kernels compute it from
    // the 2-bit fields), so a look-up sees 8 trits instead of the 6 a 12-bit index would hold; bit counts are T2 bits
    for (int e = tid; e < 6561 && t2; e += kTabThreads) {
        unsigned int d[DC_TRIT_WINDOW];
        {
            unsigned int v = (unsigned int)e;
            for (int k = DC_TRIT_WINDOW - 1; k >= 0; k--) { d[k] = v % 3u; v /= 3u; }
        }
        unsigned int used = 0, count = 0, first = 0, sym0 = 0, sym1 = 0, used2 = 0;   // used: trits
        while (used < DC_TRIT_WINDOW) {
            unsigned int val = 0, sym = 0;
            int found = 0;
            for (int l = 1; used + l <= DC_TRIT_WINDOW && l <= max_len; l++) {
                val = val * 3u + d[used + l - 1];
                if (l >= min_len && s_lencount[l] && val >= s_first[l] && val - s_first[l] < s_lencount[l]) {
                    sym = (unsigned int)(s_ssym[s_off[l] + (val - s_first[l])] & 0xFF);
                    found = l;
                    break;
                }
            }
            if (!found) break;
            if (count == 0) { first = 2u * found; sym0 = sym; }
            if (count == 1) sym1 = sym;
            used += found;
            count++;
            if (count <= 2) used2 = 2u * used;
        }
        const bool total_lut = max_len <= DC_TRIT_WINDOW;
        const unsigned int dead = 2u, ubits = 2u * used;
        tab->lut_count[e] = count ? (ubits | (count << 16) | (first << 24)) : (total_lut ? (dead | (dead << 24)) : (DC_LUT_COUNT_MARK | DC_LUT_NO_SUBTABLE));
        tab->lut_pair[e] = count ? (sym0 | (sym1 << 8) | (used2 << 16) | (first << 24) | ((count < 2 ? count : 2u) << 30))
                                 : (total_lut ? ((dead << 16) | (dead << 24) | (1u << 29)) : (DC_LUT_PAIR_MARK | DC_LUT_NO_SUBTABLE));
    }
    // ---- 7. second level for codes of 13..16 bits (tables whose longest code exceeds the 12 index bits)
    const bool need2 = bpd != 0 && !t2 && !nib && max_len * bpd > DC_LUT_BITS;   // radix 3 has its own, wider index instead (block-uniform)
    if (!need2) {
        if (tid == 0) tab->lut2_used = 0;
    } else {
        __syncthreads();
        unsigned char *s_flag = (unsigned char *)s_scnt;     // [4096] 1 = this window is the prefix of 13..16-bit codes
        unsigned short *s_sub = (unsigned short *)s_icount;  // [4096] its subtable
        // the code (13..16 bits) that a left-aligned 16-bit window starts with; 0 if none
        auto long_code = [&](unsigned int w16) -> unsigned int {
            for (int l = min_len; l <= max_len && l < 32; l++) {
                const int lb = l * bpd;
                if (lb <= DC_LUT_BITS) continue;
                if (lb > 16) break;
                unsigned int v = w16 >> (16 - lb);
                if (t2 && !t2_to_trits(w16, 16, l, &v)) continue;
                if (s_lencount[l] && v >= s_first[l] && v - s_first[l] < s_lencount[l])
                    return ((unsigned int)lb << 8) | (unsigned int)(s_ssym[s_off[l] + (v - s_first[l])] & 0xFF);
            }
            return 0u;
        };
        for (int e = tid; e < (1 << DC_LUT_BITS); e += kTabThreads) {
            unsigned int any = 0;
            if (need2 && s_lut[e] == 0)
                for (unsigned int sfx = 0; sfx < 16 && !any; sfx++) any = long_code(((unsigned int)e << 4) | sfx);
            s_flag[e] = any ? 1 : 0;
        }
        __syncthreads();
        {   // number the flagged prefixes in index order: four windows per thread, a block-wide exclusive scan
            static_assert((1 << DC_LUT_BITS) == 4 * kTabThreads, "four LUT windows per thread");
            const int lane = tid & 31, warp = tid >> 5;
            int mine = 0;
            for (int j = 0; j < 4; j++) mine += s_flag[4 * tid + j];
            int incl = mine;
            for (int d = 1; d < 32; d <<= 1) {
                const int x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += x;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            int id = incl - mine, total = 0;
            for (int w = 0; w < kTabThreads / 32; w++) {
                if (w < warp) id += s_warp[w];
                total += s_warp[w];
            }
            for (int j = 0; j < 4; j++) {
                const int e = 4 * tid + j;
                s_sub[e] = 0xFFFF;
                if (s_flag[e]) {
                    if (id < DC_LUT2_SUBTABLES) s_sub[e] = (unsigned short)id;
                    id++;
                }
            }
            if (tid == 0) {
                tab->lut2_used = total < DC_LUT2_SUBTABLES ? total : DC_LUT2_SUBTABLES;
            }
        }
        __syncthreads();
        for (int i = tid; i < (DC_LUT2_SUBTABLES + 1) * 16; i += kTabThreads) tab->lut2[i] = 0;
        __syncthreads();
        for (int e = tid; e < (1 << DC_LUT_BITS); e += kTabThreads) {
            const unsigned int id = s_sub[e];
            if (id == 0xFFFFu) continue;
            for (unsigned int sfx = 0; sfx < 16; sfx++) tab->lut2[id * 16 + sfx] = (uint16_t)long_code(((unsigned int)e << 4) | sfx);
            tab->lut_count[e] = DC_LUT_COUNT_MARK | id;
            tab->lut_pair[e] = DC_LUT_PAIR_MARK | id;
        }
    }
    // ---- 8. 14-bit count table for the decoder's synchronisation pass (tables whose longest code has 13 or 14 bits)
    if (bpd != 0 && !t2 && !nib && max_len * bpd > DC_LUT_BITS && max_len * bpd <= DC_LUT14_BITS) {
        // the code a left-aligned 14-bit window starts with, among those of at most `avail` bits: bits << 8 | 1, or 0
        auto code_bits = [&](unsigned int w14, int avail) -> unsigned int {
            for (int l = min_len; l <= max_len; l++) {
                const int lb = l * bpd;
                if (lb > avail) break;
                const unsigned int v = w14 >> (DC_LUT14_BITS - lb);
                if (s_lencount[l] && v >= s_first[l] && v - s_first[l] < s_lencount[l]) return (unsigned int)lb;
            }
            return 0u;
        };
        for (int e = tid; e < (1 << DC_LUT14_BITS); e += kTabThreads) {
            unsigned int used = 0, count = 0, first = 0;
            while (used < DC_LUT14_BITS) {
                const unsigned int nb = code_bits(((unsigned int)e << used) & ((1u << DC_LUT14_BITS) - 1u), DC_LUT14_BITS - (int)used);
                if (nb == 0) break;
                if (count == 0) first = nb;
                used += nb;
                count++;
            }
            // every code fits the index, so a window without a code starts with an unused slot: one digit, no symbol
            tab->lut14[e] = (uint16_t)(count ? (used | (count << 8) | (first << 12)) : ((unsigned int)bpd | ((unsigned int)bpd << 12)));
        }
    }
    if (tid == 0) {
        tab->n_ary = n_ary;
        tab->bits_per_digit = bpd;
        tab->max_symbol_value = max_symbol_value;
        tab->nonzero_symbols = nz;
        tab->dummy_nodes = dummies;
        tab->min_len = min_len;
        tab->max_len = max_len;
        tab->max_bits = max_len * bpd;
        tab->status = (status == DC_OK && max_len * bpd > 32) ? DC_ERR_CODE_TOO_LONG : status;
        tab->packed_radix = t2 ? 3 : 0;
        tab->total_symbols = s_totsym;
        tab->total_bits = s_totbits;
        // the byte-stepped decoder (k4_fsm.cuh) applies if the code tree has at most 256 internal nodes, no 1-bit code
        uint32_t cnt[32], ilo[32], ihi[32], base[32];
        for (int d = 0; d < 32; d++) cnt[d] = (d >= min_len && d <= max_len) ? s_lencount[d] : 0u;
        tab->fsm_states = (status == DC_OK && bpd != 0 && max_len < 16) ? fsm_geometry(s_first, cnt, min_len, max_len, bpd, n_ary, ilo, ihi, base) : 0;
        if (host_meta) {   // mapped host memory (dc_common.cuh, table_meta_*): what the entry points pick their kernels by
            const int32_t *words = (const int32_t *)tab;
            for (int i = 0; i < 10; i++) host_meta[i] = words[i];
            host_meta[10] = tab->lut2_used;
            host_meta[11] = tab->fsm_states;
            __threadfence_system();
        }
    }
}

__global__ void bits_for_hist_kernel(const unsigned long long *__restrict__ d_hist, const dc_huff_table *__restrict__ tab,
                                     unsigned long long *__restrict__ d_bits) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const int b = threadIdx.x;
    if (b < 256) {
        const unsigned long long c = d_hist[b];
        if (c) atomicAdd(&s_sum, c * (unsigned long long)(tab->lengths[b] * tab->bits_per_digit));
    }
    __syncthreads();
    if (threadIdx.x == 0) *d_bits = s_sum;
}

int launch_table(const unsigned long long *d_hist, const int32_t *d_lengths, int nsym, int n_ary, dc_huff_table *tab,
                 TableRaw raw, cudaStream_t st, int nhist, int hist_stride) {
    if (n_ary < 2 || n_ary > 512 || nsym < 1 || nsym > DC_MAX_LEAVES) return DC_ERR_ARG;
    if (tab && nsym != DC_NSLOTS) return DC_ERR_ARG;
    TableMetaTicket ticket;
    ticket.dev = nullptr;
    if (tab) table_meta_begin(tab, &ticket);
    {
        LaunchScope ls(DC_K_TABLE, st);
        table_kernel<<<1, kTabThreads, 0, st>>>(d_hist, d_lengths, nsym, n_ary, tab, raw, ticket.dev, nhist < 1 ? 1 : nhist, hist_stride);
    }
    if (tab) table_meta_end(tab, ticket, st);
    return cuda_status(cudaGetLastError());
}

}  // namespace dc

using namespace dc;

extern "C" int dc_huff_build(const uint64_t *d_hist, int n_ary, dc_huff_table *d_table, void *stream) {
    if (!d_hist || !d_table) return DC_ERR_ARG;
    TableRaw raw = {nullptr, nullptr, nullptr, nullptr};
    return launch_table((const unsigned long long *)d_hist, nullptr, DC_NSLOTS, n_ary, d_table, raw, (cudaStream_t)stream);
}

extern "C" int dc_huff_table_from_lengths(const int32_t *d_lengths, int n_ary, dc_huff_table *d_table, void *stream) {
    if (!d_lengths || !d_table) return DC_ERR_ARG;
    TableRaw raw = {nullptr, nullptr, nullptr, nullptr};
    return launch_table(nullptr, d_lengths, DC_NSLOTS, n_ary, d_table, raw, (cudaStream_t)stream);
}

extern "C" int dc_huff_table_forget(const dc_huff_table *d_table) {
    if (!d_table) return DC_ERR_ARG;
    table_meta_forget(d_table);
    return DC_OK;
}

extern "C" int dc_huff_table_download(const dc_huff_table *d_table, dc_huff_table *h_table, void *stream) {
    if (!d_table || !h_table) return DC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    DC_CUDA_TRY(cudaMemcpyAsync(h_table, d_table, sizeof(dc_huff_table), cudaMemcpyDeviceToHost, st));
    DC_CUDA_TRY(cudaStreamSynchronize(st));
    return DC_OK;
}

extern "C" int dc_huff_bits_for_hist(const uint64_t *d_hist, const dc_huff_table *d_table, uint64_t *d_bits, void *stream) {
    if (!d_hist || !d_table || !d_bits) return DC_ERR_ARG;
    LaunchScope ls(DC_K_BITS_FOR_HIST, (cudaStream_t)stream);
    bits_for_hist_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const unsigned long long *)d_hist, d_table,
                                                              (unsigned long long *)d_bits);
    return cuda_status(cudaGetLastError());
}
