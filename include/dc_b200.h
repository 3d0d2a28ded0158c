/*
 * dc_b200.h -- C-ABI of the B200-native n-ary Huffman / nybble hot path (libdc_b200.so).
 *
 * Drop-in boundary for the data-parallel path of carycode/data_compression
 * (n_ary_huffman.c, nybble_compression.c).  The reference has no plugin/FFI layer
 * (SURVEY 8b): its boundary is its non-static C functions.  Each entry point below
 * cites the reference interface it replaces (file:line into the reference tree).
 *
 * Two families:
 *   dc_*          size-explicit, DEVICE pointers, stream-ordered, never block
 *                 (except where stated), never allocate, never print, never abort.
 *   dc_host_*     HOST pointers: copy in, run the device path, copy out, synchronise.
 *                 These are what a C caller of the reference functions links against;
 *                 refapi.h maps the reference's verbatim names onto them.
 *
 * There is no CPU fallback: every function that computes launches sm_100a kernels and
 * returns DC_ERR_CUDA when no device is usable.
 *
 * Plain C types only (no CUDA, no torch types).  A stream is passed as `void *`
 * (a cudaStream_t; NULL = the legacy default stream).
 */
#ifndef DC_B200_H
#define DC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- constants */

#define DC_MAX_SYMBOL_VALUE 258 /* n_ary_huffman.c:2524 */
#define DC_NSLOTS 259           /* max_symbol_value + 1 histogram / length slots */
#define DC_MAX_LEAVES 512       /* dc_host_huffman accepts max_leaf_value < DC_MAX_LEAVES */
#define DC_LUT_BITS 12          /* decode look-up table index width */
#define DC_LUT2_SUBTABLES 256   /* second-level tables of 16 entries (codes of 13..16 bits) */
#define DC_LUT14_BITS 14        /* index width of the count table for tables whose longest code has 13 or 14 bits */
#define DC_TRIT_WINDOW 8        /* radix 3: trits a decode look-up sees (index = their base-3 value) */
#define DC_LUT_ENTRIES 6564     /* entries of the multi-symbol tables: max(2^12, 3^8), rounded up to a multiple of 4 */
/* entries of the multi-symbol tables that do not resolve in one look-up: one compare tells them apart ((int32)count entry
 * < 0, pair entry >= 0xC0000000); the low 16 bits name a second-level table, DC_LUT_NO_SUBTABLE = none (canonical search) */
#define DC_LUT_COUNT_MARK 0xFF000000u
#define DC_LUT_PAIR_MARK 0xC0000000u
#define DC_LUT_NO_SUBTABLE DC_LUT2_SUBTABLES   /* an extra, empty sub-table */

enum dc_status {
    DC_OK = 0,
    DC_ERR_ARG = -1,           /* bad argument (null pointer, misalignment, radix, phase) */
    DC_ERR_CUDA = -2,          /* CUDA runtime error or no device (no CPU fallback exists) */
    DC_ERR_CODE_TOO_LONG = -3, /* reference limits: length < 16 digits (:1414), value fits int (:1540) */
    DC_ERR_CAPACITY = -4,      /* output buffer too small */
    DC_ERR_CORRUPT = -5,       /* bitstream hits an unused code slot / ends inside a code */
    DC_ERR_SYMBOL = -6,        /* input symbol has no code (length 0), or nibble symbol >= 16 (:1093) */
    DC_ERR_RADIX = -7,         /* payload packing is defined for radices n <= 16 only */
    DC_ERR_NCCL = -8           /* NCCL could not be loaded (libnccl.so.2 / $DC_NCCL_LIB) or a collective failed */
};

/*
 * Device-resident code table.  Written by dc_huff_build / dc_huff_table_from_lengths, read by the
 * encode/decode kernels.  The layout is part of the ABI so a host may cudaMemcpy it back and read
 * lengths/values (dc_huff_table_download does exactly that).
 *   lengths[]  == canonical_lengths of huffman()                       n_ary_huffman.c:1161-1208
 *   values[]   == encode_value_table of convert_lengths_to_encode_table n_ary_huffman.c:1382-1612
 * Lengths are in DIGITS of radix n_ary (the reference's unit); a digit is bits_per_digit bits.
 * The two builders are the only writers: they also leave the table's header in host-visible (mapped) memory,
 * keyed by the table's device address, so that the encoder and the decoder can choose their kernels without
 * reading the table back.  A table that reaches a device buffer any other way (cudaMemcpy of a downloaded
 * copy into a FRESH buffer) works too, through a blocking read; do not overwrite a built table behind the
 * library's back -- build into the buffer again instead.
 */
typedef struct dc_huff_table {
    int32_t n_ary;            /* compressed_symbols */
    int32_t bits_per_digit;   /* 1, 2, 4 for n = 2, 4, 16; 2 for n = 3 (see packed_radix); 4 for n = 5 .. 15 (one nibble per digit,
                               * most significant digit first -- the stream is the payload; such tables carry no window LUTs and
                               * are decoded by the byte-stepped state machine only); 0 = table only (n > 16: no payload packing) */
    int32_t max_symbol_value; /* 258 */
    int32_t nonzero_symbols;  /* :880-886 */
    int32_t dummy_nodes;      /* :900-903, as written */
    int32_t min_len;          /* digits, over non-zero lengths (:1354-1379) */
    int32_t max_len;          /* digits (:1330-1352) */
    int32_t max_bits;         /* max_len * bits_per_digit */
    int32_t status;           /* DC_OK or DC_ERR_CODE_TOO_LONG / DC_ERR_RADIX */
    int32_t packed_radix;     /* 0: the kernels' stream is the payload; 3: n = 3, the stream has 2 bits per trit and
                               * dc_trit_pack() / dc_trit_unpack() convert it to / from the 5-trits-per-byte payload */
    uint64_t total_symbols;   /* sum of the histogram the table was built from (0 if from lengths) */
    uint64_t total_bits;      /* sum hist[s] * lengths[s] * bits_per_digit (0 if from lengths) */
    int32_t lengths[DC_NSLOTS + 1];
    uint32_t values[DC_NSLOTS + 1];
    uint32_t enc[256];        /* per byte: (value << 6) | nbits, valid when max_bits <= 26 */
    uint64_t enc64[256];      /* per byte: value | (uint64)nbits << 32 */
    /* canonical decode: per digit-length first code value, number of codes, offset into sorted[] */
    uint32_t first_code[32];
    uint32_t len_count[32];
    uint32_t len_offset[32];
    uint16_t sorted[DC_NSLOTS + 1]; /* symbols in (length, value) order */
    uint16_t lut[1 << DC_LUT_BITS]; /* index = next 12 bits: nbits << 8 | symbol; 0 = escape (longer/unused) */
    /* multi-symbol tables, same index: every code that lies completely inside the 12 bits.  Radix 3 (packed_radix == 3):
     * the index is the base-3 value of the next DC_TRIT_WINDOW = 8 trits (0 .. 6560) and the entries cover every code
     * inside those 8 trits; bit counts are those of the 2-bit-per-trit stream.
     *   lut_count: total bits | count << 16 | first code's bits << 24            (DC_LUT_COUNT_MARK | x = escape)
     *   lut_pair : symbol0 | symbol1 << 8 | bits of (up to) two codes << 16 | first code's bits << 24 (5 bits)
 *              | unused-slot flag << 29 | count(0..2) << 30                     (DC_LUT_PAIR_MARK | x = escape) */
    uint32_t lut_count[DC_LUT_ENTRIES];
    uint32_t lut_pair[DC_LUT_ENTRIES];
    /* second level, for codes of 13..16 bits: a 12-bit window that is the prefix of such codes has the marker entry
     *   lut_count = DC_LUT_COUNT_MARK | subtable,  lut_pair = DC_LUT_PAIR_MARK | subtable
     * and lut2[subtable * 16 + next 4 bits] = code bits << 8 | symbol (0 = no code of <= 16 bits there).  Longer codes
     * and unused slots have the subtable DC_LUT_NO_SUBTABLE = canonical search. */
    uint16_t lut2[(DC_LUT2_SUBTABLES + 1) * 16];
    int32_t lut2_used;        /* subtables in use */
    int32_t fsm_states;       /* internal nodes of the code tree if the byte-stepped decoder applies (k4_fsm.cuh), else 0 */
    /* the decoder's synchronisation pass only counts codes; for tables whose longest code has 13 or 14 bits (binary
     * codes of byte data, typically) it uses this 14-bit-indexed table instead, which needs no escape:
     * total bits | count << 8 | first code's bits << 12 of every code inside the next 14 bits.  Filled only then. */
    uint16_t lut14[1 << DC_LUT14_BITS];
} dc_huff_table;

/* ------------------------------------------------------------------------- library */

const char *dc_version(void);
const char *dc_status_string(int status);
/* number of CUDA devices visible, or a negative dc_status */
int dc_device_count(void);
/* total kernels this library has launched in this process (bench "gpu_launches") */
uint64_t dc_launch_count(void);

/* ------------------------------------------------------------------------- tracing */

/* kernel groups of this library (the reference's "# ..." printf narration and `make time_test` are
 * replaced by CUDA events around every launch) */
enum dc_kernel_id {
    DC_K_HISTOGRAM = 0,
    DC_K_TABLE,
    DC_K_BITS_FOR_HIST,
    DC_K_ENCODE_COUNT,
    DC_K_ENCODE_SCAN,
    DC_K_ENCODE,
    DC_K_ENCODE_MID,
    DC_K_ENCODE_WIDE,
    DC_K_DECODE_SYNC,
    DC_K_DECODE_HANDOFF,
    DC_K_DECODE_SCAN,
    DC_K_DECODE_WRITE,
    DC_K_DECODE_FAST_SYNC,
    DC_K_DECODE_FAST_SCAN,
    DC_K_DECODE_FAST_WRITE,
    DC_K_NYBBLE_PACK,
    DC_K_NYBBLE_UNPACK,
    DC_K_NYBBLE_TAIL,
    DC_K_TEXT_SUMMARY,
    DC_K_TEXT_SCAN,
    DC_K_TEXT_EMIT,
    DC_K_TRIT_PACK,
    DC_K_TRIT_UNPACK,
    DC_K_B64_PACK,
    DC_K_B64_UNPACK,
    DC_K_MTF_WALK,
    DC_K_MTF_SCAN,
    DC_K_MTF_RESOLVE,
    DC_K_TEXT_BATCH,
    DC_K_SYNTH,
    DC_K_DECODE_FSM_BUILD,
    DC_K_DECODE_FSM_SYNC,
    DC_K_DECODE_FSM_WRITE,
    DC_K_ENCODE_PLAN,
    DC_K_ENCODE_FAST,
    DC_K_SHARD_EXCHANGE,   /* the shard layer's exchange: peer-memory kernel or ncclAllGather (incl. the wait for the slowest rank) */
    DC_K_SHARD_PLAN,       /* every rank's bit total + exclusive scan; the edge-byte completion */
    DC_K_COUNT
};
/* on != 0: bracket every kernel launch with CUDA events on its launching stream */
int dc_profile_enable(int on);
/* drop accumulated timings */
int dc_profile_reset(void);
/* blocking: total device milliseconds and launch count of one kernel group since the last reset */
int dc_profile_kernel(int kernel_id, double *total_ms, uint64_t *launches);
const char *dc_profile_kernel_name(int kernel_id);

/* ------------------------------------------------------------------------- K1 histogram */

/*
 * Replaces histogram() n_ary_huffman.c:461-493 (length-explicit, 0x00 allowed).
 * d_hist[0..258] is overwritten (the reference zeroes it first, :474-476); slots 256..258 stay 0.
 */
int dc_histogram_u8(const uint8_t *d_in, size_t n, uint64_t *d_hist, void *stream);

/* ------------------------------------------------------------------------- K2 table */

/*
 * Replaces huffman() :1161-1208 + convert_lengths_to_encode_table() :1382-1612.
 * d_hist: 259 x u64 (e.g. the all-reduced histogram).  Single-CTA kernel; reproduces the as-written
 * dummy rule d = (n-1) - ((nz-1) % (n-1)) and the stable-sort tie-break exactly.
 * n_ary >= 2; payload packing needs n_ary in {2,4,16}, other radices produce lengths/values only.
 */
int dc_huff_build(const uint64_t *d_hist, int n_ary, dc_huff_table *d_table, void *stream);

/* Decode side: rebuild values/LUT from 259 lengths (digits), as read from a table header (:1727-1744). */
int dc_huff_table_from_lengths(const int32_t *d_lengths, int n_ary, dc_huff_table *d_table, void *stream);

/* The library keeps the header of every table it built, keyed by the table's device address (see dc_huff_table above).
 * Call this before the memory behind `d_table` is freed or reused for anything but another build -- an allocator may hand
 * the same address out again.  (A stale header is caught where it can be: the kernels it selects refuse a table that is
 * not theirs -- the decoder then takes its slow path, the encoder reports DC_ERR_ARG -- but forgetting is the contract.) */
int dc_huff_table_forget(const dc_huff_table *d_table);

/* Blocking copy of the table to host memory. */
int dc_huff_table_download(const dc_huff_table *d_table, dc_huff_table *h_table, void *stream);

/* bits a shard will emit for its LOCAL histogram under a (global) table: sum hist[s]*len[s]*bpd.
 * Lets every GPU know its bit total before encoding.  d_bits: 1 x u64. */
int dc_huff_bits_for_hist(const uint64_t *d_hist, const dc_huff_table *d_table, uint64_t *d_bits, void *stream);

/* ------------------------------------------------------------------------- K3 encode */

size_t dc_huff_encode_workspace_bytes(size_t n);

/*
 * Replaces represent_items_with_codes() :1621-1678 (a stub in the reference; the payload layout is the
 * one DESIGN.md defines: codes MSB-first, concatenated in input order, final byte zero-padded).
 *   d_out        16-byte aligned; byte 0 is the byte the first code bit lands in
 *   bit_phase    0..7: number of leading bits of d_out[0] owned by the previous shard (written as 0)
 *   d_total_bits 1 x u64: code bits emitted (excluding bit_phase)
 *   d_status     1 x i32: DC_OK / DC_ERR_SYMBOL / DC_ERR_CAPACITY / table status; may be NULL
 * Bytes [0, ceil((bit_phase+total_bits)/8)) of d_out are written, nothing else.
 */
int dc_huff_encode(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out,
                   size_t out_capacity, unsigned bit_phase, uint64_t *d_total_bits, int32_t *d_status,
                   void *d_workspace, size_t workspace_bytes, void *stream);

/*
 * The same encode with the run offsets PLANNED from the histogram pass instead of counted from the input a second
 * time.  dc_histogram_u8_runs() is dc_histogram_u8() that also leaves one 256 x u16 histogram per 32 KB run of the
 * input in the encode workspace (1.6 % of n); dc_huff_encode_planned() then derives the bit offset of every run from
 * those and the code lengths (sum count * length: no data pass) and encodes in one read of the input.  Call order:
 *   dc_histogram_u8_runs(d_in, n, d_hist, ws, ws_bytes, s);  [all-reduce d_hist;]  dc_huff_build(...);
 *   dc_huff_encode_planned(d_in, n, table, ..., ws, ws_bytes, s);     -- same d_in, n and workspace
 * Arguments, results and status codes are those of dc_huff_encode(); d_in must be 16-byte aligned for both calls.
 */
int dc_histogram_u8_runs(const uint8_t *d_in, size_t n, uint64_t *d_hist, void *d_encode_workspace, size_t workspace_bytes,
                         void *stream);
int dc_huff_encode_planned(const uint8_t *d_in, size_t n, const dc_huff_table *d_table, uint8_t *d_out,
                           size_t out_capacity, unsigned bit_phase, uint64_t *d_total_bits, int32_t *d_status,
                           void *d_workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------- K4 decode */

size_t dc_huff_decode_workspace_bytes(uint64_t bit_start, uint64_t nbits);

/*
 * Replaces the absent Huffman block decoder (decompress() case 'X'/'Z', :2081-2089).
 * Self-synchronising parallel decode of `nbits` code bits that start `bit_start` (< 128) bits into
 * d_bits (16-byte aligned; readable up to ceil((bit_start+nbits)/8) rounded up to a multiple of 16 bytes: the kernels load
 * 16-byte vectors).  Exactly n_out symbols
 * are expected.  May block on the stream internally while neighbouring tiles resynchronise.
 *   d_status     1 x i32: DC_OK / DC_ERR_CORRUPT / DC_ERR_CAPACITY; may be NULL
 */
int dc_huff_decode(const uint8_t *d_bits, uint64_t bit_start, uint64_t nbits, const dc_huff_table *d_table,
                   uint8_t *d_out, size_t n_out, int32_t *d_status, void *d_workspace, size_t workspace_bytes,
                   void *stream);

/* ---- opt-in: a stream with its index
 *
 * dc_huff_decode works from the bitstream alone: its first pass (F1) finds where the codes of every 256-bit subsequence
 * start and how many there are.  A producer that controls its container can keep exactly that -- the index: 2 bytes per 32
 * bytes of stream plus 8 bytes per 16 KB, 6.3 % of the payload -- and the consumer then runs the write pass only
 * (0.84 ms instead of 1.33 ms per GiB at n = 4).  The blind path stays the default and is what a stream produced elsewhere
 * needs (BASELINE config 5).
 *   dc_huff_index_build    runs F1 + F2 on a finished stream and leaves the index in d_index; fills *info (HOST memory,
 *                          to be kept with the index).  Blocking.  Returns 1 if the stream does not self-synchronise
 *                          (no index can be built: decode it with dc_huff_decode).
 *   dc_huff_decode_indexed the write pass from the index; stream-ordered.  The stream, the table and the geometry in *info
 *                          must be the ones the index was built for (DC_ERR_ARG otherwise, where the host can tell).
 */
typedef struct dc_huff_index_info {
    uint64_t magic, bit_start, nbits, n_symbols;
    uint32_t mode, start_token, reserved[2];
} dc_huff_index_info;
size_t dc_huff_index_bytes(uint64_t bit_start, uint64_t nbits);
int dc_huff_index_build(const uint8_t *d_bits, uint64_t bit_start, uint64_t nbits, const dc_huff_table *d_table, uint64_t n_symbols,
                        void *d_index, size_t index_bytes, dc_huff_index_info *info, void *d_workspace, size_t workspace_bytes,
                        void *stream);
int dc_huff_decode_indexed(const uint8_t *d_bits, const dc_huff_index_info *info, const dc_huff_table *d_table, const void *d_index,
                           size_t index_bytes, uint8_t *d_out, size_t n_out, int32_t *d_status, void *d_workspace,
                           size_t workspace_bytes, void *stream);

/* ---- one contiguous byte range ("shard") of a longer bitstream, e.g. one GPU's part (BASELINE config 5)
 *
 * The stream is cut at multiples of 1024 bytes.  A shard does not know where its first code begins; it finds it the way
 * every 16 KB segment inside a stream does, by walking the last 1024 bytes of the PREVIOUS shard until the codes have
 * self-synchronised, and reports what it assumed and where its own last code ends.  The caller exchanges the summaries
 * (an all-gather of 24 bytes per shard), checks assumed_start[g] == exit[g-1] for every g, turns the symbol counts into
 * output offsets with an exclusive scan, and calls the write phase.  A shard whose assumption was wrong (a stream that
 * does not self-synchronise within 8192 bits) is re-run with has_halo = 0 and first_code_bit = exit[g-1].
 *
 * Geometry, the same for both calls:
 *   d_bits            first byte of the shard, 16-byte aligned, at a multiple of 1024 bytes of the stream
 *   has_halo          != 0: the 1024 bytes in front of d_bits are readable and hold the previous shard's tail
 *   first_code_bit    has_halo == 0 only: exact bit offset (< 128) of the first code that starts in the shard, or the
 *                     previous shard's `exit` as reported
 *   shard_bits        bits of the stream that belong to this shard (a multiple of 8192 for all but the last shard)
 *   stream_bits_left  bits from d_bits to the end of the stream (>= shard_bits).  Readable memory: the kernels load 16-byte
 *                     vectors and fetch one 1 KB tile ahead, so the shard's bytes rounded up to 16 must be readable, and behind
 *                     a NON-final shard another 1024 bytes (the next shard's head)
 */
typedef struct dc_shard_summary {
    uint64_t symbols;        /* codes that START in the shard */
    uint32_t exit;           /* where the shard's last code ends: bits it reaches into the next shard, or (byte-stepped decoder)
                              * 0x80000000 | the code-tree state at the shard's end.  Opaque: compare with the next shard's
                              * assumed_start, feed back as first_code_bit */
    int32_t resync;          /* != 0: some 16 KB segment inside the shard did not self-synchronise; decode the stream whole */
    uint32_t assumed_start;  /* what the shard assumed in front of its first byte (same encoding as exit; == first_code_bit without a halo) */
    uint32_t reserved;
} dc_shard_summary;

int dc_huff_decode_shard_sync(const uint8_t *d_bits, int has_halo, unsigned first_code_bit, uint64_t shard_bits,
                              uint64_t stream_bits_left, const dc_huff_table *d_table, dc_shard_summary *d_summary,
                              void *d_workspace, size_t workspace_bytes, void *stream);
/* writes the shard's symbols to d_out[0 .. n_out); n_out must equal the summary's symbol count.  Uses the workspace
 * the sync phase left behind. */
int dc_huff_decode_shard_write(const uint8_t *d_bits, int has_halo, uint64_t shard_bits, uint64_t stream_bits_left,
                               const dc_huff_table *d_table, uint8_t *d_out, size_t n_out, int32_t *d_status,
                               void *d_workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------- shards over several GPUs (NCCL)
 *
 * One process per GPU; every call below is made by all ranks of a communicator with the current device set to the
 * rank's GPU (SURVEY 8b last cell, 8e; BASELINE configs 4 and 5).  NCCL is bound at run time (dlopen of libnccl.so.2,
 * or the library named by $DC_NCCL_LIB); without it these return DC_ERR_NCCL.  The data path has no bulk collective:
 * the ranks exchange histograms with edge symbols (2 KB), 1 KB halos and 24-byte summaries.
 */
typedef struct dc_shard_comm dc_shard_comm;
/* rank 0 makes an id (128 bytes = ncclUniqueId), hands it to the other ranks by any means, everybody creates */
int dc_shard_unique_id(void *id128);
int dc_shard_comm_create(const void *id128, int rank, int world, dc_shard_comm **out);
/* or: wrap a communicator the caller already has (an ncclComm_t passed as void *); it is not destroyed with the wrapper */
int dc_shard_comm_from_nccl(void *nccl_comm, int rank, int world, dc_shard_comm **out);
int dc_shard_comm_destroy(dc_shard_comm *comm);
int dc_shard_comm_rank(const dc_shard_comm *comm);
int dc_shard_comm_world(const dc_shard_comm *comm);

/*
 * Encode one logical stream whose bytes are spread over the ranks in rank order (BASELINE config 4).  Stream-ordered,
 * never blocks: local histogram -> ONE all-gather (the local histograms, symbol counts, first and last eight symbols) ->
 * global table (d_table, identical on every rank) and every rank's bit total -> encode at bit phase O_r mod 8 (O_r = bits
 * of the ranks in front) -> the bytes that neighbouring shards share are completed locally (a rank re-creates its
 * neighbours' few bits from their edge symbols).  d_out then holds stream bytes [O_r / 8, ceil((O_r + bits_r) / 8)); the
 * concatenation over the ranks IS the single-stream payload of dc_huff_encode on the concatenated input.
 *   d_total_bits  this rank's code bits (1 x u64, may be NULL);  d_status as dc_huff_encode
 * The exchange goes through peer memory (every rank stores its 2 KB into buffers its peers have mapped with CUDA IPC, over
 * NVLink, and raises a flag; one single-CTA kernel) when the ranks can map each other's buffers, else through ncclAllGather
 * (also under DC_SHARD_PEER=0).  The FIRST call on a communicator sets this up and blocks once (two small all-gathers);
 * calls on one communicator must be issued in the same order on every rank and on one stream.
 * dc_shard_huff_encode_info (blocking) reads O_r, bits_r and the stream's bit total back from the workspace;
 * dc_shard_huff_gather (blocking) places all shards in one buffer on `root`.
 */
size_t dc_shard_huff_encode_workspace_bytes(size_t n_local, int world);
int dc_shard_huff_encode(dc_shard_comm *comm, const uint8_t *d_in, size_t n_local, int n_ary, dc_huff_table *d_table,
                         uint8_t *d_out, size_t out_capacity, uint64_t *d_total_bits, int32_t *d_status, void *d_workspace,
                         size_t workspace_bytes, void *stream);
int dc_shard_huff_encode_info(const void *d_workspace, size_t n_local, int world, uint64_t *bit_offset, uint64_t *bits,
                              uint64_t *total_bits, void *stream);
int dc_shard_huff_gather(dc_shard_comm *comm, int root, const uint8_t *d_shard, const void *d_workspace, size_t n_local,
                         uint8_t *d_stream, size_t stream_capacity, void *stream);

/*
 * Decode ONE stream of total_bits bits that is cut blindly into byte ranges (BASELINE config 5): rank r holds stream bytes
 * [r * part_bytes, ...), part_bytes a multiple of 1024 and the same on every rank (the last ranks may hold less or
 * nothing).  d_buf = [1024 bytes headroom | the rank's bytes | >= 1024 bytes tailroom], 16-byte aligned; the halos are
 * exchanged here.  No side information about code boundaries: every rank synchronises over its left neighbour's tail,
 * the summaries are all-gathered, a rank whose assumption was wrong starts over from its neighbour's real exit.  Blocking.
 * On return d_out holds *n_symbols symbols that belong at *symbol_offset of the decoded output (*total_symbols in all).
 */
size_t dc_shard_huff_decode_workspace_bytes(size_t part_bytes, int world);
int dc_shard_huff_decode_stream(dc_shard_comm *comm, uint8_t *d_buf, size_t part_bytes, uint64_t total_bits,
                                const dc_huff_table *d_table, uint8_t *d_out, size_t out_capacity, uint64_t *n_symbols,
                                uint64_t *symbol_offset, uint64_t *total_symbols, int32_t *d_status, void *d_workspace,
                                size_t workspace_bytes, void *stream);

/*
 * Nybble pack / unpack over shards (write_nybble nybble_compression.c:1091-1114, split :767-773): no
 * exchange at all -- shard starts are kept on even symbol indices so that no packed byte is shared.
 * dc_shard_nybble_range gives the symbols [lo, hi) of n_total that `rank` takes; the pack / unpack calls work on the
 * shard's own symbols (d_sym -> symbol lo) and packed bytes (d_packed -> byte lo / 2).
 */
int dc_shard_nybble_range(uint64_t n_total, int rank, int world, uint64_t *lo, uint64_t *hi);
int dc_shard_nybble_pack(uint64_t n_total, int rank, int world, const uint8_t *d_sym, uint8_t *d_packed, int32_t *d_status,
                         void *stream);
int dc_shard_nybble_unpack(uint64_t n_total, int rank, int world, const uint8_t *d_packed, uint8_t *d_sym, void *stream);

/* ------------------------------------------------------------------------- K5 nybble */

/*
 * Stream form of write_nybble() nybble_compression.c:1091-1114: symbol 2i -> high nibble of byte i,
 * symbol 2i+1 -> low nibble; an odd tail leaves the low nibble 0.  Symbols must be < 16 (:1093);
 * a larger value sets *d_status = DC_ERR_SYMBOL (its low nibble is packed).  d_status may be NULL.
 */
int dc_nybble_pack(const uint8_t *d_sym, size_t n_sym, uint8_t *d_packed, int32_t *d_status, void *stream);
/* Inverse: the decoder's split nybble_compression.c:767-769 (high nibble first). */
int dc_nybble_unpack(const uint8_t *d_packed, size_t n_sym, uint8_t *d_sym, void *stream);

/* ------------------------------------------------------------------------- K7 trit payload (radix 3) */

/*
 * Radix 3 is the reference's default (n_ary_huffman.c:2529); its payload is sketched at :745-748: 5 trits per byte, byte =
 * 1 + the group's base-3 value (most significant trit first), last group padded with zero trits.  For a table built
 * with n_ary == 3 (packed_radix == 3) dc_huff_encode writes, and dc_huff_decode reads, an intermediate stream with one
 * 2-bit field per trit ("T2": total_bits == 2 * trits); these two convert between that stream and the payload of
 * ceil(trits / 5) bytes.  d_t2: 4-byte aligned (and, for dc_huff_decode, readable up to the next multiple of 16 bytes);
 * d_payload: 16-byte aligned for dc_trit_pack.  *d_status = DC_ERR_CORRUPT for a field of 3 / a byte outside 1..243.
 */
int dc_trit_pack(const uint8_t *d_t2, uint64_t ntrits, uint8_t *d_payload, int32_t *d_status, void *stream);
int dc_trit_unpack(const uint8_t *d_payload, uint64_t ntrits, uint8_t *d_t2, int32_t *d_status, void *stream);

/*
 * The base64url text form of a binary payload: the reference's unfinished packer emits 6 bits per character through
 * int2digit() (n_ary_huffman.c:371-426, :1646-1671).  Character k = int2digit(bits [6k, 6k + 6), most significant first,
 * zero padded) -- RFC 4648 without '=' padding; ceil(nbits / 6) characters.  dc_base64url_unpack accepts what digit2int()
 * accepts (:428-455: '-' or '+' for 62, '_' or '/' for 63), writes ceil(nbits / 8) bytes and sets *d_status =
 * DC_ERR_CORRUPT for any other character.  d_bits: 4-byte aligned; d_chars: 16-byte aligned for dc_base64url_pack.
 */
int dc_base64url_pack(const uint8_t *d_bits, uint64_t nbits, uint8_t *d_chars, void *stream);
int dc_base64url_unpack(const uint8_t *d_chars, uint64_t nbits, uint8_t *d_bits, int32_t *d_status, void *stream);

/* ------------------------------------------------------------------------- K6 static-table nybble compressor */

size_t dc_nybble_text_workspace_bytes(size_t n);

/*
 * Replaces compress_bytestring(source, dest, false) nybble_compression.c:887-1038 (static " etaoins" table; for
 * modify == true see dc_nybble_adaptive_compress below).  Length-explicit: d_src[0..n) must hold bytes 0x01..0x7F (:910), else *d_status = DC_ERR_SYMBOL.
 * Output: 0xAF, src[0], body -- or ' ' + source when that is not shorter (:1018-1037) -- and a NUL when there is room.
 *   d_out_len  1 x u64: bytes written (without the NUL);  needs dst_capacity >= n + 1
 */
int dc_nybble_text_compress(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t dst_capacity, uint64_t *d_out_len,
                            int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream);
/* Replaces decompress_bytestring(source, dest, false) :734-817: type byte 0xAF / ' ' / anything else; at most 2n - 3 bytes. */
int dc_nybble_text_decompress(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t dst_capacity, uint64_t *d_out_len,
                              int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------- K6 + K8 adaptive nybble compressor */

/*
 * Replace compress_bytestring(source, dest, true) = nybble_compress() nybble_compression.c:1134 and
 * decompress_bytestring(source, dest, true) = nybble_decompress() :1117: the same coder with 16 move-to-front
 * contexts (byte_to_context :517, update_context :665-687) instead of the static table.  Same arguments, limits and
 * status codes as the two functions above; the workspace is larger (dc_nybble_adaptive_workspace_bytes).
 * Compress is a parallel scan (the contexts are known from the input).  Decompress resolves the hit nibbles with one
 * serial walk over the output on the device -- the chain the format imposes; it is parallel only across strings: ONE
 * string decompresses at about 11 MB/s (a CPU core is faster).  The supported fast path for the adaptive mode is the batch
 * form below (dc_nybble_text_decompress_batch: one thread per string, 40 GB/s over many short strings).
 */
size_t dc_nybble_adaptive_workspace_bytes(size_t n);
int dc_nybble_adaptive_compress(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t dst_capacity, uint64_t *d_out_len,
                                int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream);
int dc_nybble_adaptive_decompress(const uint8_t *d_src, size_t n, uint8_t *d_dst, size_t dst_capacity, uint64_t *d_out_len,
                                  int32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream);

/*
 * Many strings per call -- the parallelism SURVEY 8e gives the adaptive mode ("replicas only"), and the reference's own
 * use (many short strings).  String i is d_src[d_src_off[i] .. d_src_off[i + 1]), its output slot
 * d_dst[d_dst_off[i] .. d_dst_off[i + 1]) (compress: >= length + 2 bytes, decompress: >= 2 * length), and
 * d_out_len[i] gets its output length; one thread walks one string as compress_bytestring() / decompress_bytestring()
 * do (modify = 0: static table, 1: adaptive contexts).  *d_status: the first error of any string.
 */
int dc_nybble_text_compress_batch(const uint8_t *d_src, const uint64_t *d_src_off, size_t count, int modify, uint8_t *d_dst,
                                  const uint64_t *d_dst_off, uint64_t *d_out_len, int32_t *d_status, void *stream);
int dc_nybble_text_decompress_batch(const uint8_t *d_src, const uint64_t *d_src_off, size_t count, int modify, uint8_t *d_dst,
                                    const uint64_t *d_dst_off, uint64_t *d_out_len, int32_t *d_status, void *stream);

/* ------------------------------------------------------------------------- synthetic inputs (bench/tests) */

/*
 * Counter-based generator shared by bench.py and the tests: byte i = value_base + rank, where
 * rank = #{k : thresholds[k] <= (splitmix64(seed + i) >> 32)} over `nthresh` ascending u32 thresholds
 * (d_thresholds on device).  The same arithmetic in numpy reproduces the stream on the host.
 */
int dc_synth_fill(uint8_t *d_out, size_t n, uint64_t seed, const uint32_t *d_thresholds, int nthresh,
                  int value_base, void *stream);

/* ------------------------------------------------------------------------- host-pointer entry points */

/* histogram(text, max_symbol_value, h) :461 -- NUL-terminated text, int counts, h[0..max_symbol_value] */
int dc_host_histogram(const char *text, int max_symbol_value, int h[]);
/* length-explicit host form */
int dc_host_histogram_u8(const uint8_t *in, size_t n, uint64_t h[DC_NSLOTS]);
/* huffman(max_leaf_value, freqs, compressed_symbols, lengths) :1161 */
int dc_host_huffman(int max_leaf_value, const int symbol_frequencies[], int compressed_symbols, int lengths[]);
/* 64-bit-count form */
int dc_host_huffman_u64(int max_leaf_value, const uint64_t symbol_frequencies[], int compressed_symbols,
                        int lengths[]);
/* convert_lengths_to_encode_table(...) :1382 */
int dc_host_convert_lengths_to_encode_table(int max_symbol_value, const int canonical_lengths[],
                                            int compressed_symbols, int encode_length_table[],
                                            unsigned int encode_value_table[]);
/* represent_items_with_codes(...) :1621 -- writes the payload at compressed_text[start...]; returns the
 * number of bytes written (>= 0) or a negative dc_status.  *total_bits (may be NULL) gets the bit count. */
int dc_host_represent_items_with_codes(int max_symbol_value, const int canonical_lengths[], int compressed_symbols,
                                       int bufsize, int original_length, const char original_text[], int start,
                                       char compressed_text[], uint64_t *total_bits);
/* whole encode path on host buffers: histogram -> table -> payload.  lengths_out[259] and total_bits
 * are the side information a decoder needs.  Returns bytes written or a negative dc_status.
 * compressed_symbols in {2, 4, 16}: total_bits = payload bits.  compressed_symbols == 3: total_bits = 2 * trits and the
 * payload is the 5-trits-per-byte form (K7), ceil(trits / 5) bytes. */
long long dc_host_huff_compress(const uint8_t *in, size_t n, int compressed_symbols, uint8_t *out,
                                size_t out_capacity, int lengths_out[DC_NSLOTS], uint64_t *total_bits);
/* inverse on host buffers */
int dc_host_huff_decompress(const uint8_t *payload, uint64_t total_bits, const int lengths[DC_NSLOTS],
                            int compressed_symbols, uint8_t *out, size_t n_out);
/* compress_bytestring(source, dest, modify) nybble_compression.c:887 / decompress_bytestring :734 on NUL-terminated
 * strings (dest must hold strlen(source) + 2, resp. 2 * strlen(source) + 1 bytes).  modify != 0 selects the adaptive
 * (move-to-front) contexts, as in nybble_compress() :1134 / nybble_decompress() :1117.  Returns the output length. */
long long dc_host_compress_bytestring(const char *source, char *dest, int modify);
long long dc_host_decompress_bytestring(const char *source, char *dest, int modify);
int dc_host_nybble_pack(const uint8_t *sym, size_t n_sym, uint8_t *packed);
int dc_host_nybble_unpack(const uint8_t *packed, size_t n_sym, uint8_t *sym);

#ifdef __cplusplus
}
#endif
#endif /* DC_B200_H */
