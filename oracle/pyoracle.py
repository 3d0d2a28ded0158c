"""ctypes/numpy binding of the CPU oracle (oracle/liboracle.so) and, when built, of the unmodified
reference compiled by oracle/Makefile (oracle/_ref/*.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (data_compression_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_HUFF_PATH = os.path.join(HERE, "_ref", "libref_huff.so")
REF_NYBBLE_PATH = os.path.join(HERE, "_ref", "libref_nybble.so")

NSLOTS = 259  # max_symbol_value = 258 (n_ary_huffman.c:2524)
MAX_SYMBOL_VALUE = 258

ORC_OK, ORC_ERR_ARG, ORC_ERR_CODE_TOO_LONG, ORC_ERR_CAPACITY, ORC_ERR_CORRUPT, ORC_ERR_SYMBOL = 0, -1, -3, -4, -5, -6


def build(force: bool = False) -> None:
    """Compile liboracle.so and (when /root/reference is present) oracle/_ref."""
    if force or not os.path.exists(LIB_PATH) or os.path.exists("/root/reference/n_ary_huffman.c"):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", HERE, "--no-print-directory"], check=True, env=env,
                       stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        u8p, u64p, i32p, u32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int),
                                 C.POINTER(C.c_uint))
        L.orc_histogram_cstr.argtypes = [C.c_char_p, C.c_int, u64p]
        L.orc_histogram_cstr.restype = None
        L.orc_histogram_u8.argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_int]
        L.orc_histogram_u8.restype = None
        L.orc_histogram_u8_mt.argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_int, C.c_int]
        L.orc_histogram_u8_mt.restype = None
        L.orc_huffman.argtypes = [C.c_int, u64p, C.c_int, i32p]
        L.orc_convert_lengths_to_encode_table.argtypes = [C.c_int, i32p, C.c_int, i32p, u32p]
        L.orc_bits_per_digit.argtypes = [C.c_int]
        L.orc_pack.argtypes = [C.c_void_p, C.c_size_t, i32p, u32p, C.c_int, C.c_uint, C.c_void_p, C.c_size_t, u64p]
        L.orc_pack_trits.argtypes = [C.c_void_p, C.c_size_t, i32p, u32p, C.c_void_p, C.c_size_t, u64p]
        L.orc_unpack_trits.argtypes = [C.c_void_p, C.c_uint64, C.c_int, i32p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_pack_mt.argtypes = [C.c_void_p, C.c_size_t, i32p, u32p, C.c_int, C.c_uint, C.c_void_p, C.c_size_t,
                                  u64p, C.c_int, u64p, C.c_size_t]
        L.orc_unpack.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, i32p, C.c_int, C.c_void_p,
                                 C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_unpack_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_int, i32p, C.c_int, C.c_void_p, C.c_size_t, u64p,
                                    C.c_size_t, C.c_int]
        for f in ("orc_nybble_pack", "orc_nybble_unpack"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
            getattr(L, f).restype = None
        for f in ("orc_nybble_pack_mt", "orc_nybble_unpack_mt"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
            getattr(L, f).restype = None
        for f in ("orc_nybble_static_compress", "orc_nybble_static_decompress",
                  "orc_nybble_adaptive_compress", "orc_nybble_adaptive_decompress"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
            getattr(L, f).restype = C.c_size_t
        L.orc_base64url_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_base64url_pack.restype = C.c_size_t
        L.orc_base64url_unpack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_base64url_unpack.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


# ----------------------------------------------------------------------------- restatement

def histogram_u8(data, nslots: int = NSLOTS, threads: int = 1) -> np.ndarray:
    d = _u8(data)
    h = np.zeros(nslots, dtype=np.uint64)
    if threads > 1:
        lib().orc_histogram_u8_mt(d.ctypes.data, d.size, _p(h, C.c_uint64), nslots, threads)
    else:
        lib().orc_histogram_u8(d.ctypes.data, d.size, _p(h, C.c_uint64), nslots)
    return h


def histogram_cstr(text: bytes, max_symbol_value: int = MAX_SYMBOL_VALUE) -> np.ndarray:
    h = np.zeros(max_symbol_value + 1, dtype=np.uint64)
    lib().orc_histogram_cstr(text, max_symbol_value, _p(h, C.c_uint64))
    return h


def huffman(freqs, n: int, max_leaf_value: int | None = None) -> np.ndarray:
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    mlv = f.size - 1 if max_leaf_value is None else max_leaf_value
    lengths = np.zeros(mlv + 1, dtype=np.int32)
    st = lib().orc_huffman(mlv, _p(f, C.c_uint64), n, _p(lengths, C.c_int))
    if st != ORC_OK:
        raise ValueError(f"orc_huffman status {st}")
    return lengths


def convert_lengths_to_encode_table(lengths, n: int, max_symbol_value: int | None = None, elen=None, evalue=None):
    """Returns (encode_length_table, encode_value_table, status).  Arrays are sized like `lengths`; pass
    pre-filled `elen`/`evalue` to observe the reference's skip-the-last-slot clearing quirk."""
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    msv = ln.size - 1 if max_symbol_value is None else max_symbol_value
    el = np.zeros(ln.size, dtype=np.int32) if elen is None else np.ascontiguousarray(elen, dtype=np.int32)
    ev = np.zeros(ln.size, dtype=np.uint32) if evalue is None else np.ascontiguousarray(evalue, dtype=np.uint32)
    st = lib().orc_convert_lengths_to_encode_table(msv, _p(ln, C.c_int), n, _p(el, C.c_int), _p(ev, C.c_uint))
    return el, ev, st


def bits_per_digit(n: int) -> int:
    return lib().orc_bits_per_digit(n)


def field_values(elen, evalue, n: int, bits: int) -> np.ndarray:
    """The canonical code values -- base-n numerals, as convert_lengths_to_encode_table() n_ary_huffman.c:1382-1612 assigns them --
    rewritten with one `bits`-wide field per digit, most significant digit first, for pack(.., bpd = bits)."""
    el = np.asarray(elen, dtype=np.int64)
    ev = np.asarray(evalue, dtype=np.uint64)
    out = np.zeros(el.size, dtype=np.uint32)
    for s in range(el.size):
        v, r = int(ev[s]), 0
        for k in range(int(el[s])):
            r |= (v % n) << (bits * k)
            v //= n
        out[s] = r
    return out


def nibble_values(elen, evalue, n: int) -> np.ndarray:
    """Radices 5 .. 15: one nibble per digit."""
    return field_values(elen, evalue, n, 4)


def unpack_nibble_digits(payload, bit_start: int, nbits: int, elen, evalue, n: int) -> np.ndarray:
    """Pure-Python reader of a nibble-per-digit payload (small cases): digits accumulate into a base-n numeral until (length,
    value) names a symbol.  Independent of every table the library builds."""
    el = np.asarray(elen, dtype=np.int64)
    ev = np.asarray(evalue, dtype=np.uint64)
    code = {(int(el[s]), int(ev[s])): s for s in range(min(el.size, 256)) if el[s] > 0}
    p = np.asarray(payload, dtype=np.uint8)
    out, length, value = [], 0, 0
    assert bit_start % 4 == 0 and nbits % 4 == 0
    for i in range(bit_start // 4, (bit_start + nbits) // 4):
        d = (int(p[i >> 1]) >> (4 if i % 2 == 0 else 0)) & 15
        assert d < n, "not a digit of this radix"
        length, value = length + 1, value * n + d
        if (length, value) in code:
            out.append(code[(length, value)])
            length, value = 0, 0
    assert length == 0, "the stream ends inside a code"
    return np.array(out, dtype=np.uint8)


def pack(data, elen, evalue, bpd: int, bit_phase: int = 0):
    """Returns (payload bytes, total_bits).  Raises on status != 0."""
    d = _u8(data)
    el = np.ascontiguousarray(elen, dtype=np.int32)
    ev = np.ascontiguousarray(evalue, dtype=np.uint32)
    cap = d.size * 4 + 16
    out = np.zeros(cap, dtype=np.uint8)
    bits = C.c_uint64(0)
    st = lib().orc_pack(d.ctypes.data, d.size, _p(el, C.c_int), _p(ev, C.c_uint), bpd, bit_phase,
                        out.ctypes.data, cap, C.byref(bits))
    if st != ORC_OK:
        raise ValueError(f"orc_pack status {st}")
    nbytes = (bits.value + bit_phase + 7) // 8 if bits.value else 0
    return out[:nbytes].copy(), bits.value


def pack_trits(data, elen, evalue):
    """n = 3 payload: returns (payload bytes, total_trits)."""
    d = _u8(data)
    el = np.ascontiguousarray(elen, dtype=np.int32)
    ev = np.ascontiguousarray(evalue, dtype=np.uint32)
    cap = d.size * 4 + 16
    out = np.zeros(cap, dtype=np.uint8)
    trits = C.c_uint64(0)
    st = lib().orc_pack_trits(d.ctypes.data, d.size, _p(el, C.c_int), _p(ev, C.c_uint), out.ctypes.data, cap, C.byref(trits))
    if st != ORC_OK:
        raise ValueError(f"orc_pack_trits status {st}")
    return out[: (trits.value + 4) // 5].copy(), trits.value


def unpack_trits(packed, total_trits: int, lengths, n_out: int) -> np.ndarray:
    b = _u8(packed)
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    out = np.zeros(max(n_out, 1), dtype=np.uint8)
    nd = C.c_size_t(0)
    st = lib().orc_unpack_trits(b.ctypes.data, total_trits, ln.size - 1, _p(ln, C.c_int), out.ctypes.data, n_out, C.byref(nd))
    if st != ORC_OK:
        raise ValueError(f"orc_unpack_trits status {st}")
    return out[: nd.value]


def pack_mt(data, elen, evalue, bpd: int, bit_phase: int = 0, threads: int = 1, block_symbols: int = 1 << 16,
            out: np.ndarray | None = None):
    d = _u8(data)
    el = np.ascontiguousarray(elen, dtype=np.int32)
    ev = np.ascontiguousarray(evalue, dtype=np.uint32)
    if out is None:
        out = np.empty(d.size * 4 + 16, dtype=np.uint8)
    nblocks = (d.size + block_symbols - 1) // block_symbols
    offs = np.zeros(nblocks + 1, dtype=np.uint64)
    bits = C.c_uint64(0)
    st = lib().orc_pack_mt(d.ctypes.data, d.size, _p(el, C.c_int), _p(ev, C.c_uint), bpd, bit_phase,
                           out.ctypes.data, out.size, C.byref(bits), threads, _p(offs, C.c_uint64), block_symbols)
    if st != ORC_OK:
        raise ValueError(f"orc_pack_mt status {st}")
    nbytes = (bits.value + bit_phase + 7) // 8 if bits.value else 0
    return out[:nbytes], bits.value, offs


def unpack(bits, bit_start: int, nbits: int, lengths, n: int, n_out: int, return_status: bool = False):
    b = _u8(bits)
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    out = np.zeros(max(n_out, 1), dtype=np.uint8)
    nd = C.c_size_t(0)
    st = lib().orc_unpack(b.ctypes.data, bit_start, nbits, ln.size - 1, _p(ln, C.c_int), n, out.ctypes.data,
                          n_out, C.byref(nd))
    if return_status:
        return out[:nd.value].copy(), st
    if st != ORC_OK:
        raise ValueError(f"orc_unpack status {st}")
    return out[:nd.value].copy()


def unpack_mt(bits, bit_start: int, lengths, n: int, n_out: int, block_offsets, block_symbols: int, threads: int,
              out: np.ndarray | None = None):
    b = _u8(bits)
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    offs = np.ascontiguousarray(block_offsets, dtype=np.uint64)
    if out is None:
        out = np.empty(max(n_out, 1), dtype=np.uint8)
    st = lib().orc_unpack_mt(b.ctypes.data, bit_start, ln.size - 1, _p(ln, C.c_int), n, out.ctypes.data, n_out,
                             _p(offs, C.c_uint64), block_symbols, threads)
    if st != ORC_OK:
        raise ValueError(f"orc_unpack_mt status {st}")
    return out[:n_out]


def nybble_pack(sym, threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    s = _u8(sym)
    if out is None:
        out = np.zeros((s.size + 1) // 2, dtype=np.uint8)
    if threads > 1:
        lib().orc_nybble_pack_mt(s.ctypes.data, s.size, out.ctypes.data, threads)
    else:
        lib().orc_nybble_pack(s.ctypes.data, s.size, out.ctypes.data)
    return out


def nybble_unpack(packed, n_sym: int, threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    p = _u8(packed)
    if out is None:
        out = np.zeros(n_sym, dtype=np.uint8)
    if threads > 1:
        lib().orc_nybble_unpack_mt(p.ctypes.data, n_sym, out.ctypes.data, threads)
    else:
        lib().orc_nybble_unpack(p.ctypes.data, n_sym, out.ctypes.data)
    return out


def nybble_static_compress(src: bytes) -> bytes:
    s = _u8(src)
    out = np.zeros(s.size * 2 + 8, dtype=np.uint8)
    n = lib().orc_nybble_static_compress(s.ctypes.data, s.size, out.ctypes.data)
    return out[:n].tobytes()


def nybble_static_decompress(src: bytes) -> bytes:
    s = _u8(src)
    out = np.zeros(s.size * 2 + 8, dtype=np.uint8)
    n = lib().orc_nybble_static_decompress(s.ctypes.data, s.size, out.ctypes.data)
    return out[:n].tobytes()


def base64url_pack(bits, nbits: int) -> bytes:
    b = _u8(bits)
    out = np.zeros((nbits + 5) // 6 + 8, dtype=np.uint8)
    n = lib().orc_base64url_pack(b.ctypes.data, nbits, out.ctypes.data)
    return out[:n].tobytes()


def base64url_unpack(chars, nbits: int) -> np.ndarray:
    c = _u8(chars)
    out = np.zeros((nbits + 7) // 8 + 8, dtype=np.uint8)
    st = lib().orc_base64url_unpack(c.ctypes.data, nbits, out.ctypes.data)
    if st != ORC_OK:
        raise ValueError(f"orc_base64url_unpack status {st}")
    return out[: (nbits + 7) // 8]


def nybble_adaptive_compress(src: bytes) -> bytes:
    s = _u8(src)
    out = np.zeros(s.size * 2 + 8, dtype=np.uint8)
    n = lib().orc_nybble_adaptive_compress(s.ctypes.data, s.size, out.ctypes.data)
    return out[:n].tobytes()


def nybble_adaptive_decompress(src: bytes) -> bytes:
    s = _u8(src)
    out = np.zeros(s.size * 2 + 8, dtype=np.uint8)
    n = lib().orc_nybble_adaptive_decompress(s.ctypes.data, s.size, out.ctypes.data)
    return out[:n].tobytes()


def build_tables(hist, n: int):
    """hist[259] -> (lengths, elen, evalue, status) through the restated huffman() + convert...()."""
    lengths = huffman(hist, n)
    el, ev, st = convert_lengths_to_encode_table(lengths, n)
    return lengths, el, ev, st


# ----------------------------------------------------------------------------- unmodified reference (oracle/_ref)

def have_ref() -> bool:
    return os.path.exists(REF_HUFF_PATH) and os.path.exists(REF_NYBBLE_PATH)


_ref_huff = None
_ref_nyb = None


def ref_huff() -> C.CDLL:
    global _ref_huff
    if _ref_huff is None:
        L = C.CDLL(REF_HUFF_PATH)
        i32p, u32p = C.POINTER(C.c_int), C.POINTER(C.c_uint)
        L.ref_histogram.argtypes = [C.c_char_p, C.c_int, i32p]
        L.ref_histogram.restype = None
        L.ref_huffman.argtypes = [C.c_int, i32p, C.c_int, i32p]
        L.ref_huffman.restype = None
        L.ref_convert_lengths_to_encode_table.argtypes = [C.c_int, i32p, C.c_int, i32p, u32p]
        L.ref_convert_lengths_to_encode_table.restype = None
        L.ref_run_tests.restype = None
        if hasattr(L, "ref_compress_block"):
            L.ref_compress_block.argtypes = [C.c_int, i32p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
            L.ref_compress_block.restype = None
        L.ref_silence_begin.restype = None
        L.ref_silence_end.restype = None
        _ref_huff = L
    return _ref_huff


def ref_nybble() -> C.CDLL:
    global _ref_nyb
    if _ref_nyb is None:
        L = C.CDLL(REF_NYBBLE_PATH)
        L.ref_write_nybble.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.ref_write_nybble.restype = None
        L.ref_compress_bytestring.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.ref_compress_bytestring.restype = None
        L.ref_decompress_bytestring.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.ref_decompress_bytestring.restype = None
        L.ref_nybble_selftest.restype = C.c_int
        _ref_nyb = L
    return _ref_nyb


def ref_histogram(text: bytes, max_symbol_value: int = MAX_SYMBOL_VALUE) -> np.ndarray:
    """Unmodified histogram() (n_ary_huffman.c:461).  `text` must be NUL-free; bytes > 126 make it print."""
    h = np.full(max_symbol_value + 1, 0xBEEF, dtype=np.int32)
    ref_huff().ref_histogram(text, max_symbol_value, _p(h, C.c_int))
    return h


def ref_compress_block(text: bytes, lengths, n: int, bufsize: int = 65000, fill: int = 0xFF) -> bytes:
    """Unmodified static compress() (n_ary_huffman.c:1688): the whole output buffer, pre-filled with `fill`."""
    ln = np.ascontiguousarray(lengths, dtype=np.int32).copy()
    src = C.create_string_buffer(text, bufsize + 1)
    dst = C.create_string_buffer(bytes([fill]) * (bufsize + 1), bufsize + 1)
    ref_huff().ref_compress_block(MAX_SYMBOL_VALUE, _p(ln, C.c_int), n, bufsize, len(text), src, dst)
    return dst.raw


def ref_huffman(freqs, n: int, max_leaf_value: int | None = None) -> np.ndarray:
    f = np.ascontiguousarray(freqs, dtype=np.int32)
    mlv = f.size - 1 if max_leaf_value is None else max_leaf_value
    lengths = np.zeros(mlv + 1, dtype=np.int32)
    ref_huff().ref_huffman(mlv, _p(f, C.c_int), n, _p(lengths, C.c_int))
    return lengths


def ref_convert_lengths_to_encode_table(lengths, n: int, max_symbol_value: int | None = None, elen=None, evalue=None):
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    msv = ln.size - 1 if max_symbol_value is None else max_symbol_value
    el = np.zeros(ln.size, dtype=np.int32) if elen is None else np.ascontiguousarray(elen, dtype=np.int32)
    ev = np.zeros(ln.size, dtype=np.uint32) if evalue is None else np.ascontiguousarray(evalue, dtype=np.uint32)
    ref_huff().ref_convert_lengths_to_encode_table(msv, _p(ln, C.c_int), n, _p(el, C.c_int), _p(ev, C.c_uint))
    return el, ev


def ref_compress_bytestring(text: bytes, modify: bool) -> bytes:
    out = C.create_string_buffer(len(text) * 2 + 16)
    ref_nybble().ref_compress_bytestring(text, out, int(modify))
    return out.value


def ref_decompress_bytestring(comp: bytes, modify: bool) -> bytes:
    out = C.create_string_buffer(len(comp) * 2 + 16)
    ref_nybble().ref_decompress_bytestring(comp, out, int(modify))
    return out.value


def ref_write_nybble_stream(sym) -> np.ndarray:
    """Pack a symbol stream by calling the unmodified write_nybble() once per symbol."""
    s = _u8(sym)
    out = np.zeros((s.size + 1) // 2, dtype=np.uint8)
    L = ref_nybble()
    base = out.ctypes.data
    for i, v in enumerate(s.tolist()):
        L.ref_write_nybble(int(v), base + (i >> 1), i & 1)
    return out
