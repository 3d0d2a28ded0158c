"""Host-buffer mirror of the reference's C functions, routed through the dc_host_* C-ABI entry points
(H2D copy, device kernels, D2H copy, synchronise).  Argument meaning follows the reference:

    histogram(text, max_symbol_value)                         n_ary_huffman.c:461
    huffman(max_leaf_value, freqs, compressed_symbols)        n_ary_huffman.c:1161
    convert_lengths_to_encode_table(msv, lengths, n)          n_ary_huffman.c:1382
    represent_items_with_codes(...)                           n_ary_huffman.c:1621
    nybble pack / unpack (write_nybble stream form)           nybble_compression.c:1091, :767

Errors the reference would assert() on come back as DcError instead of aborting.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import DC_NSLOTS, DcError, check, lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def histogram(text: bytes, max_symbol_value: int = 258) -> np.ndarray:
    """NUL-terminated semantics of the reference: counting stops at the first 0x00."""
    h = np.full(max_symbol_value + 1, 0xBEEF, dtype=np.int32)
    check(lib().dc_host_histogram(text, max_symbol_value, _p(h, C.c_int)), "dc_host_histogram")
    return h


def histogram_u8(data) -> np.ndarray:
    d = _u8(data)
    h = np.zeros(DC_NSLOTS, dtype=np.uint64)
    check(lib().dc_host_histogram_u8(d.ctypes.data, d.size, _p(h, C.c_uint64)), "dc_host_histogram_u8")
    return h


def huffman(freqs, compressed_symbols: int, max_leaf_value: int | None = None) -> np.ndarray:
    f = np.ascontiguousarray(freqs)
    mlv = f.size - 1 if max_leaf_value is None else max_leaf_value
    lengths = np.zeros(mlv + 1, dtype=np.int32)
    if f.dtype == np.uint64 or f.dtype == np.int64:
        f = f.astype(np.uint64)
        st = lib().dc_host_huffman_u64(mlv, _p(f, C.c_uint64), compressed_symbols, _p(lengths, C.c_int))
    else:
        f = f.astype(np.int32)
        st = lib().dc_host_huffman(mlv, _p(f, C.c_int), compressed_symbols, _p(lengths, C.c_int))
    check(st, "dc_host_huffman")
    return lengths


def convert_lengths_to_encode_table(lengths, compressed_symbols: int, max_symbol_value: int | None = None,
                                    elen=None, evalue=None):
    """Returns (encode_length_table, encode_value_table, status); status is DC_ERR_CODE_TOO_LONG where the
    reference would assert (:1414)."""
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    msv = ln.size - 1 if max_symbol_value is None else max_symbol_value
    el = np.zeros(ln.size, dtype=np.int32) if elen is None else np.ascontiguousarray(elen, dtype=np.int32)
    ev = np.zeros(ln.size, dtype=np.uint32) if evalue is None else np.ascontiguousarray(evalue, dtype=np.uint32)
    st = lib().dc_host_convert_lengths_to_encode_table(msv, _p(ln, C.c_int), compressed_symbols, _p(el, C.c_int),
                                                       _p(ev, C.c_uint))
    if st not in (0, -3):
        raise DcError(st, "dc_host_convert_lengths_to_encode_table")
    return el, ev, st


def represent_items_with_codes(lengths, compressed_symbols: int, text: bytes, bufsize: int | None = None, start: int = 0):
    """Returns (compressed_text bytes incl. the `start` prefix untouched, bytes_written, total_bits)."""
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    t = _u8(text)
    bufsize = max(len(t) * 4 + 64, 64) if bufsize is None else bufsize
    out = np.zeros(bufsize + 1, dtype=np.uint8)
    bits = C.c_uint64(0)
    rc = lib().dc_host_represent_items_with_codes(ln.size - 1, _p(ln, C.c_int), compressed_symbols, bufsize, t.size,
                                                  t.ctypes.data, start, out.ctypes.data, C.byref(bits))
    if rc < 0:
        raise DcError(rc, "dc_host_represent_items_with_codes")
    return out, rc, bits.value


def huff_compress(data, compressed_symbols: int, out: np.ndarray | None = None):
    """Returns (payload ndarray view, total_bits, lengths[259])."""
    d = _u8(data)
    if out is None:
        out = np.empty((2 * d.size + d.size // 2 if 5 <= compressed_symbols < 16 else d.size + d.size // 4) + 64, dtype=np.uint8)
    lengths = np.zeros(DC_NSLOTS, dtype=np.int32)
    bits = C.c_uint64(0)
    rc = lib().dc_host_huff_compress(d.ctypes.data, d.size, compressed_symbols, out.ctypes.data, out.size,
                                     _p(lengths, C.c_int), C.byref(bits))
    if rc < 0:
        raise DcError(int(rc), "dc_host_huff_compress")
    return out[:rc], bits.value, lengths


def huff_decompress(payload, total_bits: int, lengths, compressed_symbols: int, n_out: int,
                    out: np.ndarray | None = None) -> np.ndarray:
    p = _u8(payload)
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    if out is None:
        out = np.empty(max(n_out, 1), dtype=np.uint8)
    check(lib().dc_host_huff_decompress(p.ctypes.data, total_bits, _p(ln, C.c_int), compressed_symbols,
                                        out.ctypes.data, n_out), "dc_host_huff_decompress")
    return out[:n_out]


def nybble_pack(sym, out: np.ndarray | None = None) -> np.ndarray:
    s = _u8(sym)
    if out is None:
        out = np.zeros((s.size + 1) // 2, dtype=np.uint8)
    check(lib().dc_host_nybble_pack(s.ctypes.data, s.size, out.ctypes.data), "dc_host_nybble_pack")
    return out


def nybble_unpack(packed, n_sym: int, out: np.ndarray | None = None) -> np.ndarray:
    p = _u8(packed)
    if out is None:
        out = np.zeros(n_sym, dtype=np.uint8)
    check(lib().dc_host_nybble_unpack(p.ctypes.data, n_sym, out.ctypes.data), "dc_host_nybble_unpack")
    return out


def compress_bytestring(source: bytes, modify: bool = False) -> bytes:
    """compress_bytestring(source, dest, modify) nybble_compression.c:887 on the GPU: the static table, or with
    modify=True the 16 adaptive move-to-front contexts (nybble_compress() :1134)."""
    dest = C.create_string_buffer(len(source) + 2)
    n = lib().dc_host_compress_bytestring(bytes(source), C.addressof(dest), int(modify))
    if n < 0:
        raise DcError(int(n), "dc_host_compress_bytestring")
    return dest.raw[:n]


def decompress_bytestring(source: bytes, modify: bool = False) -> bytes:
    """decompress_bytestring(source, dest, modify) nybble_compression.c:734 (modify=True: nybble_decompress() :1117)."""
    dest = C.create_string_buffer(2 * len(source) + 2)
    n = lib().dc_host_decompress_bytestring(bytes(source), C.addressof(dest), int(modify))
    if n < 0:
        raise DcError(int(n), "dc_host_decompress_bytestring")
    return dest.raw[:n]
