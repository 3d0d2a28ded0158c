"""Loader for libdc_b200.so (the C-ABI of include/dc_b200.h).

The shared library is built IN-TREE by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU
fallback: if the library is missing, or no CUDA device is usable, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdc_b200.so")
REFAPI_PATH = os.path.join(HERE, "libdc_b200_refapi.so")

DC_NSLOTS = 259
DC_MAX_SYMBOL_VALUE = 258
DC_LUT_BITS = 12
DC_LUT_ENTRIES = 6564   # max(2^12, 3^8) rounded up: the multi-symbol decode tables (radix 3 indexes them by 8 trits)

DC_OK, DC_ERR_ARG, DC_ERR_CUDA, DC_ERR_CODE_TOO_LONG = 0, -1, -2, -3
DC_ERR_CAPACITY, DC_ERR_CORRUPT, DC_ERR_SYMBOL, DC_ERR_RADIX = -4, -5, -6, -7


class DcError(RuntimeError):
    def __init__(self, status: int, where: str = ""):
        self.status = status
        msg = status_string(status) if _lib is not None else str(status)
        super().__init__(f"{where}: dc status {status} ({msg})" if where else f"dc status {status} ({msg})")


class HuffTableStruct(C.Structure):
    """Mirror of ``struct dc_huff_table`` (include/dc_b200.h)."""
    _fields_ = [
        ("n_ary", C.c_int32), ("bits_per_digit", C.c_int32), ("max_symbol_value", C.c_int32),
        ("nonzero_symbols", C.c_int32), ("dummy_nodes", C.c_int32), ("min_len", C.c_int32), ("max_len", C.c_int32),
        ("max_bits", C.c_int32), ("status", C.c_int32), ("packed_radix", C.c_int32),
        ("total_symbols", C.c_uint64), ("total_bits", C.c_uint64),
        ("lengths", C.c_int32 * (DC_NSLOTS + 1)), ("values", C.c_uint32 * (DC_NSLOTS + 1)),
        ("enc", C.c_uint32 * 256), ("enc64", C.c_uint64 * 256),
        ("first_code", C.c_uint32 * 32), ("len_count", C.c_uint32 * 32), ("len_offset", C.c_uint32 * 32),
        ("sorted", C.c_uint16 * (DC_NSLOTS + 1)), ("lut", C.c_uint16 * (1 << DC_LUT_BITS)),
        ("lut_count", C.c_uint32 * DC_LUT_ENTRIES), ("lut_pair", C.c_uint32 * DC_LUT_ENTRIES),
        ("lut2", C.c_uint16 * (257 * 16)), ("lut2_used", C.c_int32), ("fsm_states", C.c_int32), ("lut14", C.c_uint16 * (1 << 14)),
    ]


TABLE_BYTES = C.sizeof(HuffTableStruct)


class HuffIndexInfo(C.Structure):
    """dc_huff_index_info: what the host keeps with a stream's index (include/dc_b200.h)."""
    _fields_ = [("magic", C.c_uint64), ("bit_start", C.c_uint64), ("nbits", C.c_uint64), ("n_symbols", C.c_uint64),
                ("mode", C.c_uint32), ("start_token", C.c_uint32), ("reserved", C.c_uint32 * 2)]


class ShardSummaryStruct(C.Structure):
    """Mirror of ``struct dc_shard_summary`` (include/dc_b200.h)."""
    _fields_ = [("symbols", C.c_uint64), ("exit", C.c_uint32), ("resync", C.c_int32), ("assumed_start", C.c_uint32),
                ("reserved", C.c_uint32)]

# every symbol include/dc_b200.h declares: (name, restype, argtypes)
_vp, _sz, _u64, _i, _u = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_uint
_ip, _up, _u64p = C.POINTER(C.c_int), C.POINTER(C.c_uint), C.POINTER(C.c_uint64)
SYMBOLS = [
    ("dc_version", C.c_char_p, []),
    ("dc_status_string", C.c_char_p, [_i]),
    ("dc_device_count", _i, []),
    ("dc_launch_count", _u64, []),
    ("dc_profile_enable", _i, [_i]),
    ("dc_profile_reset", _i, []),
    ("dc_profile_kernel", _i, [_i, C.POINTER(C.c_double), _u64p]),
    ("dc_profile_kernel_name", C.c_char_p, [_i]),
    ("dc_histogram_u8", _i, [_vp, _sz, _vp, _vp]),
    ("dc_huff_build", _i, [_vp, _i, _vp, _vp]),
    ("dc_huff_table_from_lengths", _i, [_vp, _i, _vp, _vp]),
    ("dc_huff_table_download", _i, [_vp, _vp, _vp]),
    ("dc_huff_table_forget", _i, [_vp]),
    ("dc_huff_index_bytes", _sz, [_u64, _u64]),
    ("dc_huff_index_build", _i, [_vp, _u64, _u64, _vp, _u64, _vp, _sz, _vp, _vp, _sz, _vp]),
    ("dc_huff_decode_indexed", _i, [_vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _vp, _sz, _vp]),
    ("dc_huff_bits_for_hist", _i, [_vp, _vp, _vp, _vp]),
    ("dc_huff_encode_workspace_bytes", _sz, [_sz]),
    ("dc_huff_encode", _i, [_vp, _sz, _vp, _vp, _sz, _u, _vp, _vp, _vp, _sz, _vp]),
    ("dc_histogram_u8_runs", _i, [_vp, _sz, _vp, _vp, _sz, _vp]),
    ("dc_huff_encode_planned", _i, [_vp, _sz, _vp, _vp, _sz, _u, _vp, _vp, _vp, _sz, _vp]),
    ("dc_huff_decode_workspace_bytes", _sz, [_u64, _u64]),
    ("dc_huff_decode", _i, [_vp, _u64, _u64, _vp, _vp, _sz, _vp, _vp, _sz, _vp]),
    ("dc_huff_decode_shard_sync", _i, [_vp, _i, _u, _u64, _u64, _vp, _vp, _vp, _sz, _vp]),
    ("dc_huff_decode_shard_write", _i, [_vp, _i, _u64, _u64, _vp, _vp, _sz, _vp, _vp, _sz, _vp]),
    ("dc_shard_unique_id", _i, [_vp]),
    ("dc_shard_comm_create", _i, [_vp, _i, _i, C.POINTER(C.c_void_p)]),
    ("dc_shard_comm_from_nccl", _i, [_vp, _i, _i, C.POINTER(C.c_void_p)]),
    ("dc_shard_comm_destroy", _i, [_vp]),
    ("dc_shard_comm_rank", _i, [_vp]),
    ("dc_shard_comm_world", _i, [_vp]),
    ("dc_shard_huff_encode_workspace_bytes", _sz, [_sz, _i]),
    ("dc_shard_huff_encode", _i, [_vp, _vp, _sz, _i, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    ("dc_shard_huff_encode_info", _i, [_vp, _sz, _i, _u64p, _u64p, _u64p, _vp]),
    ("dc_shard_huff_gather", _i, [_vp, _i, _vp, _vp, _sz, _vp, _sz, _vp]),
    ("dc_shard_huff_decode_workspace_bytes", _sz, [_sz, _i]),
    ("dc_shard_huff_decode_stream", _i, [_vp, _vp, _sz, _u64, _vp, _vp, _sz, _u64p, _u64p, _u64p, _vp, _vp, _sz, _vp]),
    ("dc_shard_nybble_range", _i, [_u64, _i, _i, _u64p, _u64p]),
    ("dc_shard_nybble_pack", _i, [_u64, _i, _i, _vp, _vp, _vp, _vp]),
    ("dc_shard_nybble_unpack", _i, [_u64, _i, _i, _vp, _vp, _vp]),
    ("dc_nybble_pack", _i, [_vp, _sz, _vp, _vp, _vp]),
    ("dc_nybble_unpack", _i, [_vp, _sz, _vp, _vp]),
    ("dc_trit_pack", _i, [_vp, _u64, _vp, _vp, _vp]),
    ("dc_trit_unpack", _i, [_vp, _u64, _vp, _vp, _vp]),
    ("dc_base64url_pack", _i, [_vp, _u64, _vp, _vp]),
    ("dc_base64url_unpack", _i, [_vp, _u64, _vp, _vp, _vp]),
    ("dc_nybble_text_workspace_bytes", _sz, [_sz]),
    ("dc_nybble_text_compress", _i, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    ("dc_nybble_text_decompress", _i, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    ("dc_nybble_adaptive_workspace_bytes", _sz, [_sz]),
    ("dc_nybble_adaptive_compress", _i, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    ("dc_nybble_adaptive_decompress", _i, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    ("dc_nybble_text_compress_batch", _i, [_vp, _vp, _sz, _i, _vp, _vp, _vp, _vp, _vp]),
    ("dc_nybble_text_decompress_batch", _i, [_vp, _vp, _sz, _i, _vp, _vp, _vp, _vp, _vp]),
    ("dc_synth_fill", _i, [_vp, _sz, _u64, _vp, _i, _i, _vp]),
    ("dc_host_histogram", _i, [C.c_char_p, _i, _ip]),
    ("dc_host_histogram_u8", _i, [_vp, _sz, _u64p]),
    ("dc_host_huffman", _i, [_i, _ip, _i, _ip]),
    ("dc_host_huffman_u64", _i, [_i, _u64p, _i, _ip]),
    ("dc_host_convert_lengths_to_encode_table", _i, [_i, _ip, _i, _ip, _up]),
    ("dc_host_represent_items_with_codes", _i, [_i, _ip, _i, _i, _i, _vp, _i, _vp, _u64p]),
    ("dc_host_huff_compress", C.c_longlong, [_vp, _sz, _i, _vp, _sz, _ip, _u64p]),
    ("dc_host_huff_decompress", _i, [_vp, _u64, _ip, _i, _vp, _sz]),
    ("dc_host_compress_bytestring", C.c_longlong, [C.c_char_p, _vp, _i]),
    ("dc_host_decompress_bytestring", C.c_longlong, [C.c_char_p, _vp, _i]),
    ("dc_host_nybble_pack", _i, [_vp, _sz, _vp]),
    ("dc_host_nybble_unpack", _i, [_vp, _sz, _vp]),
]
# bench/test hooks that are exported but not part of the public header
EXTRA_SYMBOLS = [
    ("dc_histogram_u8_variant", _i, [_vp, _sz, _vp, _i, _vp]),
    ("dc_debug_decode_mode", _i, [_i]),
    ("dc_debug_shard_peer_active", _i, [_vp]),
]

_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises if it has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "data_compression_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS + EXTRA_SYMBOLS:
            fn = getattr(L, name)  # AttributeError if the .so does not export what the header declares
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def status_string(status: int) -> str:
    return lib().dc_status_string(status).decode()


def check(status: int, where: str = "") -> None:
    if status != DC_OK:
        raise DcError(status, where)
