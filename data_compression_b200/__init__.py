"""data_compression_b200 -- B200-native (sm_100a) n-ary Huffman encode/decode and nybble pack/unpack.

A from-scratch implementation of the data-parallel hot path of carycode/data_compression
(n_ary_huffman.c, nybble_compression.c): hand-written CUDA kernels behind the C-ABI of
include/dc_b200.h, plus this thin host-side mirror.  PyTorch is used only for device memory, streams and
torch.distributed; all compute happens in libdc_b200.so.  There is no CPU fallback.
"""
from ._lib import (DC_ERR_ARG, DC_ERR_CAPACITY, DC_ERR_CODE_TOO_LONG, DC_ERR_CORRUPT, DC_ERR_CUDA, DC_ERR_RADIX,
                   DC_ERR_SYMBOL, DC_NSLOTS, DC_OK, DcError, lib, status_string)
from .api import (HuffTable, ShardDecoder, encode_workspace, histogram, histogram_runs, huff_bits_for_hist, huff_build, huff_compress, huff_decode, huff_decode_indexed, huff_decompress,
                  huff_encode, huff_index_build, huff_table_from_lengths, launch_count, nybble_pack, nybble_text_compress, nybble_text_decompress, nybble_adaptive_compress, nybble_adaptive_decompress, nybble_text_compress_batch, nybble_text_decompress_batch,
                  nybble_unpack, synth_fill, trit_pack, trit_unpack, base64url_pack, base64url_unpack)
from . import hostapi, synth

__all__ = [
    "DC_OK", "DC_ERR_ARG", "DC_ERR_CUDA", "DC_ERR_CODE_TOO_LONG", "DC_ERR_CAPACITY", "DC_ERR_CORRUPT", "DC_ERR_SYMBOL",
    "DC_ERR_RADIX", "DC_NSLOTS", "DcError", "lib", "status_string", "HuffTable", "ShardDecoder", "encode_workspace", "histogram", "histogram_runs", "huff_build",
    "huff_table_from_lengths", "huff_bits_for_hist", "huff_encode", "huff_decode", "huff_index_build", "huff_decode_indexed", "huff_compress", "huff_decompress",
    "nybble_pack", "nybble_unpack", "nybble_text_compress", "nybble_text_decompress", "nybble_adaptive_compress", "nybble_adaptive_decompress", "nybble_text_compress_batch", "nybble_text_decompress_batch", "synth_fill", "trit_pack", "trit_unpack", "base64url_pack", "base64url_unpack", "launch_count", "hostapi", "synth",
]
