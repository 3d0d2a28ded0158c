"""Device-side operator API: torch CUDA tensors in, torch CUDA tensors out, all compute in libdc_b200.so.

Function names follow the reference's (histogram, huffman -> huff_build, represent_items_with_codes ->
huff_encode, nybble pack/unpack); every call is stream-ordered on torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DC_NSLOTS, TABLE_BYTES, HuffTableStruct, check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: data_compression_b200 has no CPU path")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def launch_count() -> int:
    return int(lib().dc_launch_count())


class HuffTable:
    """A device-resident ``dc_huff_table`` (lengths, canonical values, encode entries, decode LUT)."""

    def __init__(self, device=None):
        self.buf = torch.empty(TABLE_BYTES, dtype=torch.uint8, device=device or torch.cuda.current_device())
        self.n_ary = None   # set by huff_build / huff_table_from_lengths (sizes the default payload buffer)

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()

    def __del__(self):
        # the library keeps the header of every table it built, keyed by address: drop it before the allocator reuses the memory
        try:
            lib().dc_huff_table_forget(self.buf.data_ptr())
        except Exception:
            pass

    def download(self) -> HuffTableStruct:
        """Blocking copy to the host."""
        h = HuffTableStruct()
        check(lib().dc_huff_table_download(self.ptr, C.addressof(h), _stream()), "dc_huff_table_download")
        return h

    def lengths(self) -> np.ndarray:
        return np.array(self.download().lengths[:DC_NSLOTS], dtype=np.int32)


def histogram(data: torch.Tensor, out: torch.Tensor | None = None, variant: int | None = None) -> torch.Tensor:
    """259 x int64 byte counts of `data` (uint8, CUDA).  Replaces histogram() n_ary_huffman.c:461."""
    _need_cuda(data, "data")
    if out is None:
        out = torch.empty(DC_NSLOTS, dtype=torch.int64, device=data.device)
    if variant is None:
        st = lib().dc_histogram_u8(data.data_ptr(), data.numel(), out.data_ptr(), _stream())
    else:
        st = lib().dc_histogram_u8_variant(data.data_ptr(), data.numel(), out.data_ptr(), variant, _stream())
    check(st, "dc_histogram_u8")
    return out


def encode_workspace(n: int, device) -> torch.Tensor:
    """A workspace for huff_encode / histogram_runs on n input bytes."""
    return torch.empty(max(lib().dc_huff_encode_workspace_bytes(n), 16), dtype=torch.uint8, device=device)


def histogram_runs(data: torch.Tensor, workspace: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """histogram() that also leaves one small histogram per 32 KB run in the encode workspace, so that
    huff_encode(..., workspace=workspace, planned=True) knows every run's bit offset without a second data pass."""
    _need_cuda(data, "data")
    if out is None:
        out = torch.empty(DC_NSLOTS, dtype=torch.int64, device=data.device)
    check(lib().dc_histogram_u8_runs(data.data_ptr(), data.numel(), out.data_ptr(), workspace.data_ptr(), workspace.numel(),
                                     _stream()), "dc_histogram_u8_runs")
    return out


def huff_build(hist: torch.Tensor, n_ary: int, table: HuffTable | None = None) -> HuffTable:
    """hist (259 x int64, CUDA) -> code table.  Replaces huffman() :1161 + convert_lengths_to_encode_table() :1382."""
    _need_cuda(hist, "hist")
    if hist.numel() != DC_NSLOTS or hist.dtype != torch.int64:
        raise ValueError("hist must be 259 x int64")
    table = table or HuffTable(hist.device)
    check(lib().dc_huff_build(hist.data_ptr(), n_ary, table.ptr, _stream()), "dc_huff_build")
    table.n_ary = n_ary
    return table


def huff_table_from_lengths(lengths: torch.Tensor, n_ary: int, table: HuffTable | None = None) -> HuffTable:
    _need_cuda(lengths, "lengths")
    if lengths.numel() != DC_NSLOTS or lengths.dtype != torch.int32:
        raise ValueError("lengths must be 259 x int32")
    table = table or HuffTable(lengths.device)
    check(lib().dc_huff_table_from_lengths(lengths.data_ptr(), n_ary, table.ptr, _stream()), "dc_huff_table_from_lengths")
    table.n_ary = n_ary
    return table


def huff_bits_for_hist(hist: torch.Tensor, table: HuffTable, out: torch.Tensor | None = None) -> torch.Tensor:
    """1 x int64: bits a shard with LOCAL histogram `hist` emits under `table`."""
    _need_cuda(hist, "hist")
    if out is None:
        out = torch.empty(1, dtype=torch.int64, device=hist.device)
    check(lib().dc_huff_bits_for_hist(hist.data_ptr(), table.ptr, out.data_ptr(), _stream()), "dc_huff_bits_for_hist")
    return out


class EncodeResult:
    __slots__ = ("payload", "total_bits", "status")

    def __init__(self, payload, total_bits, status):
        self.payload, self.total_bits, self.status = payload, total_bits, status

    def bits(self) -> int:
        """Blocking: number of code bits emitted (raises on a device-side error)."""
        st = int(self.status.item())
        check(st, "dc_huff_encode")
        return int(self.total_bits.item())


def huff_encode(data: torch.Tensor, table: HuffTable, out: torch.Tensor | None = None, bit_phase: int = 0,
                workspace: torch.Tensor | None = None, planned: bool = False) -> EncodeResult:
    """Encode `data` (uint8, CUDA) with `table`.  Replaces represent_items_with_codes() :1621.

    Returns the payload buffer (capacity-sized; the first ceil((bit_phase+bits)/8) bytes are valid) together
    with 1-element device tensors for the bit count and the status; nothing blocks."""
    _need_cuda(data, "data")
    n = data.numel()
    if out is None:
        # a Huffman code of byte data never exceeds 8 bits per symbol by more than the skew allows: n + n / 4 covers every
        # power-of-two radix; radix 3 runs on 2 bits per trit (uniform bytes: 5.08 trits = 10.2 bits per symbol)
        # nibble-per-digit radices (5 .. 15): up to log_n(256) + 1 digits of 4 bits per symbol (n = 5: 17.8 bits)
        nib = table.n_ary is not None and 5 <= table.n_ary < 16
        cap = 2 * n + n // 2 + 64 if nib else n + n // 2 + 64 if table.n_ary == 3 else n + n // 4 + 64
        out = torch.empty(cap, dtype=torch.uint8, device=data.device)
    need = lib().dc_huff_encode_workspace_bytes(n)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=data.device)
    total_bits = torch.empty(1, dtype=torch.int64, device=data.device)
    status = torch.empty(1, dtype=torch.int32, device=data.device)
    fn = lib().dc_huff_encode_planned if planned else lib().dc_huff_encode   # planned: `workspace` comes from histogram_runs(data, ...)
    check(fn(data.data_ptr(), n, table.ptr, out.data_ptr(), out.numel(), bit_phase,
             total_bits.data_ptr(), status.data_ptr(), workspace.data_ptr(), workspace.numel(),
             _stream()), "dc_huff_encode")
    return EncodeResult(out, total_bits, status)


def huff_decode(bits: torch.Tensor, nbits: int, table: HuffTable, n_out: int, bit_start: int = 0,
                out: torch.Tensor | None = None, workspace: torch.Tensor | None = None,
                status: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Self-synchronising parallel decode of `nbits` code bits -> n_out symbols.  Returns (out, status)."""
    _need_cuda(bits, "bits")
    if out is None:
        out = torch.empty(max(n_out, 1), dtype=torch.uint8, device=bits.device)
    need = lib().dc_huff_decode_workspace_bytes(bit_start, nbits)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=bits.device)
    if status is None:
        status = torch.empty(1, dtype=torch.int32, device=bits.device)
    check(lib().dc_huff_decode(bits.data_ptr(), bit_start, nbits, table.ptr, out.data_ptr(), n_out, status.data_ptr(),
                               workspace.data_ptr(), workspace.numel(), _stream()), "dc_huff_decode")
    return out[:n_out], status


def huff_index_build(bits: torch.Tensor, nbits: int, table: HuffTable, n_symbols: int, bit_start: int = 0,
                     workspace: torch.Tensor | None = None):
    """Opt-in: the index of a finished stream (what the decoder's first pass would find), to be kept beside the payload.
    Blocking.  Returns (index tensor, HuffIndexInfo), or None if the stream does not self-synchronise (decode it blindly)."""
    from ._lib import HuffIndexInfo
    _need_cuda(bits, "bits")
    need = lib().dc_huff_decode_workspace_bytes(bit_start, nbits)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=bits.device)
    index = torch.empty(max(lib().dc_huff_index_bytes(bit_start, nbits), 16), dtype=torch.uint8, device=bits.device)
    info = HuffIndexInfo()
    rc = lib().dc_huff_index_build(bits.data_ptr(), bit_start, nbits, table.ptr, n_symbols, index.data_ptr(), index.numel(),
                                   C.addressof(info), workspace.data_ptr(), workspace.numel(), _stream())
    if rc == 1:
        return None
    check(rc, "dc_huff_index_build")
    return index, info


def huff_decode_indexed(bits: torch.Tensor, index: torch.Tensor, info, table: HuffTable, out: torch.Tensor | None = None,
                        workspace: torch.Tensor | None = None, status: torch.Tensor | None = None):
    """The decoder's write pass alone, from a stream's index (huff_index_build).  Returns (out, status); nothing blocks."""
    _need_cuda(bits, "bits")
    n_out = int(info.n_symbols)
    if out is None:
        out = torch.empty(max(n_out, 1), dtype=torch.uint8, device=bits.device)
    need = lib().dc_huff_decode_workspace_bytes(int(info.bit_start), int(info.nbits))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=bits.device)
    if status is None:
        status = torch.empty(1, dtype=torch.int32, device=bits.device)
    check(lib().dc_huff_decode_indexed(bits.data_ptr(), C.addressof(info), table.ptr, index.data_ptr(), index.numel(), out.data_ptr(),
                                       n_out, status.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream()),
          "dc_huff_decode_indexed")
    return out[:n_out], status


SHARD_ALIGN = 1024  # a longer stream is cut into shards at multiples of this many bytes


class ShardDecoder:
    """One contiguous byte range of a longer bitstream (dc_huff_decode_shard_sync / _write).

    `buf` holds [1024 bytes of the previous shard's tail | the shard | >= 16 bytes of the next shard]; the halo in front
    is ignored for the stream's first shard."""

    def __init__(self, buf: torch.Tensor, shard_bytes: int, shard_bits: int, stream_bits_left: int, table: HuffTable):
        _need_cuda(buf, "buf")
        final = stream_bits_left <= shard_bits
        tail = 16 if final else SHARD_ALIGN   # the kernels load whole 1 KB tiles one ahead: a non-final shard needs the next one's head
        if buf.data_ptr() % 16 or buf.numel() < SHARD_ALIGN + (shard_bytes + 15) // 16 * 16 + tail:
            raise ValueError("buf: 16-byte aligned; 1024-byte halo + shard (rounded up to 16 bytes) + 1024 bytes behind a "
                             "non-final shard (16 behind the last one)")
        self.buf, self.table = buf, table
        self.shard_bits, self.left = shard_bits, stream_bits_left
        self.ptr = buf.data_ptr() + SHARD_ALIGN
        need = lib().dc_huff_decode_workspace_bytes(0, shard_bits + 8 * SHARD_ALIGN)
        self.ws = torch.empty(max(need, 16), dtype=torch.uint8, device=buf.device)
        self.summary = torch.empty(3, dtype=torch.int64, device=buf.device)   # 24 bytes: dc_shard_summary
        self.has_halo = 0

    def sync(self, has_halo: bool, first_code_bit: int = 0) -> torch.Tensor:
        """Phase 1.  Returns the device-resident summary as int64[3] (symbols, exit | resync << 32, assumed | 0)."""
        self.has_halo = 1 if has_halo else 0
        check(lib().dc_huff_decode_shard_sync(self.ptr, self.has_halo, first_code_bit, self.shard_bits, self.left, self.table.ptr,
                                              self.summary.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _stream()),
              "dc_huff_decode_shard_sync")
        return self.summary

    @staticmethod
    def unpack(summary_i64) -> dict:
        s = [int(x) for x in summary_i64]
        return {"symbols": s[0], "exit": s[1] & 0xFFFFFFFF, "resync": (s[1] >> 32) & 0xFFFFFFFF, "assumed_start": s[2] & 0xFFFFFFFF}

    def write(self, n_out: int):
        out = torch.empty(max(n_out, 1), dtype=torch.uint8, device=self.buf.device)
        status = torch.empty(1, dtype=torch.int32, device=self.buf.device)
        check(lib().dc_huff_decode_shard_write(self.ptr, self.has_halo, self.shard_bits, self.left, self.table.ptr, out.data_ptr(),
                                               n_out, status.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _stream()),
              "dc_huff_decode_shard_write")
        return out[:n_out], status


def huff_compress(data: torch.Tensor, n_ary: int):
    """histogram -> table -> encode on one GPU.  Returns (payload tensor, total_bits int, HuffTable)."""
    ws = encode_workspace(data.numel(), data.device)
    hist = histogram_runs(data, ws) if data.data_ptr() % 16 == 0 else histogram(data)
    table = huff_build(hist, n_ary)
    res = huff_encode(data, table, workspace=ws, planned=data.data_ptr() % 16 == 0)
    nbits = res.bits()
    return res.payload[: (nbits + 7) // 8], nbits, table


def huff_decompress(payload: torch.Tensor, nbits: int, table: HuffTable, n_out: int) -> torch.Tensor:
    out, status = huff_decode(payload, nbits, table, n_out)
    check(int(status.item()), "dc_huff_decode")
    return out


def nybble_pack(sym: torch.Tensor, out: torch.Tensor | None = None, status: torch.Tensor | None = None):
    """One 4-bit symbol per byte -> two per byte, high nibble first (write_nybble nybble_compression.c:1091)."""
    _need_cuda(sym, "sym")
    n = sym.numel()
    if out is None:
        out = torch.empty((n + 1) // 2, dtype=torch.uint8, device=sym.device)
    if status is None:
        status = torch.empty(1, dtype=torch.int32, device=sym.device)
    check(lib().dc_nybble_pack(sym.data_ptr(), n, out.data_ptr(), status.data_ptr(), _stream()), "dc_nybble_pack")
    return out, status


def nybble_unpack(packed: torch.Tensor, n_sym: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Inverse of nybble_pack (decoder split nybble_compression.c:767-769)."""
    _need_cuda(packed, "packed")
    if packed.numel() < (n_sym + 1) // 2:
        raise ValueError("packed buffer too small")
    if out is None:
        out = torch.empty(n_sym, dtype=torch.uint8, device=packed.device)
    check(lib().dc_nybble_unpack(packed.data_ptr(), n_sym, out.data_ptr(), _stream()), "dc_nybble_unpack")
    return out


def trit_pack(t2: torch.Tensor, ntrits: int, out: torch.Tensor | None = None):
    """Radix 3: the kernels' 2-bit-per-trit stream -> the 5-trits-per-byte payload (n_ary_huffman.c:745-748).
    Returns (payload[: ceil(ntrits / 5)], status)."""
    _need_cuda(t2, "t2")
    nb = (ntrits + 4) // 5
    if out is None:
        out = torch.empty(max(nb, 1) + 16, dtype=torch.uint8, device=t2.device)
    status = torch.empty(1, dtype=torch.int32, device=t2.device)
    check(lib().dc_trit_pack(t2.data_ptr(), ntrits, out.data_ptr(), status.data_ptr(), _stream()), "dc_trit_pack")
    return out[:nb], status


def trit_unpack(payload: torch.Tensor, ntrits: int):
    """Inverse of trit_pack.  Returns (t2 stream with slack for dc_huff_decode's vector loads, status)."""
    _need_cuda(payload, "payload")
    t2 = torch.empty((2 * ntrits + 7) // 8 + 64, dtype=torch.uint8, device=payload.device)
    status = torch.empty(1, dtype=torch.int32, device=payload.device)
    check(lib().dc_trit_unpack(payload.data_ptr(), ntrits, t2.data_ptr(), status.data_ptr(), _stream()), "dc_trit_unpack")
    return t2, status


def base64url_pack(bits: torch.Tensor, nbits: int) -> torch.Tensor:
    """The binary payload as base64url characters, 6 bits each through int2digit() (n_ary_huffman.c:371-426, :1646-1671):
    ceil(nbits / 6) characters, RFC 4648 without padding."""
    _need_cuda(bits, "bits")
    nchars = (nbits + 5) // 6
    out = torch.empty(max(nchars, 1) + 16, dtype=torch.uint8, device=bits.device)
    check(lib().dc_base64url_pack(bits.data_ptr(), nbits, out.data_ptr(), _stream()), "dc_base64url_pack")
    return out[:nchars]


def base64url_unpack(chars: torch.Tensor, nbits: int):
    """Inverse (digit2int() :428-455 accepts both alphabets for 62 / 63).  Returns (bytes[: ceil(nbits / 8)] + slack, status)."""
    _need_cuda(chars, "chars")
    out = torch.empty((nbits + 7) // 8 + 64, dtype=torch.uint8, device=chars.device)
    status = torch.empty(1, dtype=torch.int32, device=chars.device)
    check(lib().dc_base64url_unpack(chars.data_ptr(), nbits, out.data_ptr(), status.data_ptr(), _stream()), "dc_base64url_unpack")
    return out, status


def _nybble_text(fn_name: str, src: torch.Tensor, cap: int):
    _need_cuda(src, "src")
    n = src.numel()
    out = torch.empty(cap + 16, dtype=torch.uint8, device=src.device)
    ws_fn = lib().dc_nybble_adaptive_workspace_bytes if "adaptive" in fn_name else lib().dc_nybble_text_workspace_bytes
    ws = torch.empty(max(ws_fn(n), 16), dtype=torch.uint8, device=src.device)
    out_len = torch.empty(1, dtype=torch.int64, device=src.device)
    status = torch.empty(1, dtype=torch.int32, device=src.device)
    check(getattr(lib(), fn_name)(src.data_ptr(), n, out.data_ptr(), cap, out_len.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                  ws.numel(), _stream()), fn_name)
    return out, out_len, status


def nybble_text_compress(src: torch.Tensor):
    """compress_bytestring(src, dst, false) nybble_compression.c:887 (static " etaoins" table) on a CUDA byte tensor.
    Returns (buffer, 1 x int64 length, 1 x int32 status); nothing blocks."""
    return _nybble_text("dc_nybble_text_compress", src, src.numel() + 2)


def nybble_text_decompress(src: torch.Tensor):
    """decompress_bytestring(src, dst, false) nybble_compression.c:734.  Returns (buffer, length, status)."""
    return _nybble_text("dc_nybble_text_decompress", src, 2 * src.numel() + 2)


def nybble_adaptive_compress(src: torch.Tensor):
    """nybble_compress() = compress_bytestring(src, dst, true) nybble_compression.c:1134: 16 move-to-front contexts.
    Returns (buffer, 1 x int64 length, 1 x int32 status); nothing blocks."""
    return _nybble_text("dc_nybble_adaptive_compress", src, src.numel() + 2)


def nybble_adaptive_decompress(src: torch.Tensor):
    """nybble_decompress() = decompress_bytestring(src, dst, true) :1117.  The hit nibbles are resolved by one serial
    walk over the output on the device (the chain the format imposes).  Returns (buffer, length, status)."""
    return _nybble_text("dc_nybble_adaptive_decompress", src, 2 * src.numel() + 2)


def _nybble_text_batch(fn_name: str, strings, modify: bool, slot) -> list:
    """Many strings through one launch (one thread per string).  `strings`: a sequence of bytes objects; returns a list of bytes."""
    import numpy as np
    dev = torch.device("cuda", torch.cuda.current_device())
    lens = np.array([len(s) for s in strings], dtype=np.int64)
    src_off = np.zeros(len(strings) + 1, dtype=np.int64)
    np.cumsum(lens, out=src_off[1:])
    dst_off = np.zeros(len(strings) + 1, dtype=np.int64)
    np.cumsum(np.array([slot(int(n)) for n in lens], dtype=np.int64), out=dst_off[1:])
    flat = np.frombuffer(b"".join(bytes(s) for s in strings), dtype=np.uint8)
    d_src = torch.from_numpy(flat.copy() if flat.size else np.zeros(1, dtype=np.uint8)).to(dev)
    d_dst = torch.zeros(max(int(dst_off[-1]), 1), dtype=torch.uint8, device=dev)
    d_so, d_do = torch.from_numpy(src_off).to(dev), torch.from_numpy(dst_off).to(dev)
    d_len = torch.zeros(max(len(strings), 1), dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    check(getattr(lib(), fn_name)(d_src.data_ptr(), d_so.data_ptr(), len(strings), int(bool(modify)), d_dst.data_ptr(), d_do.data_ptr(),
                                  d_len.data_ptr(), status.data_ptr(), _stream()), fn_name)
    check(int(status.item()), fn_name)
    out, out_len = d_dst.cpu().numpy(), d_len.cpu().numpy()
    return [out[int(dst_off[i]): int(dst_off[i]) + int(out_len[i])].tobytes() for i in range(len(strings))]


def nybble_text_compress_batch(strings, modify: bool = False) -> list:
    """compress_bytestring(s, dst, modify) nybble_compression.c:887 for every s of `strings`, one GPU thread per string
    (the "replicas" parallelism of the adaptive mode).  Returns the compressed strings."""
    return _nybble_text_batch("dc_nybble_text_compress_batch", strings, modify, lambda n: n + 2)


def nybble_text_decompress_batch(strings, modify: bool = False) -> list:
    """decompress_bytestring(s, dst, modify) nybble_compression.c:734 for every s of `strings`."""
    return _nybble_text_batch("dc_nybble_text_decompress_batch", strings, modify, lambda n: 2 * n + 2)


def synth_fill(out: torch.Tensor, seed: int, thresholds: torch.Tensor, value_base: int) -> torch.Tensor:
    """Fill `out` (uint8, CUDA) with the counter-based synthetic stream (see synth.py)."""
    _need_cuda(out, "out")
    _need_cuda(thresholds, "thresholds")
    if thresholds.dtype != torch.int64 and thresholds.dtype != torch.int32:
        raise ValueError("thresholds: int32 view of the u32 table expected")
    check(lib().dc_synth_fill(out.data_ptr(), out.numel(), seed, thresholds.data_ptr(), thresholds.numel(), value_base,
                              _stream()), "dc_synth_fill")
    return out
