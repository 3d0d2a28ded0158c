"""Multi-GPU shard layer: one process per GPU, torch.distributed for the plumbing.

Both paths shard naturally; the only data-path exchanges are tiny:

  encode   contiguous byte ranges per rank
           local histogram --all_reduce(259 x i64)--> one global code table (every rank builds it, identically)
           local bit total = dot(local histogram, global lengths)   (no data pass)
           --all_gather(1 x i64)--> exclusive scan = this rank's global bit offset O_r
           the rank encodes at phase O_r mod 8 into its own buffer; the byte a shard shares with its
           neighbour is OR-combined (--all_gather(2 bytes)--), so the concatenation of the shard buffers at
           byte offsets O_r // 8 IS the single-stream payload, bit for bit
  decode   every rank decodes the codes that start in its shard (start phase and bit count are the encoder's
           side information), no exchange at all; symbol counts are all-gathered for the output offsets

The compute backend is injected: the default is the CUDA library (no CPU fallback); the CPU tests in
tests/ pass a CPU stand-in of their own so the arithmetic of this layer is exercised with gloo, world_size 2.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


class CudaBackend:
    """The product backend: every call lands in libdc_b200.so."""

    def __init__(self):
        from . import api
        self.api = api

    def histogram(self, data):
        return self.api.histogram(data)

    def build(self, hist, n_ary):
        return self.api.huff_build(hist, n_ary)

    def bits_for_hist(self, hist, table):
        return self.api.huff_bits_for_hist(hist, table)

    def encode(self, data, table, bit_phase, nbits_hint):
        cap = (nbits_hint + bit_phase + 7) // 8 + 64
        out = torch.empty(cap, dtype=torch.uint8, device=data.device)
        res = self.api.huff_encode(data, table, out=out, bit_phase=bit_phase)
        return res.payload, res.total_bits, res.status

    def decode(self, payload, nbits, table, n_out, bit_start):
        out, status = self.api.huff_decode(payload, nbits, table, n_out, bit_start=bit_start)
        return out, status


@dataclass
class ShardedPayload:
    payload: torch.Tensor       # this rank's bytes [bit_offset // 8, ceil((bit_offset + nbits) / 8)) of the stream
    nbits: int                  # code bits of this shard
    bit_offset: int             # O_r: global bit offset of the shard's first code
    total_bits: int             # bits of the whole stream
    n_symbols: int              # symbols of this shard
    table: object               # the global code table
    shard_bits: list            # bits of every shard (the side information a container would carry)

    @property
    def bit_phase(self) -> int:
        return self.bit_offset % 8

    @property
    def byte_offset(self) -> int:
        return self.bit_offset // 8

    @property
    def nbytes(self) -> int:
        return (self.bit_phase + self.nbits + 7) // 8 if self.nbits else 0


def exclusive_offsets(bits_per_rank):
    """O_r = sum of the bit totals of the ranks before r (pure function: tested on CPU)."""
    off, out = 0, []
    for b in bits_per_rank:
        out.append(off)
        off += int(b)
    return out, off


def merge_boundary_bytes(my_rank, first_last, byte_ranges):
    """Given every rank's (first byte value, last byte value) and global byte range [lo, hi) of its buffer,
    return the values this rank's first and last bytes must take so that shared bytes agree everywhere.
    A byte can be shared by more than two shards when shards are tiny."""
    lo, hi = byte_ranges[my_rank]
    if hi <= lo:
        return None, None
    first, last = first_last[my_rank]
    for q, (qlo, qhi) in enumerate(byte_ranges):
        if q == my_rank or qhi <= qlo:
            continue
        qf, ql = first_last[q]
        for idx, val in ((qlo, qf), (qhi - 1, ql)):
            if idx == lo:
                first |= val
            if idx == hi - 1:
                last |= val
    if hi - lo == 1:
        first = last = first | last
    return first, last


class ShardedHuffman:
    def __init__(self, group=None, backend=None, device=None):
        self.group = group
        self.backend = backend or CudaBackend()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device

    # --- collectives on tiny tensors
    def _all_reduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return t

    def _all_gather_list(self, t):
        if self.world == 1:
            return [t]
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t, group=self.group)
        return outs

    def encode(self, local: torch.Tensor, n_ary: int) -> ShardedPayload:
        be = self.backend
        hist = be.histogram(local)
        ghist = self._all_reduce(hist.clone())
        table = be.build(ghist, n_ary)
        my_bits = be.bits_for_hist(hist, table)
        bits = [int(t.item()) for t in self._all_gather_list(my_bits)]
        offsets, total = exclusive_offsets(bits)
        off = offsets[self.rank]
        payload, d_bits, d_status = be.encode(local, table, off % 8, bits[self.rank])
        st = int(d_status.item())
        if st != 0:
            from ._lib import DcError
            raise DcError(st, "sharded encode")
        assert int(d_bits.item()) == bits[self.rank], "dot(local histogram, lengths) disagrees with the encoder"
        sp = ShardedPayload(payload, bits[self.rank], off, total, local.numel(), table, bits)
        sp.payload = payload[: sp.nbytes]
        self._merge_boundaries(sp, offsets, bits)
        return sp

    def _merge_boundaries(self, sp: ShardedPayload, offsets, bits) -> None:
        if self.world == 1:
            return
        ranges = []
        for o, b in zip(offsets, bits):
            ranges.append((o // 8, (o + b + 7) // 8 if b else o // 8))
        mine = torch.zeros(2, dtype=torch.uint8, device=sp.payload.device)
        if sp.nbytes:
            mine[0] = sp.payload[0]
            mine[1] = sp.payload[sp.nbytes - 1]
        allb = [tuple(int(x) for x in t.cpu()) for t in self._all_gather_list(mine)]
        first, last = merge_boundary_bytes(self.rank, allb, ranges)
        if first is not None:
            sp.payload[0] = first
            sp.payload[sp.nbytes - 1] = last

    def decode(self, sp: ShardedPayload) -> torch.Tensor:
        out, status = self.backend.decode(sp.payload, sp.nbits, sp.table, sp.n_symbols, sp.bit_phase)
        st = int(status.item())
        if st != 0:
            from ._lib import DcError
            raise DcError(st, "sharded decode")
        return out

    def decode_stream(self, part: torch.Tensor, part_bytes: int, total_bits: int, table, n_total: int):
        """Decode ONE bitstream that is cut into equal byte ranges over the ranks (BASELINE config 5): rank r holds bytes
        [r * part_bytes, ...) of the payload in `part`; part_bytes is a multiple of 1024 (the last rank's part may be
        shorter).  No side information about code boundaries: every rank finds its first code by synchronising over the
        last 1024 bytes of its left neighbour, the ranks compare what they assumed with where their neighbours really
        ended (an all-gather of 24 bytes), and the symbol counts become output offsets.  Returns (symbols, offset)."""
        from . import api
        if part_bytes % api.SHARD_ALIGN:
            raise ValueError("part_bytes must be a multiple of 1024")
        dev = part.device
        r, G = self.rank, self.world
        total_bytes = (total_bits + 7) // 8
        lo = min(r * part_bytes, total_bytes)
        hi = min(lo + part_bytes, total_bytes)
        mine = hi - lo
        # halos: the last 1024 bytes and the first 1024 bytes of every part
        edge = torch.zeros(2 * api.SHARD_ALIGN, dtype=torch.uint8, device=dev)
        if mine:
            k = min(mine, api.SHARD_ALIGN)
            edge[:k] = part[:k]
            edge[2 * api.SHARD_ALIGN - k:] = part[mine - k: mine]
        edges = self._all_gather_list(edge)
        buf = torch.zeros(api.SHARD_ALIGN + mine + api.SHARD_ALIGN, dtype=torch.uint8, device=dev)
        if r > 0:
            buf[: api.SHARD_ALIGN] = edges[r - 1][api.SHARD_ALIGN:]
        buf[api.SHARD_ALIGN: api.SHARD_ALIGN + mine] = part[:mine]
        if r + 1 < G:
            buf[api.SHARD_ALIGN + mine:] = edges[r + 1][: api.SHARD_ALIGN]
        left = total_bits - 8 * lo if mine else 0
        my_bits = min(8 * mine, left)
        dec = api.ShardDecoder(buf, mine, my_bits, left, table) if my_bits > 0 else None
        summ = dec.sync(has_halo=r > 0) if dec else torch.zeros(3, dtype=torch.int64, device=dev)
        for _ in range(G + 1):
            infos = [api.ShardDecoder.unpack(t.cpu()) for t in self._all_gather_list(summ)]
            active = [g for g in range(G) if min(8 * part_bytes, max(total_bits - 8 * g * part_bytes, 0)) > 0]
            if any(infos[g]["resync"] for g in active):
                raise RuntimeError("the stream does not self-synchronise within 8192 bits; decode it on one device")
            # where does every shard's first code really start?  (the previous active shard's exit; 0 for the first)
            wrong = [g for i, g in enumerate(active) if i > 0 and infos[g]["assumed_start"] != infos[active[i - 1]]["exit"]]
            if not wrong:
                break
            if r in wrong and dec:
                prev = active[active.index(r) - 1]
                summ = dec.sync(has_halo=False, first_code_bit=infos[prev]["exit"])
        else:
            raise RuntimeError("shard starts did not settle")
        counts = [infos[g]["symbols"] if g in active else 0 for g in range(G)]
        offsets, total = exclusive_offsets(counts)
        if total != n_total:
            from ._lib import DC_ERR_CORRUPT, DcError
            raise DcError(DC_ERR_CORRUPT, f"sharded decode: {total} symbols, expected {n_total}")
        if not dec:
            return torch.empty(0, dtype=torch.uint8, device=dev), offsets[r]
        out, status = dec.write(counts[r])
        st = int(status.item())
        if st != 0:
            from ._lib import DcError
            raise DcError(st, "sharded decode")
        return out, offsets[r]

    def symbol_offsets(self, n_local: int):
        """Exclusive scan of the per-rank symbol counts (the decode side's output offsets)."""
        t = torch.tensor([n_local], dtype=torch.int64, device=self.device or "cpu")
        counts = [int(x.item()) for x in self._all_gather_list(t)]
        return exclusive_offsets(counts)

    def gather_stream(self, sp: ShardedPayload):
        """Rank 0 gets the whole logical bitstream (test/verification helper: moves C bytes)."""
        sizes = [int(x.item()) for x in self._all_gather_list(torch.tensor([sp.nbytes], dtype=torch.int64,
                                                                           device=sp.payload.device))]
        offs = [int(x.item()) for x in self._all_gather_list(torch.tensor([sp.byte_offset], dtype=torch.int64,
                                                                          device=sp.payload.device))]
        cap = max(sizes) if sizes else 0
        buf = torch.zeros(max(cap, 1), dtype=torch.uint8, device=sp.payload.device)
        buf[: sp.nbytes] = sp.payload[: sp.nbytes]
        parts = self._all_gather_list(buf)
        if self.rank != 0:
            return None
        total_bytes = (sp.total_bits + 7) // 8
        out = torch.zeros(total_bytes, dtype=torch.uint8, device=sp.payload.device)
        for p, sz, o in zip(parts, sizes, offs):
            if sz:
                out[o: o + sz] |= p[:sz]
        return out


class NcclShards:
    """The C-ABI shard layer (dc_shard_* in include/dc_b200.h, csrc/shard_nccl.cu): NCCL collectives issued by the
    library itself on a communicator of its own.  torch.distributed is only used once, to hand rank 0's NCCL id to
    the other ranks; a single process (world 1) needs no process group at all."""

    def __init__(self, device=None, group=None):
        import ctypes as C

        from ._lib import check, lib
        self.C, self.L, self.check = C, lib(), check
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            check(self.L.dc_shard_unique_id(buf), "dc_shard_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if self.world > 1:
            t = ident.to(self.device) if dist.get_backend(group) == "nccl" else ident
            dist.broadcast(t, src=0, group=group)
            ident = t.cpu()
        raw = (C.c_ubyte * 128)(*[int(x) for x in ident])
        comm = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.L.dc_shard_comm_create(raw, self.rank, self.world, C.byref(comm)), "dc_shard_comm_create")
        self.comm = comm

    def close(self):
        if self.comm:
            self.L.dc_shard_comm_destroy(self.comm)
            self.comm = None

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    # ---- encode (BASELINE config 4)
    def encode_buffers(self, n_local: int):
        dev = self.device
        from .api import HuffTable
        return {"out": torch.empty(n_local + n_local // 4 + 4096, dtype=torch.uint8, device=dev),
                "ws": torch.empty(max(self.L.dc_shard_huff_encode_workspace_bytes(n_local, self.world), 16), dtype=torch.uint8, device=dev),
                "bits": torch.empty(1, dtype=torch.int64, device=dev), "status": torch.empty(1, dtype=torch.int32, device=dev),
                "table": HuffTable(dev)}

    def encode(self, local: torch.Tensor, n_ary: int, buf=None):
        """Stream-ordered, non-blocking: returns the buffers (out, ws, bits, status, table) the call filled."""
        buf = buf or self.encode_buffers(local.numel())
        self.check(self.L.dc_shard_huff_encode(self.comm, local.data_ptr(), local.numel(), n_ary, buf["table"].ptr, buf["out"].data_ptr(),
                                               buf["out"].numel(), buf["bits"].data_ptr(), buf["status"].data_ptr(), buf["ws"].data_ptr(),
                                               buf["ws"].numel(), self._stream()), "dc_shard_huff_encode")
        return buf

    def encode_info(self, buf, n_local: int):
        """Blocking: (bit offset of this shard, its bits, bits of the whole stream)."""
        C = self.C
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.check(self.L.dc_shard_huff_encode_info(buf["ws"].data_ptr(), n_local, self.world, C.byref(a), C.byref(b), C.byref(c),
                                                    self._stream()), "dc_shard_huff_encode_info")
        self.check(int(buf["status"].item()), "dc_shard_huff_encode")
        return a.value, b.value, c.value

    def gather(self, buf, n_local: int, total_bits: int, root: int = 0, out=None):
        """Blocking: the whole stream in one buffer on `root` (None elsewhere)."""
        nbytes = (total_bits + 7) // 8
        if self.rank == root and out is None:
            out = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
        self.check(self.L.dc_shard_huff_gather(self.comm, root, buf["out"].data_ptr(), buf["ws"].data_ptr(), n_local,
                                               out.data_ptr() if out is not None else None, out.numel() if out is not None else 0,
                                               self._stream()), "dc_shard_huff_gather")
        return out[:nbytes] if self.rank == root else None

    # ---- decode of one blindly cut stream (BASELINE config 5)
    def decode_buffers(self, part_bytes: int, out_capacity: int):
        dev = self.device
        return {"buf": torch.zeros(1024 + part_bytes + 1024, dtype=torch.uint8, device=dev),
                "out": torch.empty(max(out_capacity, 16), dtype=torch.uint8, device=dev),
                "ws": torch.empty(max(self.L.dc_shard_huff_decode_workspace_bytes(part_bytes, self.world), 16), dtype=torch.uint8, device=dev),
                "status": torch.empty(1, dtype=torch.int32, device=dev)}

    def decode_stream(self, buf, part_bytes: int, total_bits: int, table):
        """buf["buf"][1024 : 1024 + mine] holds this rank's bytes.  Blocking.  Returns (symbols tensor, offset, total)."""
        C = self.C
        n, off, tot = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.check(self.L.dc_shard_huff_decode_stream(self.comm, buf["buf"].data_ptr(), part_bytes, total_bits, table.ptr, buf["out"].data_ptr(),
                                                      buf["out"].numel(), C.byref(n), C.byref(off), C.byref(tot), buf["status"].data_ptr(),
                                                      buf["ws"].data_ptr(), buf["ws"].numel(), self._stream()), "dc_shard_huff_decode_stream")
        return buf["out"][: n.value], off.value, tot.value

    # ---- nybble shards (no exchange)
    def nybble_range(self, n_total: int):
        C = self.C
        lo, hi = C.c_uint64(0), C.c_uint64(0)
        self.check(self.L.dc_shard_nybble_range(n_total, self.rank, self.world, C.byref(lo), C.byref(hi)), "dc_shard_nybble_range")
        return lo.value, hi.value

    def nybble_pack(self, n_total: int, sym_local: torch.Tensor):
        lo, hi = self.nybble_range(n_total)
        packed = torch.empty(max((hi - lo + 1) // 2, 1), dtype=torch.uint8, device=sym_local.device)
        status = torch.zeros(1, dtype=torch.int32, device=sym_local.device)
        self.check(self.L.dc_shard_nybble_pack(n_total, self.rank, self.world, sym_local.data_ptr(), packed.data_ptr(), status.data_ptr(),
                                               self._stream()), "dc_shard_nybble_pack")
        return packed[: (hi - lo + 1) // 2], status

    def nybble_unpack(self, n_total: int, packed_local: torch.Tensor):
        lo, hi = self.nybble_range(n_total)
        sym = torch.empty(max(hi - lo, 1), dtype=torch.uint8, device=packed_local.device)
        self.check(self.L.dc_shard_nybble_unpack(n_total, self.rank, self.world, packed_local.data_ptr(), sym.data_ptr(), self._stream()),
                   "dc_shard_nybble_unpack")
        return sym[: hi - lo]
