"""CPU, world_size 2 over gloo: the shard layer's arithmetic (histogram all-reduce, bit-offset scan, boundary byte
merge, shard-local decode) with an oracle-backed stand-in for the CUDA backend.  The concatenation of the shard
buffers must equal the single-stream oracle payload bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class OracleBackend:
    """Test-only compute backend: same interface as shard.CudaBackend, CPU tensors, oracle arithmetic."""

    def __init__(self):
        from oracle import pyoracle as O
        self.O = O

    def histogram(self, data):
        return torch.from_numpy(self.O.histogram_u8(data.numpy()).astype(np.int64))

    def build(self, hist, n_ary):
        lengths, el, ev, st = self.O.build_tables(hist.numpy().astype(np.uint64), n_ary)
        assert st == 0
        return {"lengths": lengths, "el": el, "ev": ev, "n": n_ary, "bpd": self.O.bits_per_digit(n_ary)}

    def bits_for_hist(self, hist, table):
        h = hist.numpy()[:256].astype(np.int64)
        return torch.tensor([int((h * table["lengths"][:256] * table["bpd"]).sum())], dtype=torch.int64)

    def encode(self, data, table, bit_phase, nbits_hint):
        payload, bits = self.O.pack(data.numpy(), table["el"], table["ev"], table["bpd"], bit_phase)
        buf = torch.zeros(payload.size + 8, dtype=torch.uint8)
        buf[: payload.size] = torch.from_numpy(payload)
        return buf, torch.tensor([bits], dtype=torch.int64), torch.tensor([0], dtype=torch.int32)

    def decode(self, payload, nbits, table, n_out, bit_start):
        out = self.O.unpack(payload.numpy(), bit_start, nbits, table["lengths"], table["n"], n_out)
        return torch.from_numpy(out), torch.tensor([0], dtype=torch.int32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_ary, sizes, results):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from data_compression_b200 import synth
        from data_compression_b200.shard import ShardedHuffman
        from oracle import pyoracle as O
        thr, base = synth.zipf_bytes_spec()
        total = sum(sizes)
        stream = synth.host_stream(total, 77, thr, base)
        lo = sum(sizes[:rank])
        local = torch.from_numpy(stream[lo: lo + sizes[rank]].copy())
        sh = ShardedHuffman(backend=OracleBackend())
        sp = sh.encode(local, n_ary)
        whole = sh.gather_stream(sp)
        back = sh.decode(sp)
        ok = bool(torch.equal(back, local))
        offs, tot = sh.symbol_offsets(local.numel())
        ok &= offs[rank] == lo and tot == total
        if rank == 0:
            lengths, el, ev, st = O.build_tables(O.histogram_u8(stream), n_ary)
            want, wbits = O.pack(stream, el, ev, O.bits_per_digit(n_ary))
            ok &= wbits == sp.total_bits
            ok &= bool(np.array_equal(whole.numpy(), want))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_ary,sizes", [(2, (5000, 7001)), (16, (4096, 4096)), (4, (3, 1)), (2, (0, 900))])
def test_sharded_encode_equals_single_stream(n_ary, sizes):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_ary, sizes, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def test_offset_scan_and_boundary_merge_pure():
    from data_compression_b200.shard import exclusive_offsets, merge_boundary_bytes
    offs, total = exclusive_offsets([13, 0, 5, 70])
    assert offs == [0, 13, 13, 18] and total == 88
    # three shards inside one byte (bits 0-2, 3-4, 5-10): byte 0 is shared by all of them
    ranges = [(0, 1), (0, 1), (0, 2)]
    fl = [(0b10100000, 0b10100000), (0b00011000, 0b00011000), (0b00000101, 0b11100000)]
    assert merge_boundary_bytes(0, fl, ranges) == (0b10111101, 0b10111101)
    assert merge_boundary_bytes(1, fl, ranges) == (0b10111101, 0b10111101)
    assert merge_boundary_bytes(2, fl, ranges) == (0b10111101, 0b11100000)
    assert merge_boundary_bytes(0, [(1, 2), (0, 0)], [(0, 5), (7, 7)]) == (1, 2)
