"""GPU: the shard layer with the real CUDA backend.  Two ranks (both on cuda:0 when the box has one GPU, one
per GPU otherwise) exchange the histogram / bit totals / boundary bytes over gloo; the concatenation of the
shard payloads must equal the single-stream oracle payload bit for bit (BASELINE configs 4 and 5 at test size),
and every rank decodes its own shard back from its global bit phase."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_ary, sizes, results):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import data_compression_b200 as dc
        from data_compression_b200 import synth
        from data_compression_b200.shard import ShardedHuffman
        from oracle import pyoracle as O
        launches0 = dc.launch_count()
        thr, base = synth.zipf_bytes_spec()
        total = sum(sizes)
        stream = synth.host_stream(total, 4242, thr, base)
        lo = sum(sizes[:rank])
        local = torch.from_numpy(stream[lo: lo + sizes[rank]].copy()).to(dev)
        sh = ShardedHuffman(device=dev)
        sp = sh.encode(local, n_ary)
        whole = sh.gather_stream(sp)
        back = sh.decode(sp)
        ok = bool(torch.equal(back[: local.numel()], local))
        offs, tot = sh.symbol_offsets(local.numel())
        ok &= offs[rank] == lo and tot == total
        ok &= dc.launch_count() > launches0          # the CUDA library did the work
        if rank == 0:
            O.build()
            lengths, el, ev, st = O.build_tables(O.histogram_u8(stream), n_ary)
            want, wbits = O.pack(stream, el, ev, O.bits_per_digit(n_ary))
            ok &= wbits == sp.total_bits
            ok &= bool(np.array_equal(whole.cpu().numpy(), want))
            # config 5: the whole (oracle-produced) stream decodes from the bitstream alone
            full = dc.huff_decompress(torch.from_numpy(want).to(dev), wbits, sp.table, total)
            ok &= bool(np.array_equal(full.cpu().numpy(), stream))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_ary,sizes", [(16, (300001, 200003)), (2, (123457, 400001)), (4, (70000, 5))])
def test_sharded_cuda_encode_equals_single_stream(n_ary, sizes):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_ary, sizes, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def _stream_worker(rank, world, port, n_ary, n, results, s_exp=1.1):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import data_compression_b200 as dc
        from data_compression_b200 import synth
        from data_compression_b200.shard import ShardedHuffman
        from oracle import pyoracle as O
        O.build()
        thr, base = synth.zipf_thresholds(255, s_exp), 1
        stream = synth.host_stream(n, 1234 + n_ary, thr, base)
        # the bitstream comes from the ORACLE (BASELINE config 5: "reference-produced"), and is cut blindly into byte ranges
        lengths, el, ev, st = O.build_tables(O.histogram_u8(stream), n_ary)
        payload, total_bits = O.pack(stream, el, ev, O.bits_per_digit(n_ary))
        part_bytes = ((payload.size + world - 1) // world + 1023) // 1024 * 1024
        lo = min(rank * part_bytes, payload.size)
        part = torch.from_numpy(payload[lo: lo + part_bytes].copy()).to(dev)
        table = dc.huff_table_from_lengths(torch.from_numpy(lengths.astype(np.int32)).to(dev), n_ary)
        sh = ShardedHuffman(device=dev)
        out, off = sh.decode_stream(part, part_bytes, total_bits, table, n)
        ok = bool(np.array_equal(out.cpu().numpy(), stream[off: off + out.numel()]))
        sizes = [int(x.item()) for x in sh._all_gather_list(torch.tensor([out.numel()], dtype=torch.int64, device=dev))]
        ok &= sum(sizes) == n and off == sum(sizes[:rank])
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_ary,n,world,s_exp", [(2, 700001, 2, 1.1), (4, 300000, 2, 1.1), (16, 123457, 2, 1.1), (2, 5000, 2, 1.1),
                                                 (4, 2000000, 4, 1.1), (4, 900001, 2, 1.5), (2, 900001, 2, 1.5)])
def test_one_stream_cut_blindly_over_ranks(n_ary, n, world, s_exp):
    """Config 5: one oracle-produced bitstream, byte ranges per rank, first codes found by boundary synchronisation.
    Zipf(1.5) gives tables whose longest code exceeds the 12-bit index (the 14-bit count table / the escape forms)."""
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_stream_worker, args=(world, _free_port(), n_ary, n, results, s_exp), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)
