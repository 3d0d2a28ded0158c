"""dc_shard_* (csrc/shard_nccl.cu) through ctypes: the multi-GPU path behind the C-ABI, NCCL issued by the library.
World 1 runs everywhere (a one-rank communicator exercises every collective call); world 2 runs when the box has
two GPUs (one process per GPU).  Every result is compared with the CPU oracle's single-stream payload."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _stream(n, seed=5):
    from data_compression_b200 import synth
    thr, base = synth.zipf_bytes_spec()
    return synth.host_stream(n, synth.SEED_BASE + seed, thr, base)


def _oracle_payload(host, n_ary):
    from oracle import pyoracle as O
    hist = O.histogram_u8(host)
    lengths, el, ev, st = O.build_tables(hist, n_ary)
    assert st == 0
    if n_ary == 3:        # the kernels' stream for radix 3: one 2-bit field per trit (the octet payload is packed from it)
        payload, bits = O.pack(host, el, O.field_values(el, ev, 3, 2), 2)
    elif 5 <= n_ary < 16:   # one nibble per digit
        payload, bits = O.pack(host, el, O.nibble_values(el, ev, n_ary), 4)
    else:
        payload, bits = O.pack(host, el, ev, O.bits_per_digit(n_ary))
    return payload, bits, lengths


def _worker(rank, world, port, n_total, n_ary, q):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from data_compression_b200 import shard
    dev = torch.device("cuda", rank)
    host = _stream(n_total)
    per = (n_total + world - 1) // world
    per += (-per) % 16
    lo, hi = min(rank * per, n_total), min((rank + 1) * per, n_total)
    local = torch.from_numpy(host[lo:hi].copy()).to(dev)
    S = shard.NcclShards(dev)
    # ---- config 4: sharded encode == the single-stream payload
    buf = S.encode(local, n_ary)
    off, bits, total = S.encode_info(buf, local.numel())
    stream = S.gather(buf, local.numel(), total, root=0)
    o_payload, o_bits, lengths = _oracle_payload(host, n_ary)
    ok = total == o_bits
    if rank == 0:
        ok = ok and np.array_equal(stream.cpu().numpy(), o_payload)
    # my shard alone must equal its byte range of the stream, shared bytes included
    nb = ((off % 8) + bits + 7) // 8 if bits else 0
    ok = ok and np.array_equal(buf["out"][:nb].cpu().numpy(), o_payload[off // 8: off // 8 + nb])
    # ---- config 5: the oracle's stream cut blindly into byte ranges
    total_bytes = (o_bits + 7) // 8
    part = ((total_bytes + world - 1) // world + 1023) // 1024 * 1024
    plo, phi = min(rank * part, total_bytes), min((rank + 1) * part, total_bytes)
    dbuf = S.decode_buffers(part, per + 4096)
    dbuf["buf"][1024: 1024 + (phi - plo)] = torch.from_numpy(o_payload[plo:phi].copy()).to(dev)
    sym, soff, stot = S.decode_stream(dbuf, part, o_bits, buf["table"])
    ok = ok and stot == n_total and int(dbuf["status"].item()) == 0
    ok = ok and np.array_equal(sym.cpu().numpy(), host[soff: soff + sym.numel()])
    counts = torch.tensor([sym.numel()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    ok = ok and int(counts.item()) == n_total
    # ---- nybble shards: even starts, no exchange
    nsym = n_total + 1   # odd on purpose
    syms = (np.arange(nsym, dtype=np.uint64) * 2654435761 >> 7).astype(np.uint8) & 15
    a, b = S.nybble_range(nsym)
    packed, st = S.nybble_pack(nsym, torch.from_numpy(syms[a:b].copy()).to(dev))
    from oracle import pyoracle as O
    whole = O.nybble_pack(syms)
    ok = ok and a % 2 == 0 and int(st.item()) == 0 and np.array_equal(packed.cpu().numpy(), whole[a // 2: a // 2 + (b - a + 1) // 2])
    back = S.nybble_unpack(nsym, packed)
    ok = ok and np.array_equal(back.cpu().numpy(), syms[a:b])
    S.close()
    if world > 1:
        dist.destroy_process_group()
    q.put((rank, bool(ok)))


@pytest.mark.parametrize("n_ary", [2, 4, 16, 10, 3])
def test_shard_c_abi_world1(n_ary):
    import queue
    q = queue.Queue()
    _worker(0, 1, 0, 3 * (1 << 20) + 12345, n_ary, q)
    assert q.get() == (0, True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (one process per GPU)")
@pytest.mark.parametrize("n_ary", [4, 16, 10, 3])
def test_shard_c_abi_world2(n_ary):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n_ary
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5 * (1 << 20) + 777, n_ary, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _edge_worker(rank, world, port, sizes, n_ary, q):
    """Sharded encode only, many small streams in one process pair: the shared bytes are completed from the neighbours' edge
    symbols (no second collective), including shards of fewer than eight symbols and empty shards."""
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from data_compression_b200 import shard
    dev = torch.device("cuda", rank)
    S = shard.NcclShards(dev)
    bad = []
    for n_total, cut in sizes:
        host = _stream(n_total, seed=9 + n_total % 7)
        lo, hi = (0, cut) if rank == 0 else (cut, n_total)
        local = torch.from_numpy(host[lo:hi].copy()).to(dev)
        buf = S.encode(local, n_ary)
        off, bits, total = S.encode_info(buf, local.numel())
        o_payload, o_bits, _ = _oracle_payload(host, n_ary)
        nb = ((off % 8) + bits + 7) // 8 if bits else 0
        if total != o_bits or not np.array_equal(buf["out"][:nb].cpu().numpy(), o_payload[off // 8: off // 8 + nb]):
            bad.append((n_total, cut))
    S.close()
    dist.destroy_process_group()
    q.put((rank, bad))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (one process per GPU)")
@pytest.mark.parametrize("n_ary", [2, 16])
def test_shard_edges_completed_locally_world2(n_ary):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    # (symbols in the stream, symbols on rank 0): cuts inside bytes, shards shorter than the eight edge symbols, empty shards
    sizes = [(5, 5), (5, 0), (19, 16), (19, 3), (40, 32), (2, 1), (9, 1), (9, 8), (4097, 4096), (4097, 1), (100003, 50001), (100003, 99999)]
    procs = [ctx.Process(target=_edge_worker, args=(r, 2, 29650 + n_ary, sizes, n_ary, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, []), (1, [])]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (one process per GPU)")
def test_shard_exchange_nccl_fallback_world2(monkeypatch):
    """DC_SHARD_PEER=0: the histogram exchange through ncclAllGather instead of peer memory -- same streams."""
    import torch.multiprocessing as mp
    monkeypatch.setenv("DC_SHARD_PEER", "0")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    sizes = [(19, 16), (4097, 4096), (100003, 50001)]
    procs = [ctx.Process(target=_edge_worker, args=(r, 2, 29671, sizes, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, []), (1, [])]
