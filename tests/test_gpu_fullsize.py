"""GPU: BASELINE.json's configurations at their full single-GPU sizes.  The payload of 1 GiB of Zipf(1.1) bytes is
compared bit for bit with the oracle's (block-parallel on the host cores, a few seconds), the decoder must return the
input from the ORACLE's bitstream, and the 2^30-symbol nybble stream must pack to the oracle's bytes and back."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N = 1 << 30


def _host_of(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy()


@pytest.mark.parametrize("n_ary", [4, 2])
def test_config3_one_gib_huffman_equals_oracle(dc, oracle, n_ary):
    from data_compression_b200 import synth
    threads = os.cpu_count() or 1
    thr, base = synth.zipf_bytes_spec()
    data = torch.empty(N, dtype=torch.uint8, device="cuda")
    dc.synth_fill(data, synth.SEED_BASE + 2, synth.device_thresholds(thr, "cuda"), base)
    host = _host_of(data)
    hist = dc.histogram(data)
    o_hist = oracle.histogram_u8(host, threads=threads)
    assert np.array_equal(hist.cpu().numpy().astype(np.uint64), o_hist)
    table = dc.huff_build(hist, n_ary)
    ln, el, ev, st = oracle.build_tables(o_hist, n_ary)
    t = table.download()
    assert st == 0 and np.array_equal(np.array(t.lengths[:259]), ln) and np.array_equal(np.array(t.values[:259], dtype=np.uint32), ev)
    res = dc.huff_encode(data, table)
    nbits = res.bits()
    want, wbits, _ = oracle.pack_mt(host, el, ev, oracle.bits_per_digit(n_ary), 0, threads=threads,
                                    out=np.empty(N + N // 4 + 64, dtype=np.uint8))
    assert nbits == wbits == t.total_bits
    nbytes = (nbits + 7) // 8
    got = _host_of(res.payload[:nbytes])
    assert np.array_equal(got, want[:nbytes])
    # decode the oracle's stream (not our own): BASELINE config 5 in its one-GPU form
    d_bits = torch.from_numpy(want[:nbytes + 64].copy()).cuda()
    del res
    out = dc.huff_decompress(d_bits, wbits, table, N)
    assert torch.equal(out, data)


def test_config2_nybble_stream_equals_oracle(dc, oracle):
    from data_compression_b200 import synth
    threads = os.cpu_count() or 1
    thr4, base4 = synth.zipf_nybble_spec()
    sym = torch.empty(N, dtype=torch.uint8, device="cuda")
    dc.synth_fill(sym, synth.SEED_BASE + 1, synth.device_thresholds(thr4, "cuda"), base4)
    packed, st = dc.nybble_pack(sym)
    assert int(st.item()) == 0
    want = oracle.nybble_pack(_host_of(sym), threads=threads)
    assert np.array_equal(_host_of(packed), want)
    assert torch.equal(dc.nybble_unpack(packed, N), sym)


def test_radix3_quarter_gib_equals_oracle(dc, oracle):
    """Row N4 at size: 256 MiB of Zipf(1.1) bytes, radix 3 (the reference's default).  The 5-trits-per-byte payload is compared
    byte for byte with the oracle's (one host thread), and the oracle's payload is decoded back."""
    from data_compression_b200 import synth
    n = 1 << 28
    thr, base = synth.zipf_bytes_spec()
    data = torch.empty(n, dtype=torch.uint8, device="cuda")
    dc.synth_fill(data, synth.SEED_BASE + 2, synth.device_thresholds(thr, "cuda"), base)
    host = _host_of(data)
    hist = dc.histogram(data)
    table = dc.huff_build(hist, 3)
    ln, el, ev, st = oracle.build_tables(hist.cpu().numpy().astype(np.uint64), 3)
    t = table.download()
    assert st == 0 and t.status == 0 and np.array_equal(np.array(t.lengths[:259]), ln)
    res = dc.huff_encode(data, table, out=torch.empty(n + n // 2 + 64, dtype=torch.uint8, device="cuda"))
    want, wtrits = oracle.pack_trits(host, el, ev)
    assert res.bits() == 2 * wtrits == t.total_bits
    payload, pst = dc.trit_pack(res.payload, wtrits)
    assert int(pst.item()) == 0 and np.array_equal(_host_of(payload), want)
    del res, payload
    t2, ust = dc.trit_unpack(torch.from_numpy(want).cuda(), wtrits)
    out, status = dc.huff_decode(t2, 2 * wtrits, table, n)
    assert int(ust.item()) == 0 and int(status.item()) == 0 and torch.equal(out, data)


def test_adaptive_nybble_compressor_quarter_gib_equals_oracle(dc, oracle):
    """Row N3 at size: 256 MiB of word-like 7-bit text through nybble_compress() (move-to-front contexts): the GPU scan must
    give the oracle's bytes (the oracle is pinned against the unmodified reference in tests/test_oracle_vs_reference.py)."""
    n = 1 << 28
    rng = np.random.default_rng(1234)
    words = [b"the ", b"and ", b"this ", b"is ", b"a ", b"test. ", b"banana ", b"Hello, ", b"world. ", b"only ", b"of ", b"to ",
             b"compression ", b"nybble ", b"context ", b"Q", b"42 ", b"\n", b"entropy ", b"static "]
    lens = np.array([len(w) for w in words])
    picks = rng.integers(0, len(words), size=n // int(lens.mean()) + 1024)
    table = np.frombuffer(b"".join(w.ljust(16, b"\0") for w in words), dtype=np.uint8).reshape(len(words), 16)
    flat = table[picks].reshape(-1)
    host = flat[flat != 0][:n].copy()
    assert host.size == n
    want = oracle.nybble_adaptive_compress(host)
    buf, ln, st = dc.nybble_adaptive_compress(torch.from_numpy(host).cuda())
    assert int(st.item()) == 0 and int(ln.item()) == len(want)
    assert np.array_equal(_host_of(buf[: len(want)]), np.frombuffer(want, dtype=np.uint8))
    # the decoder's resolve step is serial: round-trip the first MiB of text only
    small = host[: 1 << 20]
    comp = oracle.nybble_adaptive_compress(small)
    back, bl, st = dc.nybble_adaptive_decompress(torch.from_numpy(np.frombuffer(comp, dtype=np.uint8).copy()).cuda())
    assert int(st.item()) == 0 and np.array_equal(_host_of(back[: int(bl.item())]), small)


@pytest.mark.parametrize("n_ary,gib", [(4, 4), (16, 4), (2, 4)])
def test_four_gib_per_gpu_more_than_2_to_the_32_bits(dc, oracle, n_ary, gib):
    """BASELINE configs[3] / [4] put 4 GiB on a GPU at G = 2: more than 2^32 bits of payload, more than 2^32 bytes of input.
    Planned encode == the oracle's payload (compared on the device, the oracle packs block-parallel on the host cores);
    decode of the ORACLE's stream == the input."""
    from data_compression_b200 import synth
    threads = os.cpu_count() or 1
    n = gib << 30
    thr, base = synth.zipf_bytes_spec()
    data = torch.empty(n, dtype=torch.uint8, device="cuda")
    dc.synth_fill(data, synth.SEED_BASE + 3, synth.device_thresholds(thr, "cuda"), base)
    host = _host_of(data)
    ws = dc.encode_workspace(n, "cuda")
    hist = dc.histogram_runs(data, ws)
    o_hist = oracle.histogram_u8(host, threads=threads)
    assert np.array_equal(hist.cpu().numpy().astype(np.uint64), o_hist)
    table = dc.huff_build(hist, n_ary)
    ln, el, ev, st = oracle.build_tables(o_hist, n_ary)
    assert st == 0
    res = dc.huff_encode(data, table, workspace=ws, planned=True)
    nbits = res.bits()
    assert nbits > 1 << 32
    want, wbits, _ = oracle.pack_mt(host, el, ev, oracle.bits_per_digit(n_ary), 0, threads=threads,
                                    out=np.empty(n + n // 4 + 64, dtype=np.uint8))
    del host
    assert nbits == wbits
    nbytes = (nbits + 7) // 8
    d_want = torch.from_numpy(want[: nbytes + 64]).cuda()
    del want
    assert torch.equal(res.payload[:nbytes], d_want[:nbytes])
    del res, ws
    torch.cuda.empty_cache()
    out = dc.huff_decompress(d_want, wbits, table, n)
    assert torch.equal(out, data)
