"""GPU: what round 2 added around the kernels -- the host-visible table header (and a table the library did not build), the
byte-stepped decoder's corner tables (binary codes of all 256 byte values: 256 states, F1 state machine + window F3; streams
that start inside a byte: entry rows), the TMA-staged variant of F1, and the encoder's dispatch with and without a known header."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _bytes_all_256(n, seed=11):
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, 257) ** 1.05
    return rng.choice(np.arange(256, dtype=np.uint8), size=n, p=w / w.sum()).astype(np.uint8)


@pytest.mark.parametrize("n_ary", [2, 4, 16])
def test_full_byte_alphabet_roundtrip_matches_oracle(dc, oracle, n_ary):
    """All 256 byte values (0x00 included): 257 leaves with the dummy, i.e. exactly 256 internal nodes at n = 2."""
    host = _bytes_all_256((1 << 20) + 77)
    host[:256] = np.arange(256, dtype=np.uint8)
    data = torch.from_numpy(host).cuda()
    hist = dc.histogram(data)
    table = dc.huff_build(hist, n_ary)
    t = table.download()
    ln, el, ev, st = oracle.build_tables(hist.cpu().numpy().astype(np.uint64), n_ary)
    if n_ary == 2:
        assert t.fsm_states == 256   # one more than the write pass can name: F1 alone takes the state machine
    for phase in (0, 1, 6):
        res = dc.huff_encode(data, table, bit_phase=phase)
        nbits = res.bits()
        want, wbits = oracle.pack(host, el, ev, oracle.bits_per_digit(n_ary), phase)
        assert nbits == wbits
        assert np.array_equal(res.payload[: (nbits + phase + 7) // 8].cpu().numpy(), want)
        out, status = dc.huff_decode(res.payload, nbits, table, host.size, bit_start=phase)
        assert int(status.item()) == 0 and torch.equal(out, data), (n_ary, phase)


def test_decode_with_a_table_the_library_did_not_build(dc, oracle):
    """A table copied into a fresh device buffer has no host-visible header: the decoder reads it back (blocking) instead."""
    from data_compression_b200.api import HuffTable
    host = _bytes_all_256(300007, seed=5)
    data = torch.from_numpy(host).cuda()
    table = dc.huff_build(dc.histogram(data), 4)
    res = dc.huff_encode(data, table)
    nbits = res.bits()
    foreign = HuffTable(data.device)
    foreign.buf.copy_(table.buf)          # a plain device copy: the library has never seen this address
    foreign.n_ary = 4
    out, status = dc.huff_decode(res.payload, nbits, foreign, host.size)
    assert int(status.item()) == 0 and torch.equal(out, data)
    res2 = dc.huff_encode(data, foreign)   # and the encoder, which then lets its kernels pick by the table they find
    assert res2.bits() == nbits and torch.equal(res2.payload[: (nbits + 7) // 8], res.payload[: (nbits + 7) // 8])


def test_encoder_dispatch_known_and_unknown_header(dc, oracle):
    """The same payload whether the host already knows the table's longest code (build finished) or not (still queued)."""
    host = _bytes_all_256(1 << 20, seed=9)
    data = torch.from_numpy(host).cuda()
    hist = dc.histogram(data)
    for n_ary in (2, 3, 4):
        table = dc.huff_build(hist, n_ary)
        torch.cuda.synchronize()                    # the build's event has passed: one matching launch
        a = dc.huff_encode(data, table, out=torch.empty(host.size * 2 + 64, dtype=torch.uint8, device="cuda"))
        na = a.bits()
        filler = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
        filler.random_(0, 255)                      # something long in front of the build ...
        table2 = dc.huff_build(hist, n_ary)         # ... so its header is not there yet when the encoder asks
        b = dc.huff_encode(data, table2, out=torch.empty(host.size * 2 + 64, dtype=torch.uint8, device="cuda"))
        nb = b.bits()
        assert na == nb and torch.equal(a.payload[: (na + 7) // 8], b.payload[: (nb + 7) // 8]), n_ary


def test_tma_staged_sync_pass_is_bit_exact():
    """north_star (4): DC_DECODE_TMA=1 stages the bitstream tiles with cp.async.bulk + mbarrier; same results."""
    code = r'''
import numpy as np, torch, sys
sys.path.insert(0, %r)
import data_compression_b200 as dc
from data_compression_b200 import synth
thr, base = synth.zipf_bytes_spec()
for n in (1, 4097, (1 << 22) + 13):
    data = torch.empty(n, dtype=torch.uint8, device="cuda")
    dc.synth_fill(data, synth.SEED_BASE + n, synth.device_thresholds(thr, "cuda"), base)
    hist = dc.histogram(data)
    for n_ary in (2, 3, 4, 16):   # n = 2: the state machine for F1 only (COMPAT), window kernel behind it
        table = dc.huff_build(hist, n_ary)
        for phase in (0, 2):
            res = dc.huff_encode(data, table, bit_phase=phase, out=torch.empty(n * 2 + 64, dtype=torch.uint8, device="cuda"))
            out, st = dc.huff_decode(res.payload, res.bits(), table, n, bit_start=phase)
            assert int(st.item()) == 0 and torch.equal(out, data), (n, n_ary, phase)
print("tma ok")
''' % ROOT
    env = dict(os.environ, DC_DECODE_TMA="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "tma ok" in r.stdout, r.stdout + r.stderr


def test_stale_table_header_is_caught_and_forget_clears_it(dc, oracle):
    """A built table overwritten behind the library's back (the contract asks for dc_huff_table_forget): the header the host
    kept is another table's.  The decoder's state-machine kernels refuse it and the window kernels -- which only trust the
    table itself -- redo the stream; the encoder, launched for the wrong code lengths, reports DC_ERR_ARG.  After forget()
    everything takes its normal path."""
    from data_compression_b200.api import HuffTable
    host = _bytes_all_256(200003, seed=21)
    data = torch.from_numpy(host).cuda()
    hist = dc.histogram(data)
    a = dc.huff_build(hist, 4)                       # the library now knows a header for a.buf's address
    assert a.download().max_bits > 12                # (256 symbols at n = 4: codes of 7 digits)
    # another code: binary, codes of at most 12 bits (a table from lengths) -- it wants the other encoder instantiation
    ln = np.zeros(259, dtype=np.int32)
    s = 1
    for depth in range(1, 12):
        ln[s] = depth
        s += 1
    ln[s] = 12; ln[s + 1] = 12
    nsym = s + 2
    b = dc.huff_table_from_lengths(torch.from_numpy(ln).cuda(), 2)
    small = torch.from_numpy(np.random.default_rng(3).integers(1, nsym, 100001).astype(np.uint8)).cuda()
    good = dc.huff_encode(small, b, out=torch.empty(small.numel() * 2 + 64, dtype=torch.uint8, device="cuda"))
    nbits = good.bits()
    torch.cuda.synchronize()
    a.buf.copy_(b.buf)                               # behind the library's back
    a.n_ary = 2
    torch.cuda.synchronize()
    out, status = dc.huff_decode(good.payload, nbits, a, small.numel())
    assert int(status.item()) == 0 and torch.equal(out, small)      # slow path, right answer
    bad = dc.huff_encode(small, a, out=torch.empty(small.numel() * 2 + 64, dtype=torch.uint8, device="cuda"))
    assert int(bad.status.item()) == dc.DC_ERR_ARG
    assert dc.lib().dc_huff_table_forget(a.ptr) == 0
    again = dc.huff_encode(small, a, out=torch.empty(small.numel() * 2 + 64, dtype=torch.uint8, device="cuda"))
    assert again.bits() == nbits and torch.equal(again.payload[: (nbits + 7) // 8], good.payload[: (nbits + 7) // 8])
    out, status = dc.huff_decode(again.payload, nbits, a, small.numel())
    assert int(status.item()) == 0 and torch.equal(out, small)


@pytest.mark.parametrize("n_ary", [2, 3, 4, 16])
def test_indexed_decode_matches_blind_decode(dc, oracle, n_ary):
    """Opt-in index (what F1 + F2 find, kept by the producer): the write pass alone reproduces the input."""
    for n, seed in ((1, 1), (4097, 2), ((1 << 21) + 5, 3)):
        host = _bytes_all_256(n, seed=seed)
        data = torch.from_numpy(host).cuda()
        table = dc.huff_build(dc.histogram(data), n_ary)
        for phase in (0, 4):
            res = dc.huff_encode(data, table, bit_phase=phase, out=torch.empty(n * 2 + 64, dtype=torch.uint8, device="cuda"))
            nbits = res.bits()
            built = dc.huff_index_build(res.payload, nbits, table, n, bit_start=phase)
            assert built is not None
            index, info = built
            assert index.numel() >= dc.lib().dc_huff_index_bytes(phase, nbits) and index.numel() <= (nbits // 8) // 14 + 4096
            out, status = dc.huff_decode_indexed(res.payload, index, info, table)
            assert int(status.item()) == 0 and torch.equal(out, data), (n_ary, n, phase)
    # an index that belongs to another table is refused where the host can tell
    other = dc.huff_build(dc.histogram(data), 16 if n_ary != 16 else 4)
    with pytest.raises(dc.DcError):
        dc.huff_decode_indexed(res.payload, index, info, other)


@pytest.mark.parametrize("n_ary", [5, 9, 10, 12, 15])
def test_nibble_per_digit_radices(dc, oracle, n_ary):
    """SURVEY N4 (n = 9, 10 by name): radices 5 .. 15 pack one nibble per digit, most significant digit first.  The payload
    equals the oracle's (its packer fed with the canonical base-n values rewritten as nibbles by an independent routine), a
    pure-Python digit reader gets the input back from it, and the GPU decoder -- state machine only -- round-trips it."""
    for n, seed in ((1, 1), (257, 2), (4099, 3), ((1 << 20) + 9, 4)):
        host = _bytes_all_256(n, seed=seed)
        data = torch.from_numpy(host).cuda()
        hist = dc.histogram(data)
        table = dc.huff_build(hist, n_ary)
        t = table.download()
        assert t.bits_per_digit == 4 and t.packed_radix == 0 and t.fsm_states > 0
        ln, el, ev, st = oracle.build_tables(hist.cpu().numpy().astype(np.uint64), n_ary)
        assert st == 0 and np.array_equal(np.array(t.lengths[:259]), ln)
        nv = oracle.nibble_values(el, ev, n_ary)
        for phase in (0, 4):
            res = dc.huff_encode(data, table, bit_phase=phase)
            nbits = res.bits()
            want, wbits = oracle.pack(host, el, nv, 4, phase)
            assert nbits == wbits and nbits % 4 == 0
            got = res.payload[: (nbits + phase + 7) // 8].cpu().numpy()
            assert np.array_equal(got, want), (n_ary, n, phase)
            if n <= 4099:
                assert np.array_equal(oracle.unpack_nibble_digits(got, phase, nbits, el, ev, n_ary), host)
            out, status = dc.huff_decode(res.payload, nbits, table, n, bit_start=phase)
            assert int(status.item()) == 0 and torch.equal(out, data), (n_ary, n, phase)
    # a nibble that is not a digit of the radix
    bad = res.payload.clone()
    bad[1000] = 0xFF
    out, status = dc.huff_decode(bad, nbits, table, n, bit_start=4)
    assert int(status.item()) == dc.DC_ERR_CORRUPT
    # off the nibble grid: refused
    with pytest.raises(dc.DcError):
        dc.huff_decode(res.payload, nbits, table, n, bit_start=2)


def test_nibble_radix_stream_that_does_not_self_synchronise(dc, oracle):
    """124 equally likely symbols at n = 5 (the as-written dummy rule adds the 125th leaf): every code has 3 digits, so a decoder
    that starts in the wrong place never finds its way back.  The window kernels' robust path has no tables for such a radix:
    one thread walks the stream instead."""
    rng = np.random.default_rng(8)
    host = rng.integers(1, 125, size=300001).astype(np.uint8)
    host[:124] = np.arange(1, 125, dtype=np.uint8)
    data = torch.from_numpy(host).cuda()
    hist = dc.histogram(data)
    hist[1:125] = 1000                                # exactly uniform: a complete 5-ary tree of depth 3 (one leaf is the dummy)
    table = dc.huff_build(hist, 5)
    t = table.download()
    assert t.min_len == 3 and t.max_len == 3
    res = dc.huff_encode(data, table)
    nbits = res.bits()
    assert nbits == host.size * 12
    out, status = dc.huff_decode(res.payload, nbits, table, host.size)
    assert int(status.item()) == 0 and torch.equal(out, data)
    assert dc.huff_index_build(res.payload, nbits, table, host.size) is None     # no index for such a stream
    # the host entry points, the reference's call shape: huffman(.., 10, ..) -> compress -> decompress
    from data_compression_b200 import hostapi
    payload, bits, lengths = hostapi.huff_compress(host, 10)
    back = hostapi.huff_decompress(payload, bits, lengths, 10, host.size)
    assert np.array_equal(back, host)
