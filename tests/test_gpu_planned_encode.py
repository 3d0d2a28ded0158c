"""The planned encode path: dc_histogram_u8_runs -> dc_huff_build -> dc_huff_encode_planned (run offsets from the run
histograms K1 leaves in the workspace).  Must give the same bytes as the oracle -- and as dc_huff_encode -- for every table
class (codes up to 12 bits: planned single pass; 13..16: look-back single pass; longer: 64-bit entries), any phase, ragged
sizes, and report the same errors."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dc():
    import data_compression_b200 as m
    return m


@pytest.fixture(scope="module")
def oracle():
    from oracle import pyoracle as O
    O.build()
    return O


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).copy()).cuda()


def _planned(dc, data, table, phase=0, cap=None):
    ws = dc.encode_workspace(data.numel(), data.device)
    hist = dc.histogram_runs(data, ws)
    out = torch.full((cap if cap is not None else data.numel() * 4 + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    res = dc.huff_encode(data, table, out=out, bit_phase=phase, workspace=ws, planned=True)
    return hist, res, out


@pytest.mark.parametrize("n_ary", [2, 4, 16, 3])
def test_planned_matches_oracle(dc, oracle, n_ary):
    from data_compression_b200 import synth
    thr, base = synth.zipf_bytes_spec()
    host = synth.host_stream((1 << 20) + 12345, synth.SEED_BASE + 9, thr, base)
    big = _dev(host)
    hist = dc.histogram(big)
    table = dc.huff_build(hist, n_ary)
    ln, el, ev, st = oracle.build_tables(hist.cpu().numpy().astype(np.uint64), n_ary)
    bpd = oracle.bits_per_digit(n_ary) if n_ary != 3 else 2
    for size in (1, 17, 2047, 2048, 2049, 32767, 32768, 32769, 65536 + 5, host.size):
        for phase in (0, 6):
            data = big[:size]
            h2, res, out = _planned(dc, data, table, phase)
            assert np.array_equal(h2.cpu().numpy()[:256], np.bincount(host[:size], minlength=256)), (size, "run-histogram pass")
            nbits = res.bits()
            ref = dc.huff_encode(data, table, out=torch.empty(size * 4 + 64, dtype=torch.uint8, device="cuda"), bit_phase=phase)
            assert nbits == ref.bits(), (n_ary, size, phase)
            nb = (nbits + phase + 7) // 8
            assert torch.equal(out[:nb], ref.payload[:nb]), (n_ary, size, phase)
            assert (out[nb:nb + 64] == 0xEE).all(), "bytes behind the stream were touched"
            if n_ary != 3:
                want, wbits = oracle.pack(host[:size], el, ev, bpd, phase)
                assert nbits == wbits and np.array_equal(out[:nb].cpu().numpy(), want), (n_ary, size, phase)


@pytest.mark.parametrize("depths,n_ary", [(13, 2), (8, 4), (15, 4)])
def test_planned_long_codes(dc, oracle, depths, n_ary):
    """13..16-bit tables (look-back kernel) and 30-bit tables (64-bit entries, chunk offsets counted inside the kernel)."""
    ln = np.zeros(259, dtype=np.int32)
    sym = 1
    for depth in range(1, depths):
        for _ in range(n_ary - 1):
            if sym < 250:
                ln[sym] = depth; sym += 1
    for _ in range(n_ary):
        if sym < 256:
            ln[sym] = depths; sym += 1
    bpd = oracle.bits_per_digit(n_ary)
    el, ev, st = oracle.convert_lengths_to_encode_table(ln, n_ary)
    assert st == 0
    table = dc.huff_table_from_lengths(_dev(ln), n_ary)
    rng = np.random.default_rng(depths)
    used = np.flatnonzero(ln)
    w = 0.5 ** (ln[used] * bpd / 2.5)
    for size in (70001, 32768 * 3 + 100):
        data = rng.choice(used, size=size, p=w / w.sum()).astype(np.uint8)
        _, res, out = _planned(dc, _dev(data), table, phase=3)
        nbits = res.bits()
        want, wbits = oracle.pack(data, el, ev, bpd, 3)
        assert nbits == wbits
        assert np.array_equal(out[: (nbits + 3 + 7) // 8].cpu().numpy(), want)
        back, status = dc.huff_decode(out, nbits, table, size, bit_start=3)
        assert int(status.item()) == 0 and np.array_equal(back.cpu().numpy(), data)


def test_planned_errors(dc):
    data = _dev(np.array([65] * 5000 + [66] + [65] * 3000, dtype=np.uint8))
    table = dc.huff_build(dc.histogram(data[:5000]), 4)   # 66 has no code
    _, res, _ = _planned(dc, data, table)
    assert int(res.status.item()) == dc.DC_ERR_SYMBOL
    ok = _dev(np.array([65, 66] * 40000, dtype=np.uint8))
    table = dc.huff_build(dc.histogram(ok), 2)
    _, res, out = _planned(dc, ok, table)
    nb = (res.bits() + 7) // 8
    _, res2, small = _planned(dc, ok, table, cap=nb - 1)
    assert int(res2.status.item()) == dc.DC_ERR_CAPACITY


def test_radix3_default_payload_buffer_holds_uniform_bytes(dc):
    """near-uniform bytes need 10.2 bits per symbol on the 2-bit-per-trit stream; the default buffer must hold them."""
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    data = torch.randint(0, 256, (300000,), dtype=torch.uint8, device="cuda", generator=g)
    payload, nbits, table = dc.huff_compress(data, 3)
    assert nbits > 8 * data.numel()
    back, st = dc.huff_decode(payload, nbits, table, data.numel())
    assert int(st.item()) == 0 and torch.equal(back, data)


def test_trit_pack_ignores_bits_behind_the_last_trit(dc):
    """79 trits fill 20 bytes of the 2-bit stream except the last field; whatever it holds is padding."""
    t2 = torch.zeros(32, dtype=torch.uint8, device="cuda")
    t2[:20] = 0b01100001          # trits 1, 2, 0, 1 ...
    a, sa = dc.trit_pack(t2.clone(), 79)
    t2[19] |= 0b11                # a field of 3 behind the last trit
    b, sb = dc.trit_pack(t2, 79)
    assert int(sa.item()) == 0 and int(sb.item()) == 0 and torch.equal(a, b)
