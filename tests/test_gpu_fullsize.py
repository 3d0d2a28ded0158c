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
