"""GPU: the host-pointer entry points (what a C caller of the reference's functions links against)
reproduce the reference's own known-answer tests and golden vectors."""
import numpy as np
import pytest

from conftest import dense, load_golden

pytestmark = pytest.mark.gpu


def test_histogram_host(dc, oracle):
    g = load_golden("histogram.json")
    h = dc.hostapi.histogram(g["text"].encode())
    assert h.tolist() == g["hist"]          # includes the zeroed canary slot 258
    assert dc.hostapi.histogram(b"ab\x00cd")[ord("c")] == 0   # NUL terminates, like while(*c) :482
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=123457, dtype=np.uint8)
    assert np.array_equal(dc.hostapi.histogram_u8(a), oracle.histogram_u8(a))


def test_huffman_host_matches_reference_golden(dc, table_cases):
    for case in table_cases[:12]:
        h = dense(case["hist"])
        for n in (2, 3, 4, 10, 16):
            want = dense(case["radix"][str(n)]["lengths"], dtype=np.int32)
            assert np.array_equal(dc.hostapi.huffman(h.astype(np.int32), n), want), (case["name"], n)
            assert np.array_equal(dc.hostapi.huffman(h.astype(np.uint64), n), want), (case["name"], n)


def test_huffman_host_general_alphabet(dc, oracle):
    rng = np.random.default_rng(8)
    for mlv in (4, 20, 122, 300, 511):
        f = rng.integers(0, 50, size=mlv + 1).astype(np.int32)
        for n in (2, 3, 7):
            assert np.array_equal(dc.hostapi.huffman(f, n), oracle.huffman(f.astype(np.uint64), n)), (mlv, n)


def test_convert_kats(dc):
    # the reference's own KATs (n_ary_huffman.c:2821-2891) and the last-slot quirk, from the unmodified function
    for k in load_golden("convert_kats.json")["cases"]:
        pre = k.get("prefill", [0, 0])
        n = len(k["lengths"])
        el, ev, st = dc.hostapi.convert_lengths_to_encode_table(
            np.array(k["lengths"], dtype=np.int32), k["n"], max_symbol_value=k["max_symbol_value"],
            elen=np.full(n, pre[0], dtype=np.int32), evalue=np.full(n, pre[1], dtype=np.uint32))
        msv = k["max_symbol_value"]
        assert st == 0
        assert el[: msv + 1].tolist() == k["elen"][: msv + 1]
        assert ev[: msv + 1].tolist() == k["evalue"][: msv + 1]


def test_represent_items_with_codes_host(dc, oracle):
    text = (b"/* n_ary_huffman.c" b"2021-10-25: started by David Cary") * 40
    hist = oracle.histogram_u8(text)
    for n in (2, 4, 16):
        ln, el, ev, st = oracle.build_tables(hist, n)
        out, written, bits = dc.hostapi.represent_items_with_codes(ln, n, text, bufsize=65000, start=7)
        want, wbits = oracle.pack(text, el, ev, oracle.bits_per_digit(n))
        assert bits == wbits and written == want.size
        assert not out[:7].any() and np.array_equal(out[7: 7 + written], want)


def test_compress_decompress_host(dc, oracle):
    from data_compression_b200 import synth
    thr, base = synth.zipf_7bit_spec()
    data = synth.host_stream(1 << 20, synth.SEED_BASE + 1, thr, base)   # config 1 input
    for n in (2, 4, 16):
        payload, bits, lengths = dc.hostapi.huff_compress(data, n)
        ln, el, ev, st = oracle.build_tables(oracle.histogram_u8(data), n)
        assert np.array_equal(lengths, ln)
        want, wbits = oracle.pack(data, el, ev, oracle.bits_per_digit(n))
        assert bits == wbits and np.array_equal(payload, want)
        assert np.array_equal(dc.hostapi.huff_decompress(payload, bits, lengths, n, data.size), data)


def test_compress_decompress_host_radix3(dc, oracle):
    """The reference's default radix through the host entry points: 5 trits per byte, against the oracle."""
    from data_compression_b200 import synth
    thr, base = synth.zipf_7bit_spec()
    for size in (1, 7, 4096, 300001):
        data = synth.host_stream(size, synth.SEED_BASE + 9, thr, base)
        payload, bits, lengths = dc.hostapi.huff_compress(data, 3)
        ln, el, ev, st = oracle.build_tables(oracle.histogram_u8(data), 3)
        want, wtrits = oracle.pack_trits(data, el, ev)
        assert np.array_equal(lengths, ln) and bits == 2 * wtrits and np.array_equal(payload, want)
        assert payload.min() >= 1 and payload.max() <= 243            # "never uses byte 0 or 244..255" (:748)
        assert np.array_equal(dc.hostapi.huff_decompress(payload, bits, lengths, 3, data.size), data)


@pytest.mark.parametrize("n_ary", [2, 16])
def test_decompress_host_pipelined(dc, oracle, n_ary):
    """A stream of more than three 32 MiB chunks takes the chunked, overlapped path of dc_host_huff_decompress: the
    chunks are chained on the device (first-code offset, output offset) and must reassemble the input exactly; a
    wrong symbol count and a truncated stream are still reported."""
    import torch
    from data_compression_b200 import synth
    thr, base = synth.zipf_bytes_spec()
    n = 150 * (1 << 20) + 12345
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    dc.synth_fill(d, 991, synth.device_thresholds(thr, "cuda"), base)
    data = d.cpu().numpy()
    del d
    payload, bits, lengths = dc.hostapi.huff_compress(data, n_ary)
    assert payload.size > 3 * (32 << 20)
    # spot-check the payload against the oracle on the first MiB of symbols (same table, prefix of the stream)
    ln, el, ev, st = oracle.build_tables(oracle.histogram_u8(data), n_ary)
    assert np.array_equal(lengths, ln)
    want, wbits = oracle.pack(data[: 1 << 20], el, ev, oracle.bits_per_digit(n_ary))
    assert np.array_equal(payload[: wbits // 8], want[: wbits // 8])
    # the whole payload against the oracle's single stream
    import os
    full, fbits, _ = oracle.pack_mt(data, el, ev, oracle.bits_per_digit(n_ary), 0, threads=os.cpu_count() or 1,
                                    out=np.empty(n + n // 4 + 64, dtype=np.uint8))
    assert bits == fbits and np.array_equal(payload, full[: (fbits + 7) // 8])
    back = dc.hostapi.huff_decompress(payload, bits, lengths, n_ary, n)
    assert np.array_equal(back, data)
    with pytest.raises(dc.DcError) as e:
        dc.hostapi.huff_decompress(payload, bits, lengths, n_ary, n - 1)
    assert e.value.status in (dc.DC_ERR_CAPACITY, dc.DC_ERR_CORRUPT)
    with pytest.raises(dc.DcError):
        dc.hostapi.huff_decompress(payload, bits - 4099, lengths, n_ary, n)


def test_nybble_host(dc, oracle):
    for c in load_golden("nybble.json")["write_nybble"]:
        s = np.array(c["symbols"], dtype=np.uint8)
        assert dc.hostapi.nybble_pack(s).tolist() == c["packed"]
        assert dc.hostapi.nybble_unpack(np.array(c["packed"], dtype=np.uint8), s.size).tolist() == c["symbols"]


def test_radix3_host_pipelined(dc, oracle):
    """Radix 3 on host buffers at a size that takes the chunked decompress: every chunk of the 5-trits-per-byte payload is
    unpacked on the device and decoded while the next one is uploaded; payload against the oracle, then the round trip."""
    import torch
    from data_compression_b200 import synth
    thr, base = synth.zipf_bytes_spec()
    n = 140 * (1 << 20) + 777
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    dc.synth_fill(d, 4242, synth.device_thresholds(thr, "cuda"), base)
    data = d.cpu().numpy()
    del d
    payload, bits, lengths = dc.hostapi.huff_compress(data, 3)
    ln, el, ev, st = oracle.build_tables(oracle.histogram_u8(data), 3)
    want, wtrits = oracle.pack_trits(data, el, ev)
    assert st == 0 and np.array_equal(lengths, ln) and bits == 2 * wtrits and np.array_equal(payload, want)
    assert bits // 8 > 3 * 32768000                       # more than three chunks of the 2-bit-per-trit stream
    back = dc.hostapi.huff_decompress(payload, bits, lengths, 3, n)
    assert np.array_equal(back, data)
    broken = payload.copy()
    broken[payload.size // 2] = 250                       # not a payload byte (1..243)
    with pytest.raises(dc.DcError) as e:
        dc.hostapi.huff_decompress(broken, bits, lengths, 3, n)
    assert e.value.status == dc.DC_ERR_CORRUPT
