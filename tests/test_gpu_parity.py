"""GPU: bit-exact parity of every kernel with the CPU oracle / the reference-generated golden vectors,
called through the C-ABI (ctypes) exactly as a host program would."""
import numpy as np
import pytest
import torch

from conftest import dense

pytestmark = pytest.mark.gpu

RADICES = (2, 3, 4, 5, 10, 16)
PACKABLE = (2, 4, 16)


def _dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _zipf(dc, n, seed=3):
    from data_compression_b200 import synth
    thr, base = synth.zipf_bytes_spec()
    data = torch.empty(n, dtype=torch.uint8, device="cuda")
    dc.synth_fill(data, synth.SEED_BASE + seed, synth.device_thresholds(thr, "cuda"), base)
    return data


def test_synth_matches_host_twin(dc):
    from data_compression_b200 import synth
    for spec in (synth.zipf_bytes_spec, synth.zipf_7bit_spec, synth.zipf_nybble_spec):
        thr, base = spec()
        for n in (1, 17, 4096, 100003):
            d = torch.empty(n, dtype=torch.uint8, device="cuda")
            dc.synth_fill(d, 12345, synth.device_thresholds(thr, "cuda"), base)
            assert np.array_equal(d.cpu().numpy(), synth.host_stream(n, 12345, thr, base))


@pytest.mark.parametrize("variant", [0, 1])
def test_histogram_variants_match_oracle(dc, oracle, variant):
    rng = np.random.default_rng(variant)
    zipf = _zipf(dc, (1 << 22) + 5).cpu().numpy()
    cases = [zipf, rng.integers(0, 256, size=1 << 20, dtype=np.uint8), np.full(300001, 7, dtype=np.uint8),
             np.zeros(0, dtype=np.uint8)]
    cases += [rng.integers(0, 256, size=k, dtype=np.uint8) for k in (1, 15, 16, 17, 255, 7681, 65537)]
    for a in cases:
        got = dc.histogram(_dev(a) if a.size else torch.empty(0, dtype=torch.uint8, device="cuda"), variant=variant)
        assert np.array_equal(got.cpu().numpy().astype(np.uint64), oracle.histogram_u8(a)), (variant, a.size)
    # unaligned base pointer
    buf = _dev(zipf)
    got = dc.histogram(buf[3:100000], variant=variant)
    assert np.array_equal(got.cpu().numpy().astype(np.uint64), oracle.histogram_u8(zipf[3:100000]))


def test_histogram_zeroes_all_slots(dc):
    out = torch.full((259,), 0xBEEF, dtype=torch.int64, device="cuda")  # the reference's canary (:2663)
    dc.histogram(_dev(np.array([65, 66, 66], dtype=np.uint8)), out=out)
    h = out.cpu().numpy()
    assert h[65] == 1 and h[66] == 2 and h.sum() == 3 and h[258] == 0


def test_tables_match_reference_golden(dc, oracle, table_cases):
    """lengths and canonical values equal the UNMODIFIED reference's, for every golden histogram and radix."""
    for case in table_cases:
        hist = _dev(dense(case["hist"]))
        for n in RADICES:
            rec = case["radix"][str(n)]
            t = dc.huff_build(hist, n).download()
            want_len = dense(rec["lengths"], dtype=np.int32)
            assert np.array_equal(np.array(t.lengths[:259], dtype=np.int32), want_len), (case["name"], n)
            nz = int((dense(case["hist"]) != 0).sum())
            assert t.nonzero_symbols == nz
            assert t.dummy_nodes == (n - 1) - int(np.fmod(nz - 1, n - 1))  # as written (:900-903)
            if "values" in rec:
                assert t.status == 0
                assert np.array_equal(np.array(t.values[:259], dtype=np.uint32), dense(rec["values"], dtype=np.uint32)), \
                    (case["name"], n)
                # radix 3: one 2-bit field per trit in the kernels' stream; 5 .. 15: one nibble per digit
                bpd = 2 if n == 3 else 4 if 5 <= n < 16 else oracle.bits_per_digit(n)
                assert t.packed_radix == (3 if n == 3 else 0)
                assert t.bits_per_digit == bpd and t.max_bits == int(want_len.max()) * bpd
                assert t.total_bits == int((dense(case["hist"]) * want_len * bpd).sum())
            else:
                assert t.status == dc.DC_ERR_CODE_TOO_LONG


def test_tables_random_differential(dc, oracle):
    rng = np.random.default_rng(99)
    for t in range(40):
        h = np.zeros(259, dtype=np.int64)
        k = int(rng.integers(1, 257))
        idx = rng.choice(np.arange(0, 256), size=k, replace=False)
        h[idx] = rng.integers(1, 6, size=k) if t % 2 else rng.integers(1, 1 << 40, size=k)
        for n in RADICES:
            tab = dc.huff_build(_dev(h), n).download()
            ln = oracle.huffman(h.astype(np.uint64), n)
            assert np.array_equal(np.array(tab.lengths[:259], dtype=np.int32), ln), (t, n)
            el, ev, st = oracle.convert_lengths_to_encode_table(ln, n)
            if st == 0:
                assert tab.status == 0
                assert np.array_equal(np.array(tab.values[:259], dtype=np.uint32), ev), (t, n)
            else:
                assert tab.status == dc.DC_ERR_CODE_TOO_LONG


def test_table_from_lengths_and_lut(dc, oracle, table_cases):
    case = next(c for c in table_cases if c["name"] == "zipf_2p30")
    for n in PACKABLE:
        ln = dense(case["radix"][str(n)]["lengths"], dtype=np.int32)
        t = dc.huff_table_from_lengths(_dev(ln), n).download()
        el, ev, st = oracle.convert_lengths_to_encode_table(ln, n)
        assert np.array_equal(np.array(t.values[:259], dtype=np.uint32), ev)
        bpd = oracle.bits_per_digit(n)
        lut = np.array(t.lut, dtype=np.uint16)
        for s in np.nonzero(ln)[0]:
            nb = int(ln[s]) * bpd
            if nb <= 12:
                lo = int(ev[s]) << (12 - nb)
                assert (lut[lo: lo + (1 << (12 - nb))] == ((nb << 8) | int(s))).all()


@pytest.mark.parametrize("n_ary", PACKABLE)
def test_encode_matches_oracle_and_roundtrips(dc, oracle, n_ary):
    big = _zipf(dc, (1 << 20) + 3)
    host = big.cpu().numpy()
    hist = dc.histogram(big)
    table = dc.huff_build(hist, n_ary)
    ln, el, ev, st = oracle.build_tables(hist.cpu().numpy().astype(np.uint64), n_ary)
    bpd = oracle.bits_per_digit(n_ary)
    for size in (1, 2, 15, 16, 17, 255, 4095, 4096, 4097, 8192, 12289, host.size):
        for phase in ((0, 5) if size < host.size else (0, 3)):
            res = dc.huff_encode(big[:size], table, bit_phase=phase)
            nbits = res.bits()
            want, wbits = oracle.pack(host[:size], el, ev, bpd, phase)
            assert nbits == wbits, (size, phase)
            got = res.payload[: (nbits + phase + 7) // 8].cpu().numpy()
            assert np.array_equal(got, want), (n_ary, size, phase)
            out, status = dc.huff_decode(res.payload, nbits, table, size, bit_start=phase)
            assert int(status.item()) == 0, (n_ary, size, phase)
            assert torch.equal(out, big[:size]), (n_ary, size, phase)


def test_encode_only_writes_its_bytes(dc, oracle):
    data = _zipf(dc, 50000)
    table = dc.huff_build(dc.histogram(data), 2)
    out = torch.full((60000,), 0xAB, dtype=torch.uint8, device="cuda")
    res = dc.huff_encode(data, table, out=out)
    nb = (res.bits() + 7) // 8
    assert (out[nb:] == 0xAB).all()
    # capacity one byte short -> DC_ERR_CAPACITY, nothing beyond the buffer touched
    small = torch.full((nb + 15,), 0xCD, dtype=torch.uint8, device="cuda")
    res = dc.huff_encode(data, table, out=small[: nb - 1])
    assert int(res.status.item()) == dc.DC_ERR_CAPACITY
    assert (small[nb - 1:] == 0xCD).all()


def test_encode_symbol_without_code(dc):
    data = _dev(np.array([65] * 100 + [66], dtype=np.uint8))
    table = dc.huff_build(dc.histogram(data[:100]), 2)  # 66 has no code
    res = dc.huff_encode(data, table)
    assert int(res.status.item()) == dc.DC_ERR_SYMBOL


def test_wide_codes_up_to_30_bits(dc, oracle):
    """n=4 with 15-digit (30-bit) codes: the enc64 path of the encoder and the canonical slow path of the decoder."""
    ln = np.zeros(259, dtype=np.int32)
    # a complete 4-ary code: 3 symbols at each depth 1..14, 4 at depth 15 (Kraft sum == 1)
    sym = 1
    for depth in range(1, 15):
        for _ in range(3):
            ln[sym] = depth; sym += 1
    for _ in range(4):
        ln[sym] = 15; sym += 1
    el, ev, st = oracle.convert_lengths_to_encode_table(ln, 4)
    assert st == 0
    table = dc.huff_table_from_lengths(_dev(ln), 4)
    t = table.download()
    assert t.status == 0 and t.max_bits == 30
    assert np.array_equal(np.array(t.values[:259], dtype=np.uint32), ev)
    rng = np.random.default_rng(4)
    data = rng.integers(1, sym, size=70001).astype(np.uint8)
    res = dc.huff_encode(_dev(data), table, out=torch.empty(data.size * 4 + 64, dtype=torch.uint8, device="cuda"))
    nbits = res.bits()
    want, wbits = oracle.pack(data, el, ev, 2)
    assert nbits == wbits
    assert np.array_equal(res.payload[: (nbits + 7) // 8].cpu().numpy(), want)
    out, status = dc.huff_decode(res.payload, nbits, table, data.size)
    assert int(status.item()) == 0 and np.array_equal(out.cpu().numpy(), data)


@pytest.mark.parametrize("n_ary,depths", [(2, 13), (2, 14), (2, 15), (4, 7), (4, 8), (16, 4)])
def test_mid_codes_13_to_16_bits(dc, oracle, n_ary, depths):
    """Tables whose longest code is 13..16 bits: the 16-bit instantiation of the single-pass encoder and the escape
    path of the decoder (codes longer than the 12 LUT index bits).  Skewed data so the long codes are rare but present,
    plus one ragged tail."""
    ln = np.zeros(259, dtype=np.int32)
    sym = 1
    for depth in range(1, depths):           # n-1 symbols at each depth, n at the last: a complete code
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
    t = table.download()
    assert t.status == 0 and 12 < t.max_bits <= 16, t.max_bits
    rng = np.random.default_rng(depths * 31 + n_ary)
    used = np.flatnonzero(ln)
    w = 0.5 ** (ln[used] * bpd / 2.0)
    for size in (200003, 4096 * 8 * 3 + 17):
        data = rng.choice(used, size=size, p=w / w.sum()).astype(np.uint8)
        res = dc.huff_encode(_dev(data), table, out=torch.empty(data.size * 2 + 64, dtype=torch.uint8, device="cuda"))
        nbits = res.bits()
        want, wbits = oracle.pack(data, el, ev, bpd)
        assert nbits == wbits
        assert np.array_equal(res.payload[: (nbits + 7) // 8].cpu().numpy(), want)
        out, status = dc.huff_decode(res.payload, nbits, table, data.size)
        assert int(status.item()) == 0 and np.array_equal(out.cpu().numpy(), data)


def test_radix3_stream_end_on_a_byte_boundary_with_garbage_behind(dc, oracle):
    """Radix 3 look-ups index by all eight 2-bit fields of their window: bytes behind the end of the stream (here 0xFF = fields
    of 3) must not reach the index.  Found by tools/fuzz.py: streams that end on a byte boundary in a buffer with garbage."""
    rng = np.random.default_rng(99)
    for size in (7, 100, 4097, 70001):
        alphabet = rng.choice(np.arange(1, 256), size=40, replace=False).astype(np.uint8)
        data = rng.choice(alphabet, size=size).astype(np.uint8)
        hist = oracle.histogram_u8(data)
        ln, el, ev, st = oracle.build_tables(hist, 3)
        trits = np.cumsum(ln[data])
        keep = int(np.flatnonzero(trits % 4 == 0)[-1]) + 1      # a prefix whose stream ends on a byte boundary
        data = data[:keep]
        d = _dev(data)
        table = dc.huff_table_from_lengths(_dev(ln.astype(np.int32)), 3)
        res = dc.huff_encode(d, table, out=torch.empty(keep * 4 + 64, dtype=torch.uint8, device="cuda"))
        nbits = res.bits()
        assert nbits % 8 == 0 and nbits == 2 * int(trits[keep - 1])
        nb = nbits // 8
        for fill in (0xFF, 0xAA, 0x00):
            buf = torch.full((nb + 4096,), fill, dtype=torch.uint8, device="cuda")
            buf[:nb] = res.payload[:nb]
            out, status = dc.huff_decode(buf, nbits, table, keep)
            assert int(status.item()) == 0 and torch.equal(out, d), (size, fill)


@pytest.mark.parametrize("depths", [9, 12, 15])
def test_radix3_long_codes(dc, oracle, depths):
    """Radix 3 with codes of more than 8 trits: the trit-indexed tables hold no entry for them and the decoder takes the
    canonical search (the T2 + ESC instantiations); 15 trits = 30 bits is the reference's limit (lengths < 16 digits)."""
    ln = np.zeros(259, dtype=np.int32)
    sym = 1
    for depth in range(1, depths):            # two symbols at each depth, three at the last: a complete ternary code
        for _ in range(2):
            ln[sym] = depth; sym += 1
    for _ in range(3):
        ln[sym] = depths; sym += 1
    el, ev, st = oracle.convert_lengths_to_encode_table(ln, 3)
    assert st == 0
    table = dc.huff_table_from_lengths(_dev(ln), 3)
    t = table.download()
    assert t.status == 0 and t.packed_radix == 3 and t.max_bits == 2 * depths
    rng = np.random.default_rng(depths)
    used = np.flatnonzero(ln)
    w = 0.5 ** (ln[used] * 0.8)
    for size in (150001, 4096 * 8 * 2 + 5):
        data = rng.choice(used, size=size, p=w / w.sum()).astype(np.uint8)
        assert ln[data].max() == depths      # the longest codes do occur
        res = dc.huff_encode(_dev(data), table, out=torch.empty(data.size * 4 + 64, dtype=torch.uint8, device="cuda"))
        nbits = res.bits()
        want, wtrits = oracle.pack_trits(data, el, ev)
        assert nbits == 2 * wtrits
        payload, pst = dc.trit_pack(res.payload, wtrits)
        assert int(pst.item()) == 0 and np.array_equal(payload.cpu().numpy(), want)
        t2, ust = dc.trit_unpack(payload, wtrits)
        out, status = dc.huff_decode(t2, nbits, table, data.size)
        assert int(ust.item()) == 0 and int(status.item()) == 0 and np.array_equal(out.cpu().numpy(), data)


def test_base64url_text_form_of_the_binary_payload(dc, oracle):
    """Row N4, second half: the reference's unfinished packer emits the binary code 6 bits per character through int2digit()
    (n_ary_huffman.c:371-426, :1646-1671).  Against the oracle and, on whole bytes, against RFC 4648 itself."""
    import base64
    rng = np.random.default_rng(46)
    for nbits in (1, 5, 6, 7, 8, 95, 96, 97, 4096 * 8, 1000003, (1 << 23) + 5):
        nb = (nbits + 7) // 8
        raw = rng.integers(0, 256, size=nb, dtype=np.uint8)
        buf = torch.full((nb + 64,), 0xFF, dtype=torch.uint8, device="cuda")   # whatever lies behind the stream must not matter
        buf[:nb] = _dev(raw)
        chars = dc.base64url_pack(buf, nbits)
        want = oracle.base64url_pack(raw, nbits)
        assert bytes(chars.cpu().numpy()) == want, nbits
        if nbits % 8 == 0:
            assert want == base64.urlsafe_b64encode(raw.tobytes()).rstrip(b"=")
        back, st = dc.base64url_unpack(chars.clone(), nbits)
        assert int(st.item()) == 0 and np.array_equal(back[:nb].cpu().numpy(), oracle.base64url_unpack(want, nbits))
    # digit2int() also takes the RFC 4648 standard alphabet (:441-445); anything else is reported
    raw = rng.integers(0, 256, size=3000, dtype=np.uint8)
    std = np.frombuffer(base64.b64encode(raw.tobytes()), dtype=np.uint8)
    back, st = dc.base64url_unpack(_dev(std), 24000)
    assert int(st.item()) == 0 and np.array_equal(back[:3000].cpu().numpy(), raw)
    bad = std.copy(); bad[1234] = ord("!")
    _, st = dc.base64url_unpack(_dev(bad), 24000)
    assert int(st.item()) == dc.DC_ERR_CORRUPT
    # the real thing: a binary Huffman payload as text and back, then decoded
    data = _zipf(dc, 200001, seed=12)
    table = dc.huff_build(dc.histogram(data), 2)
    res = dc.huff_encode(data, table)
    nbits = res.bits()
    text = dc.base64url_pack(res.payload, nbits)
    assert text.numel() == (nbits + 5) // 6
    bits_back, st = dc.base64url_unpack(text.clone(), nbits)
    out, status = dc.huff_decode(bits_back, nbits, table, data.numel())
    assert int(st.item()) == 0 and int(status.item()) == 0 and torch.equal(out, data)


def test_degenerate_alphabets(dc, oracle):
    # one distinct symbol: 1-bit codes (binary), 128 symbols per 128-bit subsequence
    for n_ary in PACKABLE:
        data = np.full(100000, 200, dtype=np.uint8)
        d = _dev(data)
        table = dc.huff_build(dc.histogram(d), n_ary)
        res = dc.huff_encode(d, table)
        nbits = res.bits()
        ln, el, ev, st = oracle.build_tables(oracle.histogram_u8(data), n_ary)
        want, wbits = oracle.pack(data, el, ev, oracle.bits_per_digit(n_ary))
        assert nbits == wbits and np.array_equal(res.payload[: (nbits + 7) // 8].cpu().numpy(), want)
        assert torch.equal(dc.huff_decompress(res.payload, nbits, table, data.size), d)
    # all 256 byte values, equal counts (incl. 0x00): lengths 8..9 in binary
    data = np.tile(np.arange(256, dtype=np.uint8), 300)
    np.random.default_rng(1).shuffle(data)
    d = _dev(data)
    table = dc.huff_build(dc.histogram(d), 2)
    res = dc.huff_encode(d, table)
    assert torch.equal(dc.huff_decompress(res.payload, res.bits(), table, data.size), d)


def test_decode_matches_oracle_on_oracle_stream(dc, oracle):
    """'reference-produced bitstream' (config 5): the oracle packs, the GPU decodes."""
    data = _zipf(dc, 300007, seed=5).cpu().numpy()
    hist = oracle.histogram_u8(data)
    for n_ary in PACKABLE:
        ln, el, ev, st = oracle.build_tables(hist, n_ary)
        for phase in (0, 6):
            payload, bits = oracle.pack(data, el, ev, oracle.bits_per_digit(n_ary), phase)
            buf = torch.zeros(payload.size + 64, dtype=torch.uint8, device="cuda")
            buf[: payload.size] = _dev(payload)
            table = dc.huff_table_from_lengths(_dev(ln), n_ary)
            out, status = dc.huff_decode(buf, bits, table, data.size, bit_start=phase)
            assert int(status.item()) == 0
            assert np.array_equal(out.cpu().numpy(), data), (n_ary, phase)
            assert np.array_equal(oracle.unpack(payload, phase, bits, ln, n_ary, data.size), data)


@pytest.mark.parametrize("mode", [1, 2])
def test_decode_robust_path(dc, oracle, mode):
    """mode 1: the iterative hand-off kernels alone; mode 2: fast path first, then forced fallback."""
    old = dc.lib().dc_debug_decode_mode(mode)
    try:
        for n_ary in PACKABLE:
            data = _zipf(dc, 200003, seed=8 + n_ary)
            table = dc.huff_build(dc.histogram(data), n_ary)
            for phase in (0, 5):
                res = dc.huff_encode(data, table, bit_phase=phase)
                out, status = dc.huff_decode(res.payload, res.bits(), table, data.numel(), bit_start=phase)
                assert int(status.item()) == 0 and torch.equal(out, data), (mode, n_ary, phase)
    finally:
        dc.lib().dc_debug_decode_mode(old)


def test_decode_fixed_length_codes_with_phase(dc, oracle):
    """256 equiprobable bytes -> 8/9-bit codes that barely self-synchronise; a shard phase shifts every code
    off the subsequence grid.  Whatever path is taken, the result must be exact."""
    rng = np.random.default_rng(12)
    data = rng.integers(0, 256, size=300000, dtype=np.uint8)
    d = _dev(data)
    for n_ary in PACKABLE:
        table = dc.huff_build(dc.histogram(d), n_ary)
        for phase in (0, 3, 4):
            res = dc.huff_encode(d, table, bit_phase=phase)
            out, status = dc.huff_decode(res.payload, res.bits(), table, data.size, bit_start=phase)
            assert int(status.item()) == 0 and torch.equal(out, d), (n_ary, phase)


def test_decode_reports_corruption(dc, oracle):
    data = _zipf(dc, 100000, seed=6)
    table = dc.huff_build(dc.histogram(data), 2)
    res = dc.huff_encode(data, table)
    nbits = res.bits()
    # wrong symbol count
    out, status = dc.huff_decode(res.payload, nbits, table, data.numel() - 1)
    assert int(status.item()) in (dc.DC_ERR_CAPACITY, dc.DC_ERR_CORRUPT)
    # truncated stream: ends inside a code or yields fewer symbols
    out, status = dc.huff_decode(res.payload, nbits - 3, table, data.numel())
    assert int(status.item()) == dc.DC_ERR_CORRUPT
    # a stream made of the unused (dummy) slot: binary always has one
    h = np.zeros(259, dtype=np.int64); h[[97, 98, 99, 100]] = [1, 1, 2, 2]
    t2 = dc.huff_build(_dev(h), 2)
    bad = torch.full((64,), 0xFF, dtype=torch.uint8, device="cuda")
    out, status = dc.huff_decode(bad, 64, t2, 32)
    assert int(status.item()) == dc.DC_ERR_CORRUPT


def test_radix_without_packing_is_table_only(dc):
    data = _zipf(dc, 4096)
    table = dc.huff_build(dc.histogram(data), 17)     # (radices up to 16 have a payload: a nibble per digit)
    assert table.download().bits_per_digit == 0
    res = dc.huff_encode(data, table)
    assert int(res.status.item()) == dc.DC_ERR_RADIX


def test_nybble_pack_unpack(dc, oracle):
    from data_compression_b200 import synth
    thr, base = synth.zipf_nybble_spec()
    for n in (0, 1, 2, 31, 32, 33, 63, 64, 65, 4097, (1 << 20) + 1, (1 << 22)):
        sym = torch.empty(n, dtype=torch.uint8, device="cuda")
        dc.synth_fill(sym, synth.SEED_BASE + 2, synth.device_thresholds(thr, "cuda"), base)
        packed, status = dc.nybble_pack(sym)
        assert int(status.item()) == 0
        assert np.array_equal(packed.cpu().numpy(), oracle.nybble_pack(sym.cpu().numpy())), n
        assert torch.equal(dc.nybble_unpack(packed, n), sym), n
    # unaligned views take the byte-granular kernels
    sym = torch.empty(10001, dtype=torch.uint8, device="cuda")
    dc.synth_fill(sym, 9, synth.device_thresholds(thr, "cuda"), base)
    v = sym[1:]
    packed, status = dc.nybble_pack(v)
    assert np.array_equal(packed.cpu().numpy(), oracle.nybble_pack(v.cpu().numpy()))
    assert torch.equal(dc.nybble_unpack(packed, v.numel()), v)
    # a shard of an aligned stream (symbols from an even offset, bytes from half of it): the head is peeled, the bulk stays vector
    big = torch.empty(200000, dtype=torch.uint8, device="cuda")
    dc.synth_fill(big, 11, synth.device_thresholds(thr, "cuda"), base)
    whole = oracle.nybble_pack(big.cpu().numpy())
    for lo in (2, 6, 18, 30, 32, 34, 4094):
        for n in (1, 5, 29, 31, 32, 33, 100001):
            outbuf = torch.full((100100,), 0xEE, dtype=torch.uint8, device="cuda")
            dst = outbuf[lo // 2: lo // 2 + (n + 1) // 2]
            packed, status = dc.nybble_pack(big[lo: lo + n], out=dst)
            assert int(status.item()) == 0
            want = oracle.nybble_pack(big[lo: lo + n].cpu().numpy())
            assert np.array_equal(dst.cpu().numpy(), want), (lo, n)
            assert (outbuf[: lo // 2] == 0xEE).all() and (outbuf[lo // 2 + (n + 1) // 2:] == 0xEE).all(), (lo, n)
            back = torch.full((200100,), 0xEE, dtype=torch.uint8, device="cuda")
            dc.nybble_unpack(dst, n, out=back[lo: lo + n])
            assert torch.equal(back[lo: lo + n], big[lo: lo + n]) and (back[:lo] == 0xEE).all() and (back[lo + n:] == 0xEE).all(), (lo, n)
    # a symbol >= 16 is reported (assert at nybble_compression.c:1093), its low nibble is packed
    bad = _dev(np.array([1, 2, 0x13, 4] * 16, dtype=np.uint8))
    packed, status = dc.nybble_pack(bad)
    assert int(status.item()) == dc.DC_ERR_SYMBOL
    assert packed.cpu().numpy()[1] == 0x34


def test_shard_bit_totals(dc, oracle):
    """dot(local histogram, global lengths) == bits the shard really emits."""
    data = _zipf(dc, 200000)
    table = dc.huff_build(dc.histogram(data), 16)
    for lo, hi in ((0, 70000), (70000, 200000)):
        want = dc.huff_encode(data[lo:hi].clone(), table).bits()
        got = int(dc.huff_bits_for_hist(dc.histogram(data[lo:hi].clone()), table).item())
        assert got == want


@pytest.mark.parametrize("n_ary", PACKABLE)
def test_shard_decoder_halo_and_exact_start_agree(dc, oracle, n_ary):
    """A shard of a longer stream decoded (a) by synchronising over its neighbour's last 1024 bytes and (b) from the
    exact first-code offset the neighbour reports: same assumption, same symbols -- and together with shard 0 the input."""
    data = _zipf(dc, 900001, seed=n_ary)
    host = data.cpu().numpy()
    ln, el, ev, st = oracle.build_tables(oracle.histogram_u8(host), n_ary)
    payload, total_bits = oracle.pack(host, el, ev, oracle.bits_per_digit(n_ary))
    table = dc.huff_table_from_lengths(_dev(ln.astype(np.int32)), n_ary)
    cut = (payload.size // 2) // 1024 * 1024
    pad = np.zeros(1024, dtype=np.uint8)
    buf0 = _dev(np.concatenate([pad, payload[:cut], payload[cut: cut + 1024]]))
    buf1 = _dev(np.concatenate([payload[cut - 1024: cut], payload[cut:], pad]))
    d0 = dc.ShardDecoder(buf0, cut, 8 * cut, total_bits, table)
    s0 = dc.ShardDecoder.unpack(d0.sync(has_halo=False, first_code_bit=0).cpu())
    d1 = dc.ShardDecoder(buf1, payload.size - cut, total_bits - 8 * cut, total_bits - 8 * cut, table)
    s1 = dc.ShardDecoder.unpack(d1.sync(has_halo=True).cpu())
    assert s0["resync"] == 0 and s1["resync"] == 0
    assert s1["assumed_start"] == s0["exit"]
    assert s0["symbols"] + s1["symbols"] == host.size
    out0, st0 = d0.write(s0["symbols"])
    out1, st1 = d1.write(s1["symbols"])
    assert int(st0.item()) == 0 and int(st1.item()) == 0
    assert np.array_equal(np.concatenate([out0.cpu().numpy(), out1.cpu().numpy()]), host)
    s1x = dc.ShardDecoder.unpack(d1.sync(has_halo=False, first_code_bit=s0["exit"]).cpu())
    assert s1x["symbols"] == s1["symbols"] and s1x["exit"] == s1["exit"] and s1x["assumed_start"] == s0["exit"]
    out1x, st1x = d1.write(s1x["symbols"])
    assert int(st1x.item()) == 0 and torch.equal(out1x, out1)


def test_radix3_trit_payload(dc, oracle):
    """Row N4: radix 3 (the reference's default).  The kernels run on one 2-bit field per trit, K7 turns that into the
    5-trits-per-byte payload the reference's author sketches (n_ary_huffman.c:745-748); everything against the oracle."""
    for size, seed in ((1, 1), (4, 2), (5, 3), (79, 4), (80, 5), (81, 6), (70001, 7), ((1 << 20) + 3, 8)):
        data = _zipf(dc, size, seed=seed)
        host = data.cpu().numpy()
        hist = dc.histogram(data)
        table = dc.huff_build(hist, 3)
        t = table.download()
        ln, el, ev, st = oracle.build_tables(hist.cpu().numpy().astype(np.uint64), 3)
        assert t.status == 0 and st == 0 and t.packed_radix == 3 and t.bits_per_digit == 2
        assert np.array_equal(np.array(t.lengths[:259]), ln) and np.array_equal(np.array(t.values[:259], dtype=np.uint32), ev)
        res = dc.huff_encode(data, table, out=torch.empty(size * 4 + 64, dtype=torch.uint8, device="cuda"))
        nbits = res.bits()
        want, wtrits = oracle.pack_trits(host, el, ev)
        assert nbits == 2 * wtrits == t.total_bits
        payload, pst = dc.trit_pack(res.payload, wtrits)
        assert int(pst.item()) == 0 and np.array_equal(payload.cpu().numpy(), want), size
        assert np.array_equal(oracle.unpack_trits(want, wtrits, ln, size), host) if size <= 70001 else True
        t2, ust = dc.trit_unpack(torch.from_numpy(want).cuda(), wtrits)
        assert int(ust.item()) == 0
        nb = (nbits + 7) // 8
        assert torch.equal(t2[: nb - 1], res.payload[: nb - 1])      # the same 2-bit stream (the last byte may differ in padding)
        out, status = dc.huff_decode(t2, nbits, table, size)
        assert int(status.item()) == 0 and torch.equal(out, data), size
        if size == 70001:   # the robust decode path (tile hand-off) with the trit-aware canonical search
            old = dc.lib().dc_debug_decode_mode(1)
            try:
                out, status = dc.huff_decode(t2, nbits, table, size)
            finally:
                dc.lib().dc_debug_decode_mode(old)
            assert int(status.item()) == 0 and torch.equal(out, data)
    # a 2-bit field of 3 in the kernels' stream is not a trit: reported, never decoded as something else
    broken = t2.clone()
    broken[nb // 2] = 0xFF
    out, status = dc.huff_decode(broken, nbits, table, size)
    assert int(status.item()) == dc.DC_ERR_CORRUPT
    # a payload byte outside 1..243 is reported
    bad = torch.tensor([0, 5, 244], dtype=torch.uint8, device="cuda")
    _, ust = dc.trit_unpack(bad, 15)
    assert int(ust.item()) == dc.DC_ERR_CORRUPT
