"""CPU: the oracle restatement against the committed golden vectors (generated from the UNMODIFIED
reference by tests/golden/make_golden.py) and against the reference's own known-answer tests."""
import numpy as np
import pytest

from conftest import dense, load_golden

RADICES = (2, 3, 4, 5, 10, 16)


def test_huffman_lengths_match_reference_golden(oracle, table_cases):
    for case in table_cases:
        h = dense(case["hist"])
        for n in RADICES:
            want = dense(case["radix"][str(n)]["lengths"], dtype=np.int32)
            got = oracle.huffman(h.astype(np.uint64), n)
            assert np.array_equal(got, want), (case["name"], n)


def test_canonical_values_match_reference_golden(oracle, table_cases):
    checked = 0
    for case in table_cases:
        for n in RADICES:
            rec = case["radix"][str(n)]
            if "values" not in rec:
                continue
            lengths = dense(rec["lengths"], dtype=np.int32)
            el, ev, st = oracle.convert_lengths_to_encode_table(lengths, n)
            assert st == 0
            assert np.array_equal(el, lengths)
            assert np.array_equal(ev, dense(rec["values"], dtype=np.uint32)), (case["name"], n)
            checked += 1
    assert checked > 100


def test_convert_kats_incl_last_slot_quirk(oracle):
    # n_ary_huffman.c:2821-2891 (3 KATs) + the i < max_symbol_value quirk (:1336, :1360, :1421)
    for k in load_golden("convert_kats.json")["cases"]:
        pre = k.get("prefill", [0, 0])
        n = len(k["lengths"])
        el, ev, st = oracle.convert_lengths_to_encode_table(
            np.array(k["lengths"], dtype=np.int32), k["n"], max_symbol_value=k["max_symbol_value"],
            elen=np.full(n, pre[0], dtype=np.int32), evalue=np.full(n, pre[1], dtype=np.uint32))
        msv = k["max_symbol_value"]
        assert el[: msv + 1].tolist() == k["elen"][: msv + 1]
        assert ev[: msv + 1].tolist() == k["evalue"][: msv + 1]


def test_summarize_tree_kats(oracle):
    # n_ary_huffman.c:1112-1154: {a:9,b:9} -> 1,1 ; {a:9,b:9,c:8} -> a=1,b=2,c=2 (plus the as-written dummy, F2)
    h = np.zeros(259, dtype=np.uint64)
    h[ord("a")] = 9; h[ord("b")] = 9; h[ord("c")] = 8
    ln = oracle.huffman(h, 3)  # trinary, 3 symbols -> d=2 as written: lengths still 1..2
    assert ln[ord("a")] >= 1 and ln.sum() > 0


def test_survey_golden_lengths(oracle):
    # SURVEY 4, run-derived: n=2 {a:1,b:1,c:2,d:2} -> 3,3,2,2 ; one symbol -> 1 ; {5,7} -> 2,1 ; n=4 {5,7,9,11} -> 2,1,1,1
    h = np.zeros(259, dtype=np.uint64); h[[97, 98, 99, 100]] = [1, 1, 2, 2]
    assert oracle.huffman(h, 2)[[97, 98, 99, 100]].tolist() == [3, 3, 2, 2]
    h = np.zeros(259, dtype=np.uint64); h[65] = 5
    assert oracle.huffman(h, 2)[65] == 1
    h[66] = 7
    assert oracle.huffman(h, 2)[[65, 66]].tolist() == [2, 1]
    h = np.zeros(259, dtype=np.uint64); h[[1, 2, 3, 4]] = [5, 7, 9, 11]
    assert oracle.huffman(h, 4)[[1, 2, 3, 4]].tolist() == [2, 1, 1, 1]
    h = np.zeros(259, dtype=np.uint64); h[[0, 1, 2, 255]] = [50, 30, 10, 10]
    ln = oracle.huffman(h, 2)
    assert ln[[0, 1, 2, 255]].tolist() == [1, 2, 4, 3]
    assert oracle.convert_lengths_to_encode_table(ln, 2)[1][[0, 1, 2, 255]].tolist() == [0, 2, 14, 6]


def test_histogram_golden(oracle):
    g = load_golden("histogram.json")
    text = g["text"].encode()
    assert oracle.histogram_cstr(text).tolist() == g["hist"]
    assert oracle.histogram_u8(text).tolist() == g["hist"]
    assert g["hist"][258] == 0  # the canary slot is zeroed (:2663-2666)
    assert oracle.histogram_cstr(b"ab\x00cd")[ord("c")] == 0  # NUL terminates (:482)
    assert oracle.histogram_u8(b"ab\x00cd")[ord("c")] == 1
    big = np.random.default_rng(3).integers(0, 256, size=1 << 20, dtype=np.uint8)
    assert np.array_equal(oracle.histogram_u8(big, threads=4)[:256], np.bincount(big, minlength=256))


def test_nybble_golden(oracle):
    g = load_golden("nybble.json")
    for c in g["write_nybble"]:
        s = np.array(c["symbols"], dtype=np.uint8)
        assert oracle.nybble_pack(s).tolist() == c["packed"]
        assert oracle.nybble_unpack(np.array(c["packed"], dtype=np.uint8), s.size).tolist() == c["symbols"]
    for c in g["static"]:
        text, comp = bytes.fromhex(c["text"]), bytes.fromhex(c["compressed"])
        assert oracle.nybble_static_compress(text) == comp
        assert oracle.nybble_static_decompress(comp) == text
    for c in g["adaptive"]:
        text, comp = bytes.fromhex(c["text"]), bytes.fromhex(c["compressed"])
        assert oracle.nybble_adaptive_compress(text) == comp
        assert oracle.nybble_adaptive_decompress(comp) == text
    assert len(bytes.fromhex(g["adaptive"][0]["compressed"])) <= 70  # nybble_compression.c:1187
    main_text = bytes.fromhex(g["static"][0]["text"])
    assert len(bytes.fromhex(g["static"][0]["compressed"])) <= 70  # nybble_compression.c:1162
    assert len(main_text) == 80


@pytest.mark.parametrize("n", [2, 4, 16])
def test_pack_unpack_roundtrip_and_layout(oracle, n):
    rng = np.random.default_rng(n)
    data = rng.choice(np.arange(1, 40, dtype=np.uint8), size=5000, p=np.arange(39, 0, -1) / 780.0)
    hist = oracle.histogram_u8(data)
    lengths, el, ev, st = oracle.build_tables(hist, n)
    assert st == 0
    bpd = oracle.bits_per_digit(n)
    for phase in (0, 3, 7):
        payload, bits = oracle.pack(data, el, ev, bpd, phase)
        assert bits == int((hist[:256].astype(np.int64) * lengths[:256] * bpd).sum())
        assert payload.size == (bits + phase + 7) // 8
        # MSB-first layout, checked against an independent bit-string construction
        s = "0" * phase + "".join(format(int(ev[b]), "b").zfill(int(el[b]) * bpd) for b in data)
        s += "0" * (-len(s) % 8)
        assert payload.tobytes() == int(s, 2).to_bytes(len(s) // 8, "big")
        back = oracle.unpack(payload, phase, bits, lengths, n, data.size)
        assert np.array_equal(back, data)
        p2, b2, offs = oracle.pack_mt(data, el, ev, bpd, phase, threads=3, block_symbols=777)
        assert b2 == bits and np.array_equal(p2, payload)
        assert np.array_equal(oracle.unpack_mt(payload, phase, lengths, n, data.size, offs, 777, 3), data)


def test_trit_payload_layout_and_roundtrip(oracle):
    """Radix 3 (the reference's default, n_ary_huffman.c:2529): 5 trits per octet, byte = 1 + base-3 value, the scheme the
    author sketches at :745-748 ("never uses byte 0 or 244..255")."""
    # hand vector: A x3, B x3, C once; the as-written dummy rule (:786) adds two dummies here, so C sits at depth 2
    data = np.frombuffer(b"ABBAAB", dtype=np.uint8)
    h = oracle.histogram_u8(np.frombuffer(b"ABBAABC", dtype=np.uint8))
    lengths, el, ev, st = oracle.build_tables(h, 3)
    assert st == 0 and [int(lengths[s]) for s in (65, 66, 67)] == [1, 1, 2]
    payload, trits = oracle.pack_trits(data, el, ev)
    assert trits == 6
    t = [int(ev[b]) for b in data]
    g0 = 1 + t[0] * 81 + t[1] * 27 + t[2] * 9 + t[3] * 3 + t[4]
    g1 = 1 + t[5] * 81
    assert payload.tolist() == [g0, g1]
    assert oracle.unpack_trits(payload, trits, lengths, data.size).tobytes() == b"ABBAAB"

    rng = np.random.default_rng(33)
    for size in (1, 4, 5, 6, 4999):
        data = rng.choice(np.arange(1, 60, dtype=np.uint8), size=size, p=np.arange(59, 0, -1) / 1770.0)
        hist = oracle.histogram_u8(data)
        hist[200] += 1  # keep at least two symbols
        lengths, el, ev, st = oracle.build_tables(hist, 3)
        assert st == 0
        payload, trits = oracle.pack_trits(data, el, ev)
        assert trits == int(sum(int(lengths[b]) for b in data))
        assert payload.size == (trits + 4) // 5
        assert payload.min() >= 1 and payload.max() <= 243
        # independent construction: codes as base-3 numerals, MSB trit first
        digits = []
        for b in data:
            v, L = int(ev[b]), int(lengths[b])
            digits += [(v // 3 ** (L - 1 - k)) % 3 for k in range(L)]
        digits += [0] * (-len(digits) % 5)
        want = [1 + sum(d * 3 ** (4 - k) for k, d in enumerate(digits[i:i + 5])) for i in range(0, len(digits), 5)]
        assert payload.tolist() == want
        assert np.array_equal(oracle.unpack_trits(payload, trits, lengths, data.size), data)


def test_unpack_flags_unused_slot(oracle):
    # binary always carries one dummy leaf (F2): its code slot must be reported, not decoded
    h = np.zeros(259, dtype=np.uint64); h[[97, 98, 99, 100]] = [1, 1, 2, 2]
    lengths, el, ev, st = oracle.build_tables(h, 2)
    # c=00 d=01 a=100 b=101 -> the dummy leaf owns the depth-2 slot '11' 
    assert [(int(ev[s]), int(lengths[s])) for s in (97, 98, 99, 100)] == [(4, 3), (5, 3), (0, 2), (1, 2)]
    bits = np.array([0b11000000], dtype=np.uint8)
    out, status = oracle.unpack(bits, 0, 2, lengths, 2, 1, return_status=True)
    assert status == oracle.ORC_ERR_CORRUPT


def test_nybble_compressors_round_trip_property(oracle):
    """Both modes of the oracle's nybble compressor invert on arbitrary 7-bit strings, never emit a NUL, and never grow the
    text by more than the type byte (the ' ' + raw fall-back, nybble_compression.c:1018-1037)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.binary(min_size=1, max_size=600).map(lambda b: bytes((x % 127) + 1 for x in b)))
    def check(text):
        for comp_fn, dec_fn in ((oracle.nybble_static_compress, oracle.nybble_static_decompress),
                                (oracle.nybble_adaptive_compress, oracle.nybble_adaptive_decompress)):
            comp = comp_fn(text)
            assert 0 not in comp and len(comp) <= len(text) + 1
            assert dec_fn(comp) == text

    check()


def test_base64url_text_form(oracle):
    """int2digit() n_ary_huffman.c:371-426 is the RFC 4648 base64url alphabet and the unfinished packer emits 6 bits per
    character (:1646-1671): on whole bytes the text form is Python's urlsafe base64 without padding."""
    import base64
    rng = np.random.default_rng(64)
    for nbytes in (0, 1, 2, 3, 4, 11, 12, 13, 1000, 4099):
        raw = rng.integers(0, 256, size=nbytes, dtype=np.uint8).tobytes()
        text = oracle.base64url_pack(raw, 8 * nbytes)
        assert text == base64.urlsafe_b64encode(raw).rstrip(b"=")
        assert oracle.base64url_unpack(text, 8 * nbytes).tobytes() == raw
        std = base64.b64encode(raw).rstrip(b"=")                       # digit2int() also takes '+' and '/' (:441-445)
        assert oracle.base64url_unpack(std, 8 * nbytes).tobytes() == raw
    # a bit count that is no multiple of 6 or 8: zero padding in the last character, bits behind the end ignored
    raw = bytes([0b10110111, 0b01011111])
    assert oracle.base64url_pack(raw, 11) == b"t0"                       # 101101 = 45 = 't' | 11010(0) = 52 = '0'
    assert oracle.base64url_unpack(b"t0", 11).tobytes() == bytes([0b10110111, 0b01000000])
    with pytest.raises(ValueError):
        oracle.base64url_unpack(b"t!", 11)


def test_nibble_per_digit_layout_statement_roundtrips_on_cpu():
    """Radices 5 .. 15 (9 and 10 by name): the oracle's statement of the payload -- canonical base-n values rewritten
    with one nibble per digit, packed MSB first -- read back by an independent pure-Python digit reader."""
    import numpy as np
    from oracle import pyoracle as O
    O.build()
    rng = np.random.default_rng(4)
    for n in (5, 9, 10, 15):
        w = 1.0 / np.arange(1, 201) ** 1.1
        data = rng.choice(np.arange(1, 201, dtype=np.uint8), size=5000, p=w / w.sum()).astype(np.uint8)
        ln, el, ev, st = O.build_tables(O.histogram_u8(data), n)
        assert st == 0
        nv = O.nibble_values(el, ev, n)
        for s in range(1, 201):   # every nibble is a digit of the radix, and the numeral is the canonical value
            digits = [(int(nv[s]) >> (4 * k)) & 15 for k in range(int(el[s]))][::-1]
            assert all(d < n for d in digits)
            v = 0
            for d in digits:
                v = v * n + d
            assert el[s] == 0 or v == int(ev[s])
        for phase in (0, 4):
            payload, bits = O.pack(data, el, nv, 4, phase)
            assert bits == int(sum(int(el[b]) for b in data)) * 4
            assert np.array_equal(O.unpack_nibble_digits(payload, phase, bits, el, ev, n), data)
