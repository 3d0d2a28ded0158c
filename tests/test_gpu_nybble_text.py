"""GPU: the nybble compressor / decompressor -- static table and adaptive move-to-front contexts
(row N3) -- against the golden vectors generated from the unmodified compress_bytestring()/decompress_bytestring()
and against the oracle on text-like data."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _dev(b: bytes) -> torch.Tensor:
    return torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()).cuda()


def _run(fn, src: bytes):
    buf, n, st = fn(_dev(src) if src else torch.empty(0, dtype=torch.uint8, device="cuda"))
    return bytes(buf[: int(n.item())].cpu().numpy()), int(st.item()), buf, int(n.item())


def _textlike(rng, n, p_letter):
    letters = np.frombuffer(b" etaoins", dtype=np.uint8)
    others = np.array([c for c in range(1, 128) if c not in letters], dtype=np.uint8)
    pick = rng.random(n) < p_letter
    return np.where(pick, rng.choice(letters, n), rng.choice(others, n)).astype(np.uint8).tobytes()


def test_reference_golden_vectors(dc):
    for c in load_golden("nybble.json")["static"]:
        text, comp = bytes.fromhex(c["text"]), bytes.fromhex(c["compressed"])
        if not text:
            continue
        got, st, buf, n = _run(dc.nybble_text_compress, text)
        assert st == 0 and got == comp
        assert int(buf[n].item()) == 0                      # NUL-terminated like the reference's strings
        back, st, _, _ = _run(dc.nybble_text_decompress, comp)
        assert st == 0 and back == text
        assert dc.hostapi.compress_bytestring(text) == comp
        assert dc.hostapi.decompress_bytestring(comp) == text
    main = bytes.fromhex(load_golden("nybble.json")["static"][0]["text"])
    assert len(dc.hostapi.compress_bytestring(main)) <= 70  # nybble_compression.c:1162


@pytest.mark.parametrize("p_letter", [0.0, 0.3, 0.62, 0.95, 1.0])
@pytest.mark.parametrize("n", [1, 2, 3, 15, 16, 17, 4095, 4096, 4097, 70001, 1 << 20])
def test_matches_oracle(dc, oracle, n, p_letter):
    rng = np.random.default_rng(n * 7 + int(p_letter * 100))
    text = _textlike(rng, n, p_letter)
    want = oracle.nybble_static_compress(text)
    got, st, _, _ = _run(dc.nybble_text_compress, text)
    assert st == 0 and got == want, (n, p_letter)
    back, st, _, _ = _run(dc.nybble_text_decompress, want)
    assert st == 0 and back == oracle.nybble_static_decompress(want) == text


def test_long_runs_cross_tiles(dc, oracle):
    # hit runs longer than a 4 KB tile, odd and even, so the parity state crosses tile and warp boundaries
    for run in (4097, 8192, 12289):
        text = b"X" + b"e" * run + b"Q" + b"t" * (run + 1) + b"Z" * 3 + b" " * 7
        want = oracle.nybble_static_compress(text)
        got, st, _, _ = _run(dc.nybble_text_compress, text)
        assert st == 0 and got == want
        back, st, _, _ = _run(dc.nybble_text_decompress, want)
        assert st == 0 and back == text


def test_decoder_accepts_nibble_granular_literals(dc, oracle):
    # streams the compressor never writes but the reference decoder parses: literals that start on a low nibble,
    # a dangling half literal at the end, the ' ' (raw) type and an unknown type byte
    rng = np.random.default_rng(5)
    for n in (2, 3, 9, 100, 5000, 70000):
        body = rng.integers(1, 256, size=n, dtype=np.uint8).tobytes()
        for head in (b"\xafA", b" ", b"Q"):
            comp = head + body
            want = oracle.nybble_static_decompress(comp)
            got, st, _, _ = _run(dc.nybble_text_decompress, comp)
            assert st == 0 and got == want, (n, head)


def test_errors(dc):
    got, st, _, _ = _run(dc.nybble_text_compress, b"abc\x80def")
    assert st == dc.DC_ERR_SYMBOL                              # assert( source[i] < 0x80 ) :910
    got, st, _, _ = _run(dc.nybble_adaptive_compress, b"abc\x80def")
    assert st == dc.DC_ERR_SYMBOL
    assert dc.hostapi.compress_bytestring(b"") == b""
    assert dc.hostapi.decompress_bytestring(b"") == b""


# ------------------------------------------------------------------------------------------ adaptive contexts (row N3)

def _english(rng, n):
    words = [b"the", b"and", b"this", b"is", b"a", b"test", b"banana", b"Hello,", b"world.", b"only", b"of", b"to", b"in", b"it",
             b"compression", b"nybble", b"context", b"Q", b"42", b"\n"]
    out = bytearray()
    while len(out) < n:
        out += words[int(rng.integers(len(words)))] + b" "
    return bytes(out[:n])


def test_adaptive_reference_golden_vectors(dc):
    g = load_golden("nybble.json")["adaptive"]
    for c in g:
        text, comp = bytes.fromhex(c["text"]), bytes.fromhex(c["compressed"])
        got, st, buf, n = _run(dc.nybble_adaptive_compress, text)
        assert st == 0 and got == comp, len(text)
        assert int(buf[n].item()) == 0
        back, st, _, _ = _run(dc.nybble_adaptive_decompress, comp)
        assert st == 0 and back == text
        assert dc.hostapi.compress_bytestring(text, modify=True) == comp          # nybble_compress() :1134
        assert dc.hostapi.decompress_bytestring(comp, modify=True) == text        # nybble_decompress() :1117
    assert len(dc.hostapi.compress_bytestring(bytes.fromhex(g[0]["text"]), modify=True)) <= 70   # :1178


@pytest.mark.parametrize("kind", ["english", "letters", "uniform", "one_context"])
@pytest.mark.parametrize("n", [1, 2, 3, 16, 17, 511, 512, 513, 1025, 32768, 32769, 70001, (1 << 20) + 5])
def test_adaptive_matches_oracle(dc, oracle, n, kind):
    rng = np.random.default_rng(n * 13 + len(kind))
    if kind == "english":
        text = _english(rng, n)
    elif kind == "letters":
        text = _textlike(rng, n, 0.7)
    elif kind == "uniform":
        text = rng.integers(1, 128, size=n, dtype=np.uint8).tobytes()   # every context, mostly misses
    else:
        text = rng.choice(np.frombuffer(b"abcdefg`", dtype=np.uint8), size=n).tobytes()  # one context, 8 letters: all hits soon
    want = oracle.nybble_adaptive_compress(text)
    got, st, _, _ = _run(dc.nybble_adaptive_compress, text)
    assert st == 0 and got == want, (n, kind)
    if n <= 70001:   # the decoder's resolve step is one serial walk
        back, st, _, _ = _run(dc.nybble_adaptive_decompress, want)
        assert st == 0 and back == oracle.nybble_adaptive_decompress(want) == text


def test_adaptive_lists_cross_blocks_and_chunks(dc, oracle):
    # a letter seen once, then 40 000 bytes of other contexts, then used again: its list position has to survive
    # 78 blocks of 512 bytes and a chunk boundary (64 blocks); and a context that never appears in between
    rng = np.random.default_rng(8)
    filler = rng.choice(np.frombuffer(b"hijklmno", dtype=np.uint8), size=40000).tobytes()   # context 13 only
    text = b"A~A~A~" + filler + b"A~A~" + filler[:700] + b"Az~"
    want = oracle.nybble_adaptive_compress(text)
    got, st, _, _ = _run(dc.nybble_adaptive_compress, text)
    assert st == 0 and got == want
    back, st, _, _ = _run(dc.nybble_adaptive_decompress, want)
    assert st == 0 and back == text


def test_adaptive_decoder_on_arbitrary_streams(dc, oracle):
    rng = np.random.default_rng(6)
    for n in (2, 3, 9, 100, 5000):
        body = rng.integers(1, 256, size=n, dtype=np.uint8).tobytes()
        for head in (b"\xafA", b" ", b"Q"):
            comp = head + body
            want = oracle.nybble_adaptive_decompress(comp)
            got, st, _, _ = _run(dc.nybble_adaptive_decompress, comp)
            assert st == 0 and got == want, (n, head)


# ------------------------------------------------------------------------------------------ many strings per call (replicas)

@pytest.mark.parametrize("modify", [False, True])
def test_batch_matches_reference_and_oracle(dc, oracle, modify):
    g = load_golden("nybble.json")["adaptive" if modify else "static"]
    texts = [bytes.fromhex(c["text"]) for c in g]
    comps = [bytes.fromhex(c["compressed"]) for c in g]
    assert dc.nybble_text_compress_batch(texts, modify) == comps            # the unmodified reference's own outputs
    assert dc.nybble_text_decompress_batch(comps, modify) == texts
    rng = np.random.default_rng(17 + int(modify))
    many = [b"", b"a", b"e", b"ab", b" e"]
    for k in range(3000):
        n = int(rng.integers(1, 200))
        many.append(_english(rng, n) if k % 3 else rng.integers(1, 128, size=n, dtype=np.uint8).tobytes())
    comp_fn = oracle.nybble_adaptive_compress if modify else oracle.nybble_static_compress
    dec_fn = oracle.nybble_adaptive_decompress if modify else oracle.nybble_static_decompress
    want = [comp_fn(t) for t in many]
    got = dc.nybble_text_compress_batch(many, modify)
    assert got == want
    assert dc.nybble_text_decompress_batch(want, modify) == many
    # arbitrary streams through the decoder (nibble-granular literals, unknown type bytes)
    junk = [bytes([0xAF, 0x41]) + rng.integers(1, 256, size=int(rng.integers(0, 60)), dtype=np.uint8).tobytes() for _ in range(500)]
    assert dc.nybble_text_decompress_batch(junk, modify) == [dec_fn(j) for j in junk]
    with pytest.raises(dc.DcError) as e:
        dc.nybble_text_compress_batch([b"ok", b"bad\x80byte"], modify)
    assert e.value.status == dc.DC_ERR_SYMBOL
