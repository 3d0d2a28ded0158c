"""Regenerates tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile with -DNDEBUG).  Run in the authoring container only:

    python tests/golden/make_golden.py

The vectors travel with the repo; /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RADICES = (2, 3, 4, 5, 10, 16)


def histograms():
    rng = np.random.default_rng(0x5EED0001)
    cases = []

    def add(name, h):
        cases.append((name, np.asarray(h, dtype=np.int64)))

    for t in range(16):
        h = np.zeros(259, dtype=np.int64)
        k = int(rng.integers(1, 256))
        idx = rng.choice(np.arange(1, 256), size=k, replace=False)
        mode = t % 4
        if mode == 0:
            h[idx] = rng.integers(1, 5, size=k)            # heavy ties
        elif mode == 1:
            h[idx] = rng.integers(1, 100000, size=k)
        elif mode == 2:
            h[idx] = (1e6 / (np.arange(1, k + 1) ** 1.1)).astype(np.int64) + 1
        else:
            h[idx] = 7                                     # all equal
        add(f"random{t}", h)
    h = np.zeros(259, dtype=np.int64); h[65] = 5; add("one_symbol", h)
    h = np.zeros(259, dtype=np.int64); h[65] = 5; h[66] = 7; add("two_symbols", h)
    h = np.zeros(259, dtype=np.int64); h[[97, 98, 99, 100]] = [1, 1, 2, 2]; add("abcd_1122", h)
    h = np.zeros(259, dtype=np.int64); h[[1, 2, 3, 4]] = [5, 7, 9, 11]; add("four_symbols", h)
    h = np.zeros(259, dtype=np.int64); h[1:17] = np.arange(16) + 3; add("sixteen_symbols", h)
    h = np.zeros(259, dtype=np.int64); h[1:256] = 4; add("equal_255", h)
    h = np.zeros(259, dtype=np.int64); h[0:256] = 3; add("equal_256_with_nul", h)
    h = np.zeros(259, dtype=np.int64); h[0] = 50; h[1] = 30; h[2] = 10; h[255] = 10; add("nul_used", h)
    # expected Zipf(1.1) counts at N = 2^30 over bytes 1..255
    w = np.arange(1, 256, dtype=np.float64) ** -1.1
    h = np.zeros(259, dtype=np.int64); h[1:256] = np.floor(w / w.sum() * (1 << 30)).astype(np.int64); add("zipf_2p30", h)
    # Fibonacci-like counts: deepest possible tree for the symbol count (binary max length 15 is the ref limit)
    h = np.zeros(259, dtype=np.int64); fib = [1, 1]
    while len(fib) < 15: fib.append(fib[-1] + fib[-2])
    h[1:16] = fib; add("fibonacci_15", h)
    return cases


def main():
    assert O.have_ref(), "oracle/_ref is not built (make -C oracle needs /root/reference)"
    tables = []
    for name, h in histograms():
        entry = {"name": name, "hist": {str(i): int(v) for i, v in enumerate(h) if v}, "radix": {}}
        for n in RADICES:
            lengths = O.ref_huffman(h.astype(np.int32), n)
            rec = {"lengths": {str(i): int(v) for i, v in enumerate(lengths) if v}}
            if lengths.max() < 16 and n ** int(lengths.max()) < 2 ** 31 and lengths.max() > 0:
                el, ev = O.ref_convert_lengths_to_encode_table(lengths, n)
                assert np.array_equal(el, lengths)
                rec["values"] = {str(i): int(ev[i]) for i in range(259) if lengths[i]}
            entry["radix"][str(n)] = rec
        tables.append(entry)
    with open(os.path.join(HERE, "huffman_tables.json"), "w") as f:
        json.dump({"source": "unmodified n_ary_huffman.c huffman() + convert_lengths_to_encode_table(), -DNDEBUG",
                   "cases": tables}, f, separators=(",", ":"))

    # the reference's own KATs for convert_lengths_to_encode_table (n_ary_huffman.c:2821-2891), re-run here
    kats = []
    for lens in ([0, 0, 1, 1, 1], [0, 0] + [2] * 8, [0, 0] + [2] * 9):
        l = np.zeros(80, dtype=np.int32); l[:len(lens)] = lens
        el, ev = O.ref_convert_lengths_to_encode_table(l, 3, max_symbol_value=20)
        kats.append({"max_symbol_value": 20, "n": 3, "lengths": l.tolist(), "elen": el.tolist(), "evalue": ev.tolist()})
    # last-slot quirk (:1336/:1360/:1421): a non-zero length in slot max_symbol_value
    l = np.zeros(8, dtype=np.int32); l[[1, 2, 7]] = [1, 2, 2]
    def prefill():
        return np.full(8, 77, dtype=np.int32), np.full(8, 88, dtype=np.uint32)
    el0, ev0 = prefill()
    el, ev = O.ref_convert_lengths_to_encode_table(l, 2, max_symbol_value=7, elen=el0, evalue=ev0)
    kats.append({"max_symbol_value": 7, "n": 2, "lengths": l.tolist(), "elen": el.tolist(), "evalue": ev.tolist(),
                 "prefill": [77, 88]})
    l = np.zeros(8, dtype=np.int32); l[[1, 2, 7]] = [1, 2, 3]   # slot 7 longer than max over i<7: never assigned
    el0, ev0 = prefill()
    el, ev = O.ref_convert_lengths_to_encode_table(l, 2, max_symbol_value=7, elen=el0, evalue=ev0)
    kats.append({"max_symbol_value": 7, "n": 2, "lengths": l.tolist(), "elen": el.tolist(), "evalue": ev.tolist(),
                 "prefill": [77, 88]})
    with open(os.path.join(HERE, "convert_kats.json"), "w") as f:
        json.dump({"source": "unmodified convert_lengths_to_encode_table()", "cases": kats}, f)

    # histogram() on 7-bit text (n_ary_huffman.c:461): the reference's embedded self-test text + canary slot 258
    text = (b"/* n_ary_huffman.c" b"2021-10-25: started by David Cary")
    h = O.ref_histogram(text)
    with open(os.path.join(HERE, "histogram.json"), "w") as f:
        json.dump({"source": "unmodified histogram()", "text": text.decode(), "hist": h.tolist()}, f)

    # nybble_compression.c: its main() text in static and adaptive mode, plus write_nybble streams
    text = b"Hello, world. This is a test. This is only a test. Banana banana banana banana. "
    rng = np.random.default_rng(7)
    texts = [text, b"a", b"ab", b"e e e e", b"xyz", b" etaoins" * 5, b"The quick brown fox jumps over the lazy dog."]
    for _ in range(6):
        k = int(rng.integers(2, 200))
        alphabet = np.frombuffer(b" etaoinsxyzQ.,", dtype=np.uint8)
        texts.append(bytes(rng.choice(alphabet, size=k).tolist()))
    ny = {"source": "unmodified compress_bytestring()/write_nybble()", "static": [], "adaptive": [], "write_nybble": []}
    for t in texts:
        c = O.ref_compress_bytestring(t, False)
        assert O.ref_decompress_bytestring(c, False) == t
        ny["static"].append({"text": t.hex(), "compressed": c.hex()})
    long_text = (text * 40)[:3001] + bytes(rng.choice(np.frombuffer(b" etaoinsrhldcu.,\nTHE", dtype=np.uint8), size=5000).tolist())
    for t in texts + [long_text]:
        c = O.ref_compress_bytestring(t, True)   # nybble_compress() :1134
        assert O.ref_decompress_bytestring(c, True) == t
        ny["adaptive"].append({"text": t.hex(), "compressed": c.hex()})
    for k in (0, 1, 2, 7, 32, 33, 95):
        s = rng.integers(0, 16, size=k).astype(np.uint8)
        ny["write_nybble"].append({"symbols": s.tolist(), "packed": O.ref_write_nybble_stream(s).tolist()})
    with open(os.path.join(HERE, "nybble.json"), "w") as f:
        json.dump(ny, f)
    print("golden vectors written to", HERE)


def container_golden():
    """The reference's own block writer, static compress() (n_ary_huffman.c:1688-1815): what it leaves in its output buffer.
    Under -DNDEBUG it always ends in the raw pass-through block "<len>:\\n\\n<text>," (no line feed behind the comma); with
    compressed_symbols > 2 the table block "265:\\nX258:<259 digits>," it formatted first is still in the buffer behind a
    short text (its first bytes are overwritten by the raw block)."""
    rng = np.random.default_rng(11)
    cases = []
    texts = [b"A", b"hello world, hello netstrings" * 3,
             bytes(rng.choice(np.frombuffer(b" etaoinsrhldcu.,\nTHE", dtype=np.uint8), size=3000).tolist())]
    for n in (2, 3, 4):
        for t in texts:
            h = O.ref_histogram(t)
            lengths = O.ref_huffman(h, n)
            buf = O.ref_compress_block(t, lengths, n)
            end = len(buf)
            while end > 0 and buf[end - 1] == 0xFF:
                end -= 1
            cases.append({"n": n, "text": t.hex(), "lengths": lengths.tolist(), "buffer": buf[:min(end, len(t) + 400)].hex()})
    with open(os.path.join(HERE, "container.json"), "w") as f:
        json.dump({"source": "unmodified static compress() via oracle/ref_harness_huff.c, buffer pre-filled with 0xFF", "cases": cases}, f)


if __name__ == "__main__":
    main()
    container_golden()
