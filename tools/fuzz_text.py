#!/usr/bin/env python
"""Randomised cross-check of the nybble compressors (static table, adaptive contexts, single strings and batches) against the
oracle; run on the GPU box.   python tools/fuzz_text.py [seconds]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from oracle import pyoracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(os.environ.get("SEED", 2)))
t_end = time.time() + budget
it = fails = 0
alphabets = [b" etaoins", b" etaoinsrhldcu.,\nTHE", bytes(range(1, 128)), b"abcdefg`", b"e", b" e", bytes(range(0x20, 0x30))]
def dev(b): return torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()).cuda()
def run(fn, b):
    buf, n, st = fn(dev(b))
    return bytes(buf[: int(n.item())].cpu().numpy()), int(st.item())
while time.time() < t_end:
    it += 1
    a = np.frombuffer(alphabets[int(rng.integers(len(alphabets)))], dtype=np.uint8)
    size = int(rng.choice([1, 2, 3, 17, 100, 511, 512, 513, 4096, 8193, 40000, 70001]))
    p = rng.random(a.size) ** float(rng.choice([1.0, 3.0])); p /= p.sum()
    text = rng.choice(a, size=size, p=p).astype(np.uint8).tobytes()
    for name, cfn, dfn, ocf, odf in (("static", dc.nybble_text_compress, dc.nybble_text_decompress, O.nybble_static_compress, O.nybble_static_decompress),
                                     ("adaptive", dc.nybble_adaptive_compress, dc.nybble_adaptive_decompress, O.nybble_adaptive_compress, O.nybble_adaptive_decompress)):
        want = ocf(text)
        got, st = run(cfn, text)
        ok = st == 0 and got == want
        if size <= 40000 or name == "static":
            back, st2 = run(dfn, want)
            ok = ok and st2 == 0 and back == text
        junk = bytes([0xAF, 0x41]) + rng.integers(1, 256, size=int(rng.integers(0, 3000)), dtype=np.uint8).tobytes()
        j, st3 = run(dfn, junk)
        ok = ok and st3 == 0 and j == odf(junk)
        if not ok:
            fails += 1
            print("MISMATCH", name, it, size, len(a), flush=True)
    if it % 20 == 0:   # batches
        many = [rng.choice(a, size=int(rng.integers(0, 300)), p=p).astype(np.uint8).tobytes() for _ in range(500)]
        for modify, ocf, odf in ((False, O.nybble_static_compress, O.nybble_static_decompress), (True, O.nybble_adaptive_compress, O.nybble_adaptive_decompress)):
            want = [ocf(t) if t else b"" for t in many]
            if dc.nybble_text_compress_batch(many, modify) != want or dc.nybble_text_decompress_batch(want, modify) != many:
                fails += 1
                print("BATCH MISMATCH", modify, it, flush=True)
print("iterations", it, "mismatches", fails)
sys.exit(1 if fails else 0)
