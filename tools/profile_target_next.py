#!/usr/bin/env python
"""Profiling target for the "next" rows (K6 nybble text compressor, K8 move-to-front contexts, K7 trits): two rounds, the first
warms up.  ncu --set full -k regex:'tx_|mtf_|trit_' -s <launches of round 1> ... python tools/profile_target_next.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from data_compression_b200 import synth

n = int(os.environ.get("N", 256 << 20))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(7)
letters = torch.tensor(list(b" etaoins"), dtype=torch.uint8, device=dev)
others = torch.tensor([c for c in range(33, 127) if c not in b" etaoins"], dtype=torch.uint8, device=dev)
pick = torch.rand(n, device=dev, generator=g) < 0.62
text = torch.where(pick, letters[torch.randint(0, 8, (n,), device=dev, generator=g)], others[torch.randint(0, others.numel(), (n,), device=dev, generator=g)])
del pick
thr, base = synth.zipf_bytes_spec()
data = torch.empty(n, dtype=torch.uint8, device=dev)
dc.synth_fill(data, synth.SEED_BASE + 2, synth.device_thresholds(thr, dev), base)
table = dc.huff_build(dc.histogram(data), 3)
out = torch.empty(n + n // 2, dtype=torch.uint8, device=dev)
for _ in range(2):
    buf, ln, st = dc.nybble_text_compress(text)
    comp = buf[: int(ln.item())].clone()
    dc.nybble_text_decompress(comp)
    dc.nybble_adaptive_compress(text)
    res = dc.huff_encode(data, table, out=out)
    ntr = res.bits() // 2
    pay, _ = dc.trit_pack(res.payload, ntr)
    t2, _ = dc.trit_unpack(pay, ntr)
    dc.huff_decode(t2, 2 * ntr, table, n)
    torch.cuda.synchronize()
print("profile target (next rows) done")
