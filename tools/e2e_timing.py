#!/usr/bin/env python
"""Development tool: wall-clock split of the host-buffer round trip (pinned buffers) into compress and decompress."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from data_compression_b200 import synth, hostapi

n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
n_ary = int(os.environ.get("N_ARY", 4))
dev = torch.device("cuda:0")
thr, base = synth.zipf_bytes_spec()
d = torch.empty(n, dtype=torch.uint8, device=dev)
dc.synth_fill(d, synth.SEED_BASE + 2, synth.device_thresholds(thr, dev), base)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.copy_(d); del d
h_pay = torch.empty(n + n // 4 + 64, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
a, b, c = h_in.numpy(), h_pay.numpy(), h_out.numpy()
for it in range(4):
    t0 = time.perf_counter()
    p, bits, lens = hostapi.huff_compress(a, n_ary, out=b)
    t1 = time.perf_counter()
    hostapi.huff_decompress(p, bits, lens, n_ary, n, out=c)
    t2 = time.perf_counter()
    print(f"compress {1e3*(t1-t0):.2f} ms  decompress {1e3*(t2-t1):.2f} ms  total {n/(t2-t0)/1e9:.2f} GB/s  payload {p.size/1e6:.1f} MB")
assert np.array_equal(a, c)
# raw copy rates for reference
g = torch.empty(n, dtype=torch.uint8, device=dev)
for name, fn in (("H2D", lambda: g.copy_(h_in, non_blocking=True)), ("D2H", lambda: h_out.copy_(g, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name} {n/dt/1e9:.1f} GB/s")
# both directions at once (the decompress pipeline's floor: payload up while decoded bytes go down)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gp = torch.empty(p.size, dtype=torch.uint8, device=dev)
hp = h_pay[: p.size]
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        gp.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(g, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D {p.size/1e6:.0f} MB + D2H {n/1e6:.0f} MB concurrently: {1e3*dt:.2f} ms")
