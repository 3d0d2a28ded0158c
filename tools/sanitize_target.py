#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from data_compression_b200 import synth

dev = torch.device("cuda:0")
thr, base = synth.zipf_bytes_spec()
for n in (1, 513, 40000, 300007):
    data = torch.empty(n, dtype=torch.uint8, device=dev)
    dc.synth_fill(data, 99 + n, synth.device_thresholds(thr, dev), base)
    hist = dc.histogram(data)
    for n_ary in (2, 4, 16):
        table = dc.huff_build(hist, n_ary)
        for phase in (0, 5):
            res = dc.huff_encode(data, table, bit_phase=phase)
            nbits = res.bits()
            out, st = dc.huff_decode(res.payload, nbits, table, n, bit_start=phase)
            assert int(st.item()) == 0 and torch.equal(out, data), (n, n_ary, phase)
    # robust decode path
    table = dc.huff_build(hist, 2)
    res = dc.huff_encode(data, table)
    nbits = res.bits()
    old = dc.lib().dc_debug_decode_mode(1)
    out, st = dc.huff_decode(res.payload, nbits, table, n)
    dc.lib().dc_debug_decode_mode(old)
    assert int(st.item()) == 0 and torch.equal(out, data)
    sym = data & 15
    packed, st = dc.nybble_pack(sym)
    assert torch.equal(dc.nybble_unpack(packed, n), sym)
    text = (data & 63) + 32
    buf, ln, st = dc.nybble_text_compress(text)
    c = buf[: int(ln.item())].clone()
    back, bl, st2 = dc.nybble_text_decompress(c)
    assert int(st.item()) == 0 and int(bl.item()) == n and torch.equal(back[:n], text)
# wide and mid tables
ln = np.zeros(259, dtype=np.int32)
s = 1
for depth in range(1, 15):
    for _ in range(3):
        ln[s] = depth; s += 1
for _ in range(4):
    ln[s] = 15; s += 1
table = dc.huff_table_from_lengths(torch.from_numpy(ln).to(dev), 4)
data = torch.randint(1, s, (50001,), dtype=torch.uint8, device=dev)
res = dc.huff_encode(data, table, out=torch.empty(data.numel() * 4 + 64, dtype=torch.uint8, device=dev))
out, st = dc.huff_decode(res.payload, res.bits(), table, data.numel())
assert int(st.item()) == 0 and torch.equal(out, data)
torch.cuda.synchronize()
print("sanitize target ok")
