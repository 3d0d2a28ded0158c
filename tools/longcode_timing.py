"""Decode timings on tables whose longest code exceeds the 12-bit index (Zipf exponents 1.5 .. 2.0); run on the GPU box."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from data_compression_b200 import synth
dev = torch.device("cuda:0"); n = 1 << 30; L = dc.lib()
for s_exp in (1.1, 1.5, 2.0):
    thr = synth.zipf_thresholds(255, s_exp)
    data = torch.empty(n, dtype=torch.uint8, device=dev); dc.synth_fill(data, 5, synth.device_thresholds(thr, dev), 1)
    for n_ary in (2, 4):
        t = dc.huff_build(dc.histogram(data), n_ary)
        if t.download().status != 0:
            print("zipf", s_exp, "n", n_ary, "table status", t.download().status, "(reference limit: codes of < 16 digits)", flush=True); continue
        out = torch.empty(n + n // 2, dtype=torch.uint8, device=dev)
        res = dc.huff_encode(data, t, out=out); nb = res.bits(); dec = torch.empty(n, dtype=torch.uint8, device=dev)
        o, st = dc.huff_decode(res.payload, nb, t, n, out=dec); assert int(st.item()) == 0 and torch.equal(o, data)
        L.dc_profile_reset(); L.dc_profile_enable(1)
        for _ in range(3):
            dc.huff_encode(data, t, out=out); dc.huff_decode(res.payload, nb, t, n, out=dec)
        torch.cuda.synchronize(); L.dc_profile_enable(0)
        row = []
        for kid in range(40):
            ms, cnt = C.c_double(0), C.c_uint64(0)
            L.dc_profile_kernel(kid, C.byref(ms), C.byref(cnt))
            if cnt.value and ms.value / cnt.value > 0.02: row.append(f"{L.dc_profile_kernel_name(kid).decode()}={ms.value/cnt.value:.3f}")
        print("zipf", s_exp, "n", n_ary, "max_bits", t.download().max_bits, "bits/sym", round(nb / n, 3), " ".join(row), flush=True)
        del out, dec, res
