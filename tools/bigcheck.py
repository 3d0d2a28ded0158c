#!/usr/bin/env python
"""Large inputs on tables with long codes (every encoder instantiation and decoder mode at a scale where offsets leave 32 bits):
payload against the oracle (block-parallel on the host cores), then the round trip.  Run on the GPU box."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from data_compression_b200 import synth
from oracle import pyoracle as O

n = int(os.environ.get("N", 768 << 20))
threads = os.cpu_count() or 1
dev = torch.device("cuda:0")
fails = 0
for n_ary, s_exp in ((2, 1.5), (4, 1.5), (4, 2.0), (4, 3.0), (16, 2.0), (16, 3.0), (3, 1.5), (3, 2.5)):
    thr = synth.zipf_thresholds(255, s_exp)
    data = torch.empty(n, dtype=torch.uint8, device=dev)
    dc.synth_fill(data, 77 + n_ary, synth.device_thresholds(thr, dev), 1)
    host = data.cpu().numpy()
    hist = dc.histogram(data)
    o_hist = O.histogram_u8(host, threads=threads)
    table = dc.huff_build(hist, n_ary)
    t = table.download()
    ln, el, ev, st = O.build_tables(o_hist, n_ary)
    if t.status != 0 or st != 0:
        print("n", n_ary, "zipf", s_exp, "table status", t.status, st, "(skipped)", flush=True)
        continue
    res = dc.huff_encode(data, table, out=torch.empty(n + n // 2 + 64, dtype=torch.uint8, device=dev))
    nbits = res.bits()
    t0 = time.time()
    if n_ary == 3:
        sub = 64 << 20   # the trit packer of the oracle is single-threaded: compare a prefix of the payload, decode everything
        want, wtr = O.pack_trits(host[:sub], el, ev)
        pay, pst = dc.trit_pack(res.payload, nbits // 2)
        whole = (wtr // 5) - 1
        ok = int(pst.item()) == 0 and np.array_equal(pay[:whole].cpu().numpy(), want[:whole])
        stream, ust = dc.trit_unpack(pay, nbits // 2)
        ok = ok and int(ust.item()) == 0
    else:
        want, wbits, _ = O.pack_mt(host, el, ev, O.bits_per_digit(n_ary), 0, threads=threads, out=np.empty(n + n // 2 + 64, dtype=np.uint8))
        ok = nbits == wbits and np.array_equal(res.payload[: (nbits + 7) // 8].cpu().numpy(), want[: (nbits + 7) // 8])
        stream = res.payload
    out, status = dc.huff_decode(stream, nbits, table, n)
    ok = ok and int(status.item()) == 0 and torch.equal(out, data)
    print("n", n_ary, "zipf", s_exp, "max_bits", t.max_bits, "bits/sym", round(nbits / n, 3), "OK" if ok else "MISMATCH", f"(oracle {time.time() - t0:.1f} s)", flush=True)
    fails += 0 if ok else 1
    del data, res, out, stream
sys.exit(1 if fails else 0)
