#!/usr/bin/env python
"""Randomised cross-check of the host-buffer entry points (dc_host_huff_compress / _decompress, every radix) against the
oracle.  Run with DC_PIPE_CHUNK_MIB=1 so that streams of a few MiB take the chunked, overlapped decompress.
    DC_PIPE_CHUNK_MIB=1 python tools/fuzz_host.py [seconds]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from data_compression_b200 import hostapi
from oracle import pyoracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(os.environ.get("SEED", 3)))
t_end = time.time() + budget
it = fails = skipped = 0
while time.time() < t_end:
    it += 1
    n_ary = int(rng.choice([2, 3, 4, 16]))
    nsym = int(rng.integers(2, 256))
    alphabet = rng.choice(np.arange(1, 256), size=nsym, replace=False).astype(np.uint8)
    skew = float(rng.choice([0.0, 1.0, 1.5, 2.0, 3.0]))
    w = (np.arange(1, nsym + 1) ** -skew).astype(np.float64)
    size = int(rng.choice([1, 1000, 70001, (3 << 20) + 17, (5 << 20) + 4093, 8 << 20]))
    data = rng.choice(alphabet, size=size, p=w / w.sum()).astype(np.uint8)
    ln, el, ev, st = O.build_tables(O.histogram_u8(data), n_ary)
    try:
        payload, bits, lengths = hostapi.huff_compress(data, n_ary)
    except dc.DcError as e:
        assert st != 0 or int(ln.max()) * max(O.bits_per_digit(n_ary) if n_ary != 3 else 2, 1) > 32, (e.status, st)
        skipped += 1
        continue
    ok = st == 0 and np.array_equal(lengths, ln)
    if n_ary == 3:
        want, wtr = O.pack_trits(data, el, ev)
        ok = ok and bits == 2 * wtr and np.array_equal(payload, want)
    else:
        want, wbits = O.pack(data, el, ev, O.bits_per_digit(n_ary))
        ok = ok and bits == wbits and np.array_equal(payload, want)
    back = hostapi.huff_decompress(payload, bits, lengths, n_ary, size)
    ok = ok and np.array_equal(back, data)
    if not ok:
        fails += 1
        print("MISMATCH", it, n_ary, nsym, skew, size, int(ln.max()), flush=True)
print("iterations", it, "skipped (code too long)", skipped, "mismatches", fails)
sys.exit(1 if fails else 0)
