#!/usr/bin/env python
"""Profiling target: runs every kernel of the hot path a fixed, small number of times on one GPU
(2 rounds: the first warms up, ncu captures the second with -s).  Usage:
    ncu --set full -k regex:<kernel> -s <first-round launches of that kernel> -c 1 ... python tools/profile_target.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc  # noqa: E402
from data_compression_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size-mib", type=int, default=256)
    ap.add_argument("--n-ary", type=int, default=4)
    ap.add_argument("--rounds", type=int, default=2)
    args = ap.parse_args()
    n = args.size_mib << 20
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    thr, base = synth.zipf_bytes_spec()
    data = torch.empty(n, dtype=torch.uint8, device=dev)
    dc.synth_fill(data, synth.SEED_BASE + 2, synth.device_thresholds(thr, dev), base)
    thr4, base4 = synth.zipf_nybble_spec()
    sym = torch.empty(n, dtype=torch.uint8, device=dev)
    dc.synth_fill(sym, synth.SEED_BASE + 1, synth.device_thresholds(thr4, dev), base4)
    payload = torch.empty(n + n // 4 + 64, dtype=torch.uint8, device=dev)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    packed = torch.empty(n // 2, dtype=torch.uint8, device=dev)
    for _ in range(args.rounds):
        ws = dc.encode_workspace(n, dev)
        hist = dc.histogram_runs(data, ws)
        table = dc.huff_build(hist, args.n_ary)
        res = dc.huff_encode(data, table, out=payload, workspace=ws, planned=True)
        nbits = res.bits()
        back, st = dc.huff_decode(payload, nbits, table, n, out=out)
        assert int(st.item()) == 0
        dc.nybble_pack(sym, out=packed)
        dc.nybble_unpack(packed, n, out=out)
    torch.cuda.synchronize()
    print("profile target done", nbits)


if __name__ == "__main__":
    main()
