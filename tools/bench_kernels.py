#!/usr/bin/env python
"""Per-kernel timings (CUDA events, inputs resident in HBM, > L2) for every kernel and histogram layout.
Development tool: prints one JSON object per line; bench.py is the contract benchmark."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc  # noqa: E402
from data_compression_b200 import synth  # noqa: E402

PEAK = 6537.3
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def report(name, alg_bytes, n, best, med, **kw):
    print(json.dumps({"kernel": name, "ms_best": round(best, 4), "ms_median": round(med, 4),
                      "uncompressed_GBps": round(n / best / 1e6, 1), "algorithmic_GBps": round(alg_bytes / best / 1e6, 1),
                      "frac_of_hbm_peak": round(alg_bytes / best / 1e6 / PEAK, 4), **kw}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size-mib", type=int, default=1024)
    ap.add_argument("--radices", default="2,3,4,16")
    ap.add_argument("--hist-variants", default="0,1")
    args = ap.parse_args()
    n = args.size_mib << 20
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    thr, base = synth.zipf_bytes_spec()
    data = torch.empty(n, dtype=torch.uint8, device=dev)
    dc.synth_fill(data, synth.SEED_BASE + 2, synth.device_thresholds(thr, dev), base)
    hist = torch.empty(259, dtype=torch.int64, device=dev)

    for v in [int(x) for x in args.hist_variants.split(",") if x != ""]:
        best, med = timeit(lambda: dc.histogram(data, out=hist, variant=v))
        report(f"histogram[variant {v}]", n, n, best, med)
    uni = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev)
    for v in [int(x) for x in args.hist_variants.split(",") if x != ""]:
        best, med = timeit(lambda: dc.histogram(uni, out=hist, variant=v))
        report(f"histogram[variant {v}, uniform bytes]", n, n, best, med)
    del uni
    dc.histogram(data, out=hist)

    L = dc.lib()
    payload = torch.empty(n + n // 4 + 64, dtype=torch.uint8, device=dev)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    ws = torch.empty(L.dc_huff_encode_workspace_bytes(n), dtype=torch.uint8, device=dev)
    for n_ary in [int(x) for x in args.radices.split(",") if x != ""]:
        table = dc.HuffTable(dev)
        best, med = timeit(lambda: dc.huff_build(hist, n_ary, table))
        report(f"table[n={n_ary}]", 259 * 8, 0, best, med, max_bits=table.download().max_bits)
        res = dc.huff_encode(data, table, out=payload, workspace=ws)
        nbits = res.bits()
        c = (nbits + 7) // 8
        import ctypes as C
        L.dc_profile_reset(); L.dc_profile_enable(1)
        best, med = timeit(lambda: dc.huff_encode(data, table, out=payload, workspace=ws))
        L.dc_profile_enable(0)
        parts = {}
        for kid in range(64):
            ms, cnt = C.c_double(0), C.c_uint64(0)
            L.dc_profile_kernel(kid, C.byref(ms), C.byref(cnt))
            if cnt.value:
                parts[L.dc_profile_kernel_name(kid).decode()] = [round(ms.value / cnt.value, 4), cnt.value]
        report(f"encode[n={n_ary}]", n + c, n, best, med, compressed_ratio=round(c / n, 4), kernels_ms_avg_and_launches=parts)
        # the planned path: histogram with run histograms -> table -> plan + encode
        hist2 = torch.empty(259, dtype=torch.int64, device=dev)
        best, med = timeit(lambda: dc.histogram_runs(data, ws, out=hist2))
        report(f"histogram_runs", n, n, best, med)
        assert torch.equal(hist2, hist)
        ref_payload = payload[:c].clone()
        payload.zero_()
        L.dc_profile_reset(); L.dc_profile_enable(1)
        best, med = timeit(lambda: dc.huff_encode(data, table, out=payload, workspace=ws, planned=True))
        L.dc_profile_enable(0)
        parts = {}
        for kid in range(64):
            ms, cnt = C.c_double(0), C.c_uint64(0)
            L.dc_profile_kernel(kid, C.byref(ms), C.byref(cnt))
            if cnt.value:
                parts[L.dc_profile_kernel_name(kid).decode()] = [round(ms.value / cnt.value, 4), cnt.value]
        assert torch.equal(payload[:c], ref_payload), "planned encode differs"
        report(f"encode_planned[n={n_ary}]", n + c, n, best, med, kernels_ms_avg_and_launches=parts)
        del ref_payload
        dws = torch.empty(L.dc_huff_decode_workspace_bytes(0, nbits), dtype=torch.uint8, device=dev)
        st = torch.empty(1, dtype=torch.int32, device=dev)
        L.dc_profile_reset(); L.dc_profile_enable(1)
        best, med = timeit(lambda: dc.huff_decode(payload, nbits, table, n, out=out, workspace=dws, status=st), reps=3, warm=1)
        L.dc_profile_enable(0)
        assert int(st.item()) == 0 and torch.equal(out, data)
        parts = {}
        import ctypes as C
        for kid in range(64):
            ms, cnt = C.c_double(0), C.c_uint64(0)
            L.dc_profile_kernel(kid, C.byref(ms), C.byref(cnt))
            if cnt.value:
                parts[L.dc_profile_kernel_name(kid).decode()] = [round(ms.value / cnt.value, 4), cnt.value]
        report(f"decode[n={n_ary}]", n + c, n, best, med, kernels_ms_avg_and_launches=parts)
        del dws
        if n_ary == 3:  # the kernels' stream has 2 bits per trit; K7 converts to / from the 5-trits-per-byte payload
            ntr = nbits // 2
            pay = torch.empty((ntr + 4) // 5 + 16, dtype=torch.uint8, device=dev)
            best, med = timeit(lambda: dc.trit_pack(payload, ntr, out=pay))
            report("trit_pack[n=3]", c + (ntr + 4) // 5, n, best, med, payload_ratio=round((ntr + 4) // 5 / n, 4))
            best, med = timeit(lambda: dc.trit_unpack(pay, ntr))
            report("trit_unpack[n=3]", c + (ntr + 4) // 5, n, best, med)
            del pay

    thr4, base4 = synth.zipf_nybble_spec()
    sym = data
    dc.synth_fill(sym, synth.SEED_BASE + 1, synth.device_thresholds(thr4, dev), base4)
    packed = payload[: n // 2]
    stt = torch.empty(1, dtype=torch.int32, device=dev)
    best, med = timeit(lambda: dc.nybble_pack(sym, out=packed, status=stt))
    report("nybble_pack", n + n // 2, n, best, med)
    best, med = timeit(lambda: dc.nybble_unpack(packed, n, out=out))
    report("nybble_unpack", n + n // 2, n, best, med)
    assert torch.equal(out, sym)
    # static-table nybble compressor on text-like bytes: ~62 % of the characters are one of " etaoins"
    g = torch.Generator(device=dev); g.manual_seed(7)
    letters = torch.tensor(list(b" etaoins"), dtype=torch.uint8, device=dev)
    others = torch.tensor([c for c in range(33, 127) if c not in b" etaoins"], dtype=torch.uint8, device=dev)
    pick = torch.rand(n, device=dev, generator=g) < 0.62
    text = torch.where(pick, letters[torch.randint(0, 8, (n,), device=dev, generator=g)],
                       others[torch.randint(0, others.numel(), (n,), device=dev, generator=g)])
    del pick
    buf, ln, stt2 = dc.nybble_text_compress(text)
    clen = int(ln.item())
    assert int(stt2.item()) == 0
    best, med = timeit(lambda: dc.nybble_text_compress(text), reps=3, warm=1)
    report("nybble_text_compress", 2 * n + clen, n, best, med, compressed_ratio=round(clen / n, 4))
    comp = buf[:clen].clone()
    del buf
    back, bl, _ = dc.nybble_text_decompress(comp)
    assert int(bl.item()) == n and torch.equal(back[:n], text)
    del back
    best, med = timeit(lambda: dc.nybble_text_decompress(comp), reps=3, warm=1)
    report("nybble_text_decompress", 2 * clen + n, n, best, med)
    del comp
    # the adaptive (move-to-front) mode on the same text: K8 positions + K6
    buf, ln, stt2 = dc.nybble_adaptive_compress(text)
    alen = int(ln.item())
    assert int(stt2.item()) == 0
    del buf
    best, med = timeit(lambda: dc.nybble_adaptive_compress(text), reps=3, warm=1)
    report("nybble_adaptive_compress", 2 * n + alen, n, best, med, compressed_ratio=round(alen / n, 4))
    del text
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    best, med = timeit(lambda: a.copy_(sym))
    report("torch copy_ (reference point)", 2 * n, n, best, med)


if __name__ == "__main__":
    main()
