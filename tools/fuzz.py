#!/usr/bin/env python
"""Randomised cross-check against the oracle (run on the GPU box): random alphabets, skews, radices and sizes through
histogram -> table -> encode -> decode, then random damage to the payload (must be reported or decode to something, never crash).
    python tools/fuzz.py [seconds]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from oracle import pyoracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(os.environ.get("SEED", 1)))
t_end = time.time() + budget
it = fails = 0
while time.time() < t_end:
    it += 1
    n_ary = int(rng.choice([2, 3, 4, 16, 5, 7, 9, 10, 13]))   # 5 .. 15: one nibble per digit
    nsym = int(rng.integers(1, 256))
    alphabet = rng.choice(np.arange(1, 256), size=nsym, replace=False).astype(np.uint8)
    skew = float(rng.choice([0.0, 0.5, 1.0, 1.5, 2.5, 4.0]))
    w = (np.arange(1, nsym + 1) ** -skew).astype(np.float64)
    size = int(rng.choice([1, 2, 7, 100, 4095, 4097, 32768, 100003, 1 << 20]))
    data = rng.choice(alphabet, size=size, p=w / w.sum()).astype(np.uint8)
    d = torch.from_numpy(data).cuda()
    hist = dc.histogram(d)
    o_hist = O.histogram_u8(data)
    assert np.array_equal(hist.cpu().numpy().astype(np.uint64), o_hist)
    # tables are exact for every radix (the payload only for 2, 3, 4, 16): one more radix per iteration, tables only
    n_other = int(rng.integers(5, 40))
    t_o = dc.huff_build(hist, n_other).download()
    ln_o, el_o, ev_o, st_o = O.build_tables(o_hist, n_other)
    if not (np.array_equal(np.array(t_o.lengths[:259]), ln_o) and ((st_o != 0) or np.array_equal(np.array(t_o.values[:259], dtype=np.uint32), ev_o))):
        fails += 1
        print("TABLE MISMATCH", it, n_other, nsym, skew, size, flush=True)
    table = dc.huff_build(hist, n_ary)
    t = table.download()
    ln, el, ev, st = O.build_tables(o_hist, n_ary)
    if st != 0 or t.status != 0:
        assert (st != 0) == (t.status != 0), (st, t.status)
        continue
    assert np.array_equal(np.array(t.lengths[:259]), ln) and np.array_equal(np.array(t.values[:259], dtype=np.uint32), ev)
    if t.max_bits > 32:
        continue
    res = dc.huff_encode(d, table, out=torch.empty(size * 4 + 64, dtype=torch.uint8, device="cuda"))
    nbits = res.bits()
    if n_ary == 3:
        want, wtr = O.pack_trits(data, el, ev)
        assert nbits == 2 * wtr
        pay, pst = dc.trit_pack(res.payload, wtr)
        ok = int(pst.item()) == 0 and np.array_equal(pay.cpu().numpy(), want)
        stream, ust = dc.trit_unpack(pay, wtr)
    else:
        nib = 5 <= n_ary < 16
        sv, bpd = (O.nibble_values(el, ev, n_ary), 4) if nib else (ev, O.bits_per_digit(n_ary))
        want, wbits = O.pack(data, el, sv, bpd)
        ok = nbits == wbits and np.array_equal(res.payload[: (nbits + 7) // 8].cpu().numpy(), want)
        stream = res.payload
    out, status = dc.huff_decode(stream, nbits, table, size)
    ok = ok and int(status.item()) == 0 and np.array_equal(out.cpu().numpy(), data)
    if not ok:
        fails += 1
        print("MISMATCH", it, n_ary, nsym, skew, size, t.max_bits, flush=True)
    # a shard's view: the same input at a non-zero bit phase (leading bits belong to the previous shard)
    if n_ary != 3:
        phase = 4 if nib else int(rng.integers(1, 8))   # (a nibble code cannot be decoded off the nibble grid)
        res_p = dc.huff_encode(d, table, out=torch.empty(size * 4 + 64, dtype=torch.uint8, device="cuda"), bit_phase=phase)
        want_p, wb = O.pack(data, el, sv, bpd, phase)
        okp = res_p.bits() == wb and np.array_equal(res_p.payload[: (wb + phase + 7) // 8].cpu().numpy(), want_p)
        out_p, st_p = dc.huff_decode(res_p.payload, wb, table, size, bit_start=phase)
        okp = okp and int(st_p.item()) == 0 and np.array_equal(out_p.cpu().numpy(), data)
        if not okp:
            fails += 1
            print("PHASE MISMATCH", it, n_ary, nsym, skew, size, t.max_bits, phase, flush=True)
    # damage: flip a few bytes of the stream, decode must terminate with some status
    if nbits >= 64:
        broken = stream[: (nbits + 7) // 8 + 64].clone()
        for _ in range(3):
            broken[int(rng.integers(0, (nbits + 7) // 8))] = int(rng.integers(0, 256))
        try:
            out, status = dc.huff_decode(broken, nbits, table, size)
            int(status.item())
        except dc.DcError:
            pass
print("iterations", it, "mismatches", fails)
sys.exit(1 if fails else 0)
