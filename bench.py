#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native n-ary Huffman hot path.

One "step" = one pass of the hot path over one batch of synthetic input that is already resident in HBM:
    encode  : byte histogram (+ run histograms) -> (all-gather of the histograms over ranks) -> code table -> planned single-pass encode
    decode  : self-synchronising parallel decode of that payload back to the bytes
on BASELINE.json's 1-GPU Huffman configuration (configs[2]: n=4, 1 GiB of Zipf(1.1) bytes per GPU).
`value` is uncompressed GB/s through the encode+decode round trip, aggregated over all ranks (weak scaling:
every rank holds its own 1 GiB shard of one logical stream; the only data-path exchange is the 2 KB
histogram all-reduce and an all-gather of one u64 bit total per rank).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Prints ONE JSON line (see the task contract): metric/value/unit, ms_per_step, clocks, e2e (host buffers
through the C-ABI, copies inside the timed region), gpu_launches, roofline (dominant kernel, measured live
with CUDA events around every launch), cpu_baseline (the oracle port on the host cores, bounded sample).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "huffman_encode_decode_roundtrip_uncompressed_GBps"
UNIT = "GB/s"
CONFIG_INDEX = 2  # BASELINE.json configs[2]


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML (the recipe's `nvidia-smi --query-gpu=clocks.sm,...,
    clocks_event_reasons.*` line, read in-process every 4 ms).  The thread is started before warm-up (nvmlInit takes
    longer than a short timed region); begin()/stop() bracket the timed region and only its samples are reported."""

    def __init__(self, gpu_index: int):
        self.gpu, self.samples, self.stop_flag, self.thread, self.max_mhz = gpu_index, [], False, None, None
        self.err, self.t_begin, self.t_end, self.ready = None, None, None, threading.Event()

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.ready.set()
            while not self.stop_flag:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
                time.sleep(0.004)
        except Exception as e:  # noqa: BLE001 - the bench must not die because NVML is unavailable
            self.err = repr(e)
            self.ready.set()

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def begin(self):
        self.ready.wait(timeout=10)
        self.t_begin = time.perf_counter()

    def stop(self) -> dict:
        self.t_end = time.perf_counter()
        time.sleep(0.01)
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        inside = [x for x in self.samples if self.t_begin is not None and self.t_begin <= x[0] <= self.t_end + 0.005]
        reasons = sorted({name for _, _, r in inside for bit, name in names.items() if r & bit})
        out = {"sm_mhz": float(np.median([x[1] for x in inside])) if inside else None, "sm_max_mhz": self.max_mhz,
               "samples": len(inside), "reasons": reasons}
        if self.err:
            out["error"] = self.err
        return out


# --------------------------------------------------------------------------------------------- CPU legs

def _cpu_roundtrip(O, sample: np.ndarray, n_ary: int, threads: int):
    """One encode+decode of `sample` with the oracle port on `threads` host threads; returns seconds (enc, dec)."""
    bpd = O.bits_per_digit(n_ary)
    block = 1 << 16
    t0 = time.perf_counter()
    hist = O.histogram_u8(sample, threads=threads)
    lengths, el, ev, st = O.build_tables(hist, n_ary)
    out = np.empty(sample.size + sample.size // 4 + 64, dtype=np.uint8)
    payload, bits, offs = O.pack_mt(sample, el, ev, bpd, 0, threads=threads, block_symbols=block, out=out)
    t1 = time.perf_counter()
    back = np.empty(sample.size, dtype=np.uint8)
    O.unpack_mt(payload, 0, lengths, n_ary, sample.size, offs, block, threads, out=back)
    t2 = time.perf_counter()
    assert np.array_equal(back[:4096], sample[:4096]) and np.array_equal(back[-4096:], sample[-4096:])
    return t1 - t0, t2 - t1


def _host_sample(nbytes: int, rank_offset: int = 0) -> np.ndarray:
    from data_compression_b200 import synth
    thr, base = synth.zipf_bytes_spec()
    return synth.host_stream(nbytes, synth.SEED_BASE + CONFIG_INDEX, thr, base, start=rank_offset)


def cpu_baseline(n_ary: int, budget_s: float = 12.0) -> dict:
    from oracle import pyoracle as O
    O.build()
    threads = os.cpu_count() or 1
    probe = _host_sample(16 << 20)
    te, td = _cpu_roundtrip(O, probe, n_ary, threads)
    rate = probe.size / (te + td)
    nbytes = int(min(512 << 20, max(32 << 20, rate * budget_s)))
    nbytes -= nbytes % (1 << 20)
    sample = probe if nbytes <= probe.size else _host_sample(nbytes)
    te, td = _cpu_roundtrip(O, sample, n_ary, threads)
    return {"value": sample.size / (te + td) / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample.size >> 20} MiB of the same Zipf(1.1) stream, one encode+decode, OpenMP oracle "
                      f"(decode is given the encoder's block offsets)",
            "encode_gbs": sample.size / te / 1e9, "decode_gbs": sample.size / td / 1e9}


def run_reference(args) -> None:
    """--impl reference: the path's CPU implementation on the host cores.  The reference itself cannot encode
    or decode (represent_items_with_codes and the Huffman block decoder are assert(0) stubs), so the
    whole path is the oracle port (kind "port"); the reference functions that do exist are timed beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as O
    O.build()
    threads = os.cpu_count() or 1
    sample = _host_sample(args.ref_sample_mib << 20)
    for _ in range(args.warmup):
        _cpu_roundtrip(O, sample[: 8 << 20], args.n_ary, threads)
    t0 = time.perf_counter()
    te = td = 0.0
    for _ in range(args.steps):
        a, b = _cpu_roundtrip(O, sample, args.n_ary, threads)
        te += a; td += b
    total = time.perf_counter() - t0
    value = sample.size * args.steps / total / 1e9
    extra = {}
    if O.have_ref():
        from data_compression_b200 import synth
        thr, base = synth.zipf_7bit_spec()  # the unmodified histogram() prints on bytes > 126
        text = synth.host_stream(64 << 20, synth.SEED_BASE + 1, thr, base).tobytes()
        t = time.perf_counter(); h = O.ref_histogram(text); th = time.perf_counter() - t
        t = time.perf_counter(); O.ref_huffman(h, args.n_ary); tt = time.perf_counter() - t
        extra = {"reference_histogram_gbs_1thread": len(text) / th / 1e9, "reference_huffman_ms": tt * 1e3}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"n={args.n_ary} Huffman encode+decode of Zipf(1.1) bytes (BASELINE configs[{CONFIG_INDEX}])",
                   "n_ary": args.n_ary, "bytes_per_step": int(sample.size),
                   "note": "bounded sample of the 1 GiB workload per step; CPU port of the path on all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample.size >> 20} MiB per step", "encode_gbs": sample.size * args.steps / te / 1e9,
                         "decode_gbs": sample.size * args.steps / td / 1e9, **extra},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# --------------------------------------------------------------------------------------------- GPU arm

def _kernel_times(L) -> dict:
    kern = {}
    for kid in range(64):
        ms, cnt = C.c_double(0), C.c_uint64(0)
        if L.dc_profile_kernel(kid, C.byref(ms), C.byref(cnt)) != 0:
            break
        if cnt.value:
            kern[L.dc_profile_kernel_name(kid).decode()] = {"ms_total": ms.value, "launches": cnt.value, "ms_avg": ms.value / cnt.value}
    return kern


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import data_compression_b200 as dc
    from data_compression_b200 import shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: data_compression_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = dc.lib()
    peak, peak_src = _peaks()

    n = args.size_mib << 20
    n_ary = args.n_ary
    thr, base = synth.zipf_bytes_spec()
    d_thr = synth.device_thresholds(thr, dev)
    seed0 = synth.SEED_BASE + CONFIG_INDEX

    def fill(t, start):   # bytes [start, start + len(t)) of the one logical stream
        L.dc_synth_fill(t.data_ptr(), t.numel(), seed0 + start, d_thr.data_ptr(), d_thr.numel(), base, torch.cuda.current_stream().cuda_stream)

    data = torch.empty(n, dtype=torch.uint8, device=dev)
    fill(data, rank * n)   # rank r holds bytes [r*n, (r+1)*n)
    decoded = torch.empty(n, dtype=torch.uint8, device=dev)
    dec_status = torch.empty(1, dtype=torch.int32, device=dev)
    state = {"nbits": None, "phase": 0, "dec_ws": None}
    S = shard.NcclShards(dev) if world > 1 else None   # dc_shard_*: NCCL issued by the library (C-ABI)

    def all_max(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        payload = torch.empty(n + n // 4 + 64, dtype=torch.uint8, device=dev)
        enc_ws = dc.encode_workspace(n, dev)
        hist = torch.empty(dc.DC_NSLOTS, dtype=torch.int64, device=dev)
        table = dc.HuffTable(dev)

        def encode_path():
            # histogram (+ one small histogram per 32 KB run) -> table -> planned single-pass encode
            dc.histogram_runs(data, enc_ws, out=hist)
            dc.huff_build(hist, n_ary, table)
            return dc.huff_encode(data, table, out=payload, workspace=enc_ws, planned=True)
    else:
        ebuf = S.encode_buffers(n)
        payload, table = ebuf["out"], ebuf["table"]

        def encode_path():
            # ONE C call per rank: local histogram -> all-gather of the histograms (NCCL) -> global table and every rank's
            # bit total -> encode at this rank's bit phase -> shared edge bytes completed from the neighbours' edge symbols
            S.encode(data, n_ary, ebuf)
            return None

    def decode_path():
        dc.huff_decode(payload, state["nbits"], table, n, bit_start=state["phase"], out=decoded, workspace=state["dec_ws"],
                       status=dec_status)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # first pass: learn the bit count (side information a container header carries), check the round trip
    res = encode_path()
    if world == 1:
        state["nbits"] = res.bits()
    else:
        off, bits, total = S.encode_info(ebuf, n)
        state["nbits"], state["phase"] = bits, off % 8
    state["dec_ws"] = torch.empty(L.dc_huff_decode_workspace_bytes(state["phase"], state["nbits"]), dtype=torch.uint8, device=dev)
    decode_path()
    dc._lib.check(int(dec_status.item()), "decode")
    assert torch.equal(decoded, data), "round trip failed"
    c_bytes = (state["nbits"] + state["phase"] + 7) // 8

    for _ in range(max(args.warmup - 1, 0)):
        encode_path(); decode_path()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    L.dc_profile_reset(); L.dc_profile_enable(1)
    launches0 = dc.launch_count()
    barrier()
    if rank == 0:
        sampler.begin()
    ev[0].record()
    for s in range(args.steps):
        encode_path()
        ev[2 * s + 1].record()
        decode_path()
        ev[2 * s + 2].record()
    barrier()
    launches = dc.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    L.dc_profile_enable(0)
    total_ms = ev[0].elapsed_time(ev[-1])
    enc_ms = sum(ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps))
    dec_ms = sum(ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps))
    dc._lib.check(int(dec_status.item()), "decode")
    total_ms, enc_ms, dec_ms = all_max(total_ms, enc_ms, dec_ms)

    # ---- per-kernel durations (CUDA events around every launch, on the launching stream)
    kern = _kernel_times(L)
    alg_bytes = {"histogram": n, "encode_count": n, "encode": n + c_bytes, "encode_fast": n + c_bytes, "decode_sync": c_bytes,
                 "decode_write": c_bytes + n, "decode_fast_sync": c_bytes, "decode_fast_write": c_bytes + n,
                 "decode_fsm_sync": c_bytes, "decode_fsm_write": c_bytes + n}
    dom = max((k for k in kern if k in alg_bytes), key=lambda k: kern[k]["ms_total"])
    achieved = alg_bytes[dom] / (kern[dom]["ms_avg"] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get(dom), tj.get("_source", "profiles/ncu_traffic.json (one ncu --set full capture, not measured in this run)")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes[dom], "peak_source": peak_src}
    breakdown = {
        "encode_gbs": n * world * args.steps / (enc_ms * 1e-3) / 1e9,
        "decode_gbs": n * world * args.steps / (dec_ms * 1e-3) / 1e9,
        "encode_ms": enc_ms / args.steps, "decode_ms": dec_ms / args.steps,
        "encode_path_frac_of_hbm": (2 * n + c_bytes) * args.steps / (enc_ms * 1e-3) / 1e9 / peak,
        "decode_path_frac_of_hbm": (c_bytes + n) * args.steps / (dec_ms * 1e-3) / 1e9 / peak,
        "kernels": kern,
    }

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return all_max(a.elapsed_time(b) / reps)[0]

    # ---- opt-in: a stream kept with its index (what the decoder's first pass finds): the write pass alone
    if not args.no_extras:
        nb, ph = state["nbits"], state["phase"]
        built = dc.huff_index_build(payload, nb, table, n, bit_start=ph, workspace=state["dec_ws"])
        assert built is not None, "the bench stream self-synchronises"
        index, info = built
        t_build = timed(lambda: dc.huff_index_build(payload, nb, table, n, bit_start=ph, workspace=state["dec_ws"]), reps=3, warm=1)
        t_idx = timed(lambda: dc.huff_decode_indexed(payload, index, info, table, out=decoded, workspace=state["dec_ws"], status=dec_status))
        dc._lib.check(int(dec_status.item()), "decode_indexed")
        assert torch.equal(decoded, data), "indexed decode failed"
        breakdown["decode_with_index"] = {
            "what": "opt-in: the producer keeps sub_info + segment offsets (dc_huff_index_build), the consumer runs the write pass only "
                    "(dc_huff_decode_indexed); not part of `value`",
            "index_bytes": int(index.numel()), "index_frac_of_payload": index.numel() / float(c_bytes),
            "index_build_ms": t_build, "decode_ms": t_idx, "decode_gbs": n * world / (t_idx * 1e-3) / 1e9,
            "decode_frac_of_hbm": (c_bytes + index.numel() + n) / (t_idx * 1e-3) / 1e9 / peak}
        del index

    # ---- BASELINE configs[1]: nybble pack / unpack of 2^30 four-bit symbols per GPU (shards start on even symbols: no exchange)
    if not args.no_extras:
        thr4, base4 = synth.zipf_nybble_spec()
        d_thr4 = synth.device_thresholds(thr4, dev)
        sym = decoded   # reuse: n symbols, one per byte
        L.dc_synth_fill(sym.data_ptr(), n, synth.SEED_BASE + 1 + rank * n, d_thr4.data_ptr(), d_thr4.numel(), base4,
                        torch.cuda.current_stream().cuda_stream)
        n_total_sym = n * world
        if world == 1:
            packed = torch.empty(n // 2, dtype=torch.uint8, device=dev)
            st4 = torch.zeros(1, dtype=torch.int32, device=dev)
            back = torch.empty(n, dtype=torch.uint8, device=dev)
            t_pack = timed(lambda: dc.nybble_pack(sym, out=packed, status=st4))
            t_unpack = timed(lambda: dc.nybble_unpack(packed, n, out=back))
            assert torch.equal(back, sym) and int(st4.item()) == 0
        else:
            lo, hi = S.nybble_range(n_total_sym)
            assert (lo, hi) == (rank * n, (rank + 1) * n)
            packed = torch.empty(n // 2, dtype=torch.uint8, device=dev)
            st4 = torch.zeros(1, dtype=torch.int32, device=dev)
            back = torch.empty(n, dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream().cuda_stream
            t_pack = timed(lambda: L.dc_shard_nybble_pack(n_total_sym, rank, world, sym.data_ptr(), packed.data_ptr(), st4.data_ptr(), stream))
            t_unpack = timed(lambda: L.dc_shard_nybble_unpack(n_total_sym, rank, world, packed.data_ptr(), back.data_ptr(), stream))
            assert torch.equal(back, sym) and int(st4.item()) == 0
        breakdown["config2_nybble"] = {
            "symbols_per_gpu": n, "pack_ms": t_pack, "unpack_ms": t_unpack,
            "pack_gbs_of_symbols": n * world / (t_pack * 1e-3) / 1e9, "unpack_gbs_of_symbols": n * world / (t_unpack * 1e-3) / 1e9,
            "pack_frac_of_hbm": 1.5 * n / (t_pack * 1e-3) / 1e9 / peak, "unpack_frac_of_hbm": 1.5 * n / (t_unpack * 1e-3) / 1e9 / peak}
        del packed, back

    # ---- BASELINE configs[3] and [4] on N > 1 GPUs, through dc_shard_* (NCCL inside the library)
    if world > 1 and not args.no_extras:
        fill(data, rank * n)
        # config 4: n = 16 encode of one logical stream into ONE contiguous buffer on rank 0
        b16 = S.encode_buffers(n)
        t_enc16 = timed(lambda: S.encode(data, 16, b16), reps=5, warm=2)
        off16, bits16, total16 = S.encode_info(b16, n)
        S.gather(b16, n, total16, root=0)   # (the first exchange between two ranks sets up the NCCL peer connection)
        barrier()
        t0 = time.perf_counter()
        stream16 = S.gather(b16, n, total16, root=0)
        barrier()
        t_gather = all_max(time.perf_counter() - t0)[0] * 1e3
        check = None
        if rank == 0 and n * world <= (16 << 30):   # the same bytes encoded as ONE stream on one GPU must give the same payload
            whole = torch.empty(n * world, dtype=torch.uint8, device=dev)
            fill(whole, 0)
            p1, bits1, _ = dc.huff_compress(whole, 16)
            check = bool(bits1 == total16 and torch.equal(p1, stream16))
            del whole, p1
        breakdown["config4_n16_sharded_encode"] = {
            "bytes_per_gpu": n, "encode_ms": t_enc16, "encode_gbs": n * world / (t_enc16 * 1e-3) / 1e9,
            "frac_of_aggregate_hbm": (2 * n + (bits16 + 7) // 8) / (t_enc16 * 1e-3) / 1e9 / peak,
            "gather_into_one_buffer_ms": t_gather, "stream_bytes": (total16 + 7) // 8,
            "equals_single_gpu_stream": check,
            "collectives_per_encode": "one ncclAllGather (262 x u64 per rank: histogram, symbol count, first and last 8 symbols); shared edge bytes completed locally"}
        del b16, stream16
        torch.cuda.empty_cache()
        # config 5: n = 2 decode of ONE stream cut blindly into equal byte ranges, boundary sync between neighbours
        b2 = S.encode_buffers(n)
        S.encode(data, 2, b2)
        off2, bits2, total2 = S.encode_info(b2, n)
        total_bytes = (total2 + 7) // 8
        whole2 = torch.zeros(total_bytes + 16, dtype=torch.uint8, device=dev)
        got = S.gather(b2, n, total2, root=0, out=whole2 if rank == 0 else None)
        dist.broadcast(whole2, src=0)
        part = ((total_bytes + world - 1) // world + 1023) // 1024 * 1024
        plo, phi = min(rank * part, total_bytes), min((rank + 1) * part, total_bytes)
        dbuf = S.decode_buffers(part, n + n // 2)
        dbuf["buf"][1024: 1024 + (phi - plo)] = whole2[plo:phi]
        table2 = b2["table"]
        del whole2, got
        res5 = {}

        def dec5():
            res5["r"] = S.decode_stream(dbuf, part, total2, table2)
        t_dec5 = timed(dec5, reps=3, warm=1)
        sym5, soff5, stot5 = res5["r"]
        exp = torch.empty(sym5.numel(), dtype=torch.uint8, device=dev)
        fill(exp, soff5)
        ok5 = bool(stot5 == n * world and int(dbuf["status"].item()) == 0 and torch.equal(sym5, exp))
        okt = torch.tensor([1 if ok5 else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        breakdown["config5_n2_blind_cut_decode"] = {
            "stream_bytes": total_bytes, "bytes_per_gpu_of_stream": part, "decode_ms": t_dec5,
            "decode_gbs_uncompressed": n * world / (t_dec5 * 1e-3) / 1e9,
            "frac_of_aggregate_hbm": (total_bytes + n * world) / world / (t_dec5 * 1e-3) / 1e9 / peak,
            "output_equals_input": bool(int(okt.item()) == 1),
            "exchange": "1 KB halos by ncclSend/Recv, 24-byte summaries by ncclAllGather, host reads them (blocking call)"}
        del b2, dbuf, exp
        torch.cuda.empty_cache()

    # ---- end to end: host buffers (pinned) through the C-ABI, copies inside the timed region
    e2e = None
    if args.e2e_steps > 0:
        fill(data, rank * n)
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.copy_(data)
        h_payload = torch.empty(n + n // 4 + 64, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        # the floor of this box: the same bytes over PCIe with no kernels at all, every rank at once (H2D n, D2H c; then H2D c
        # and D2H n concurrently, as the pipelined decompress issues them)
        cb0 = c_bytes
        up, dn = torch.cuda.Stream(), torch.cuda.Stream()
        d_tmp = torch.empty(n, dtype=torch.uint8, device=dev)

        def copies():
            with torch.cuda.stream(up):
                d_tmp.copy_(h_in, non_blocking=True)
                h_payload[:cb0].copy_(payload[:cb0], non_blocking=True)
            up.synchronize()
            with torch.cuda.stream(up):
                payload[:cb0].copy_(h_payload[:cb0], non_blocking=True)
            with torch.cuda.stream(dn):
                h_out.copy_(d_tmp, non_blocking=True)
            up.synchronize(); dn.synchronize()
        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            copies()
        floor_s = all_max(time.perf_counter() - t0)[0]
        del d_tmp
        if world == 1:
            del payload
        del decoded  # the host entry points own their device arena
        torch.cuda.empty_cache()
        from data_compression_b200 import hostapi
        np_in, np_payload, np_out = h_in.numpy(), h_payload.numpy(), h_out.numpy()
        p, bits, lens = hostapi.huff_compress(np_in, n_ary, out=np_payload)           # warm-up (arena allocation)
        hostapi.huff_decompress(p, bits, lens, n_ary, n, out=np_out)
        assert np.array_equal(np_out[:65536], np_in[:65536])
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            p, bits, lens = hostapi.huff_compress(np_in, n_ary, out=np_payload)
            hostapi.huff_decompress(p, bits, lens, n_ary, n, out=np_out)
        torch.cuda.synchronize()
        dt = all_max(time.perf_counter() - t0)[0]
        cb = (bits + 7) // 8
        e2e = {"value": n * world * args.e2e_steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n + cb),
               "d2h_bytes_per_step": int(cb + n + 259 * 4), "steps": args.e2e_steps,
               "api": "dc_host_huff_compress + dc_host_huff_decompress (pinned host buffers)",
               "e2e_copy_floor": {"value": n * world * args.e2e_steps / floor_s / 1e9, "unit": UNIT,
                                  "what": "the same bytes moved with bare pinned cudaMemcpyAsync by all ranks at once, no kernels"}}

    peer_exchange = bool(L.dc_debug_shard_peer_active(S.comm)) if S is not None else False
    if S is not None:
        S.close()
    if rank == 0:
        secs = total_ms * 1e-3
        line = {
            "metric": METRIC, "value": n * world * args.steps / secs / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"n={n_ary} Huffman encode+decode of {args.size_mib} MiB Zipf(1.1) bytes per GPU "
                                   f"(BASELINE configs[{CONFIG_INDEX}])",
                       "n_ary": n_ary, "bytes_per_gpu": n, "compressed_bytes_per_gpu": int(c_bytes),
                       "parallelism": (f"dp{world}: contiguous shards of one logical stream through dc_shard_huff_encode (C-ABI, NCCL inside the "
                                       f"library): one all-gather (local histograms + edge symbols), bit offsets from the gathered histograms, shared edge bytes completed locally"
                                       if world > 1 else "1 GPU"),
                       "shard_exchange": (("peer memory (CUDA IPC: stores over NVLink + flags, one single-CTA kernel)"
                                           if peer_exchange else "ncclAllGather") if world > 1 else None),
                       "l2": "inputs larger than L2 (1 GiB vs 126 MB); no explicit flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline,
            "breakdown": breakdown,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n_ary)
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1), so
    fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def _emit(line: dict) -> None:
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


_OUT = sys.stdout


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-ary", type=int, default=4)
    ap.add_argument("--size-mib", type=int, default=1024, help="uncompressed MiB per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-sample-mib", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the breakdown lines of configs 1, 3 and 4")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
