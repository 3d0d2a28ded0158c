#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native n-ary Huffman hot path.

One "step" = one pass of the hot path over one batch of synthetic input that is already resident in HBM:
    encode  : byte histogram -> (all-reduce over ranks) -> code table -> single-pass payload encode
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

def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import data_compression_b200 as dc
    from data_compression_b200 import synth

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

    n = args.size_mib << 20
    n_ary = args.n_ary
    thr, base = synth.zipf_bytes_spec()
    d_thr = synth.device_thresholds(thr, dev)
    data = torch.empty(n, dtype=torch.uint8, device=dev)
    # one logical stream: rank r holds bytes [r*n, (r+1)*n)
    dc.lib().dc_synth_fill(data.data_ptr(), n, synth.SEED_BASE + CONFIG_INDEX + rank * n, d_thr.data_ptr(), d_thr.numel(),
                           base, torch.cuda.current_stream().cuda_stream)
    payload = torch.empty(n + n // 4 + 64, dtype=torch.uint8, device=dev)
    decoded = torch.empty(n, dtype=torch.uint8, device=dev)
    enc_ws = torch.empty(L.dc_huff_encode_workspace_bytes(n), dtype=torch.uint8, device=dev)
    hist = torch.empty(dc.DC_NSLOTS, dtype=torch.int64, device=dev)
    ghist = torch.empty(dc.DC_NSLOTS, dtype=torch.int64, device=dev)
    table = dc.HuffTable(dev)
    my_bits = torch.empty(1, dtype=torch.int64, device=dev)
    all_bits = torch.empty(world, dtype=torch.int64, device=dev)
    dec_status = torch.empty(1, dtype=torch.int32, device=dev)
    state = {"nbits": None, "phase": 0, "dec_ws": None}

    def encode_path():
        dc.histogram(data, out=hist)
        if world > 1:
            ghist.copy_(hist)
            dist.all_reduce(ghist)                      # one global code table (2 KB over NVLink)
            dc.huff_build(ghist, n_ary, table)
            dc.huff_bits_for_hist(hist, table, out=my_bits)
            dist.all_gather_into_tensor(all_bits, my_bits)  # global bit-offset scan
            if state["nbits"] is None:
                ab = all_bits.cpu().numpy()
                state["phase"] = int(ab[:rank].sum() % 8)
                state["nbits"] = int(ab[rank])
        else:
            dc.huff_build(hist, n_ary, table)
        return dc.huff_encode(data, table, out=payload, bit_phase=state["phase"], workspace=enc_ws)

    def decode_path():
        dc.huff_decode(payload, state["nbits"], table, n, bit_start=state["phase"], out=decoded, workspace=state["dec_ws"],
                       status=dec_status)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # first pass: learn the bit count (side information a container header carries), check the round trip
    res = encode_path()
    if state["nbits"] is None:
        state["nbits"] = res.bits()
    dc._lib.check(int(res.status.item()), "encode")
    state["dec_ws"] = torch.empty(L.dc_huff_decode_workspace_bytes(state["phase"], state["nbits"]), dtype=torch.uint8, device=dev)
    decode_path()
    dc._lib.check(int(dec_status.item()), "decode")
    assert torch.equal(decoded, data), "round trip failed"
    c_bytes = (state["nbits"] + state["phase"] + 7) // 8

    for _ in range(max(args.warmup - 1, 0)):
        encode_path(); decode_path()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    t = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms = (float(x) for x in t.cpu())

    # ---- per-kernel durations (CUDA events around every launch, on the launching stream)
    kern = {}
    for kid in range(32):
        ms, cnt = C.c_double(0), C.c_uint64(0)
        L.dc_profile_kernel(kid, C.byref(ms), C.byref(cnt))
        if cnt.value:
            kern[L.dc_profile_kernel_name(kid).decode()] = {"ms_total": ms.value, "launches": cnt.value,
                                                            "ms_avg": ms.value / cnt.value}
    alg_bytes = {"histogram": n, "encode_count": n, "encode": n + c_bytes, "decode_sync": c_bytes, "decode_write": c_bytes + n,
                 "decode_fast_sync": c_bytes, "decode_fast_write": c_bytes + n}
    peak, peak_src = _peaks()
    dom = max((k for k in kern if k in alg_bytes), key=lambda k: kern[k]["ms_total"])
    achieved = alg_bytes[dom] / (kern[dom]["ms_avg"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes[dom], "peak_source": peak_src}

    # ---- end to end: host buffers (pinned) through the C-ABI, copies inside the timed region
    e2e = None
    if args.e2e_steps > 0:
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.copy_(data)
        h_payload = torch.empty(n + n // 4 + 64, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        del payload, decoded  # the host entry points own their device arena
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
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        cb = (bits + 7) // 8
        e2e = {"value": n * world * args.e2e_steps / float(dt.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n + cb),
               "d2h_bytes_per_step": int(cb + n + 259 * 4), "steps": args.e2e_steps,
               "api": "dc_host_huff_compress + dc_host_huff_decompress (pinned host buffers)"}

    if rank == 0:
        secs = total_ms * 1e-3
        line = {
            "metric": METRIC, "value": n * world * args.steps / secs / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"n={n_ary} Huffman encode+decode of {args.size_mib} MiB Zipf(1.1) bytes per GPU "
                                   f"(BASELINE configs[{CONFIG_INDEX}])",
                       "n_ary": n_ary, "bytes_per_gpu": n, "compressed_bytes_per_gpu": int(c_bytes),
                       "parallelism": f"dp{world}: contiguous shards, histogram all-reduce + bit-offset all-gather",
                       "l2": "inputs larger than L2 (1 GiB vs 126 MB); no explicit flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline,
            "breakdown": {
                "encode_gbs": n * world * args.steps / (enc_ms * 1e-3) / 1e9,
                "decode_gbs": n * world * args.steps / (dec_ms * 1e-3) / 1e9,
                "encode_path_frac_of_hbm": (2 * n + c_bytes) * args.steps / (enc_ms * 1e-3) / 1e9 / peak,
                "decode_path_frac_of_hbm": (c_bytes + n) * args.steps / (dec_ms * 1e-3) / 1e9 / peak,
                "kernels": kern,
            },
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
