"""Synthetic inputs of BASELINE.json / SURVEY 8d, identical on host (numpy) and device (dc_synth_fill).

byte i = value_base + rank_i, rank_i = #{k : thresholds[k] <= (splitmix64(seed + i) >> 32)}, where
`thresholds` is the u32 inverse CDF of Zipf(s) over `nranks` ranks (P(rank r) ~ (r+1)^-s, r = 0..nranks-1).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0x5EED0001  # + config index


def zipf_thresholds(nranks: int, s: float = 1.1) -> np.ndarray:
    """nranks-1 ascending u32 thresholds: rank = number of thresholds <= u, u uniform in [0, 2^32)."""
    w = np.arange(1, nranks + 1, dtype=np.float64) ** (-s)
    cdf = np.cumsum(w) / np.sum(w)
    thr = np.floor(cdf[:-1] * 4294967296.0)
    return np.minimum(thr, 4294967295.0).astype(np.uint32)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def host_stream(n: int, seed: int, thresholds: np.ndarray, value_base: int, start: int = 0) -> np.ndarray:
    """numpy twin of dc_synth_fill: bytes [start, start+n) of the stream."""
    out = np.empty(n, dtype=np.uint8)
    chunk = 1 << 22
    thr = thresholds.astype(np.uint64)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        with np.errstate(over="ignore"):
            idx = np.arange(start + lo, start + hi, dtype=np.uint64) + np.uint64(seed)
        r = splitmix64(idx) >> np.uint64(32)
        out[lo:hi] = (np.searchsorted(thr, r, side="right") + value_base).astype(np.uint8)
    return out


# the three named distributions
def zipf_bytes_spec():      # configs 3-5: ranks 1..255 -> byte = rank (never 0x00)
    return zipf_thresholds(255), 1


def zipf_7bit_spec():       # config 1: ranks 1..126 so the unmodified reference histogram() can run
    return zipf_thresholds(126), 1


def zipf_nybble_spec():     # config 2: 4-bit symbols 0..15
    return zipf_thresholds(16), 0


def device_thresholds(thresholds: np.ndarray, device):
    import torch
    return torch.from_numpy(thresholds.astype(np.int64).astype(np.uint32).view(np.int32).copy()).to(device)
