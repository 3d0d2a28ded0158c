#!/usr/bin/env python
"""Randomised round trips through the netstring container (host/container.c over the GPU entry points); run on the GPU box."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from data_compression_b200 import hostapi
L = ctypes.CDLL(os.path.join(ROOT, "data_compression_b200", "libdc_b200_refapi.so"))
L.dc_container_compress.restype = ctypes.c_size_t
L.dc_container_compress.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
L.dc_container_decompress.restype = ctypes.c_size_t
L.dc_container_decompress.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(int(os.environ.get("SEED", 4)))
t_end = time.time() + budget
it = fails = 0
alphabets = [bytes(range(1, 127)), b"0123456789:,\n XZ#", b" etaoinsrhldcu.,\n", b"ab", b"\n", b"x"]
while time.time() < t_end:
    it += 1
    a = np.frombuffer(alphabets[int(rng.integers(len(alphabets)))], dtype=np.uint8)
    size = int(rng.choice([1, 2, 50, 1000, 32767, 32768, 32769, 70001, 150000]))
    p = rng.random(a.size) ** float(rng.choice([1.0, 4.0])); p /= p.sum()
    text = rng.choice(a, size=size, p=p).astype(np.uint8).tobytes()
    radix = int(rng.choice([2, 3, 4, 16, 10]))
    lengths = hostapi.huffman(hostapi.histogram(text), radix).astype(np.int32)
    cap = 2 * len(text) + 8192
    out = ctypes.create_string_buffer(cap + 1)
    n = L.dc_container_compress(radix, lengths.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), text, len(text), out, cap)
    ok = n != ctypes.c_size_t(-1).value
    if ok:
        back = ctypes.create_string_buffer(len(text) + 2)
        m = L.dc_container_decompress(radix, out.raw[:n], n, back, len(text) + 2)
        ok = m == len(text) and back.raw[:m] == text
    if not ok:
        fails += 1
        print("MISMATCH", it, radix, size, len(a), n, flush=True)
print("iterations", it, "mismatches", fails)
sys.exit(1 if fails else 0)
