"""Per-kernel timings of the adaptive nybble compressor (K8 + K6) on text-like data; run on the GPU box."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
dev = torch.device("cuda:0"); n = int(os.environ.get("N", 1 << 28))
g = torch.Generator(device=dev); g.manual_seed(7)
letters = torch.tensor(list(b" etaoinsrhld"), dtype=torch.uint8, device=dev)
others = torch.tensor([c for c in range(33, 127) if c not in b" etaoinsrhld"], dtype=torch.uint8, device=dev)
pick = torch.rand(n, device=dev, generator=g) < 0.8
text = torch.where(pick, letters[torch.randint(0, letters.numel(), (n,), device=dev, generator=g)], others[torch.randint(0, others.numel(), (n,), device=dev, generator=g)])
del pick
L = dc.lib()
buf, ln, st = dc.nybble_adaptive_compress(text); clen = int(ln.item()); print("n", n, "compressed", clen, "status", int(st.item()))
small = buf[: min(clen, 1 << 20)].clone(); del buf
for name, fn in (("compress", lambda: dc.nybble_adaptive_compress(text)), ("decompress_1MiB", lambda: dc.nybble_adaptive_decompress(small))):
    L.dc_profile_reset(); L.dc_profile_enable(1)
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); L.dc_profile_enable(0)
    tot = 0.0
    for kid in range(40):
        ms, cnt = C.c_double(0), C.c_uint64(0)
        L.dc_profile_kernel(kid, C.byref(ms), C.byref(cnt))
        if cnt.value:
            print(name, L.dc_profile_kernel_name(kid).decode(), "ms/call", round(ms.value / cnt.value, 4), "calls", cnt.value); tot += ms.value / 3
    print(name, "total ms", round(tot, 3))
