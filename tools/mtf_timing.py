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
# many short strings per call (one thread per string): 1M strings of 64..192 bytes cut from the same text
import numpy as np
cnt = 1 << 20
rng = np.random.default_rng(3)
lens = rng.integers(64, 193, size=cnt).astype(np.int64)
so = np.zeros(cnt + 1, dtype=np.int64); np.cumsum(lens, out=so[1:])
do = np.zeros(cnt + 1, dtype=np.int64); np.cumsum(lens + 2, out=do[1:])
d_so, d_do = torch.from_numpy(so).to(dev), torch.from_numpy(do).to(dev)
d_dst = torch.empty(int(do[-1]), dtype=torch.uint8, device=dev); d_len = torch.empty(cnt, dtype=torch.int64, device=dev)
st = torch.zeros(1, dtype=torch.int32, device=dev)
for modify in (0, 1):
    f = lambda: L.dc_nybble_text_compress_batch(text.data_ptr(), d_so.data_ptr(), cnt, modify, d_dst.data_ptr(), d_do.data_ptr(), d_len.data_ptr(), st.data_ptr(), None)
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); b.synchronize()
    ms = a.elapsed_time(b)
    print("batch compress modify", modify, "strings", cnt, "bytes", int(so[-1]), "ms", round(ms, 3), "GB/s", round(int(so[-1]) / ms / 1e6, 1), "status", int(st.item()))
    # decompress what was just written: compact the outputs into slots of 2 * len + 2
    clen = d_len.cpu().numpy(); co = np.zeros(cnt + 1, dtype=np.int64); np.cumsum(clen, out=co[1:])
    comp = torch.empty(int(co[-1]), dtype=torch.uint8, device=dev)
    idx = torch.arange(int(co[-1]), device=dev)
    which = torch.searchsorted(torch.from_numpy(co).to(dev), idx, right=True) - 1
    comp.copy_(d_dst[d_do[which] + (idx - torch.from_numpy(co).to(dev)[which])])
    xo = np.zeros(cnt + 1, dtype=np.int64); np.cumsum(2 * clen + 2, out=xo[1:])
    d_x = torch.empty(int(xo[-1]), dtype=torch.uint8, device=dev); d_co, d_xo = torch.from_numpy(co).to(dev), torch.from_numpy(xo).to(dev)
    g2 = lambda: L.dc_nybble_text_decompress_batch(comp.data_ptr(), d_co.data_ptr(), cnt, modify, d_x.data_ptr(), d_xo.data_ptr(), d_len.data_ptr(), st.data_ptr(), None)
    g2(); torch.cuda.synchronize(); a.record(); g2(); b.record(); b.synchronize()
    ms = a.elapsed_time(b)
    ok = bool((d_len.cpu().numpy() == lens).all())
    print("batch decompress modify", modify, "ms", round(ms, 3), "GB/s of text", round(int(so[-1]) / ms / 1e6, 1), "lengths ok", ok, "status", int(st.item()))
