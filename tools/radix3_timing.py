import sys, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import data_compression_b200 as dc
from data_compression_b200 import synth
dev=torch.device("cuda:0"); n=1<<30
thr,base=synth.zipf_bytes_spec()
data=torch.empty(n,dtype=torch.uint8,device=dev); dc.synth_fill(data,5,synth.device_thresholds(thr,dev),base)
def tm(f,reps=3):
    f(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); f(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for n_ary in (3,4):
    t=dc.huff_build(dc.histogram(data),n_ary); out=torch.empty(n+n//2,dtype=torch.uint8,device=dev)
    res=dc.huff_encode(data,t,out=out); nb=res.bits()
    dec=torch.empty(n,dtype=torch.uint8,device=dev)
    h=t.download()
    print("n",n_ary,"max_bits",h.max_bits,"lut2_used",h.lut2_used,"encode ms",round(tm(lambda: dc.huff_encode(data,t,out=out)),3),"decode ms", round(tm(lambda: dc.huff_decode(res.payload,nb,t,n,out=dec)),3))
    o,st=dc.huff_decode(res.payload,nb,t,n,out=dec); assert int(st.item())==0 and torch.equal(o,data)
    if n_ary == 3:
        ntr = nb // 2
        pay = torch.empty((ntr + 4) // 5 + 16, dtype=torch.uint8, device=dev)
        print("   trit_pack ms", round(tm(lambda: dc.trit_pack(res.payload, ntr, out=pay)), 3), "trit_unpack ms", round(tm(lambda: dc.trit_unpack(pay, ntr)), 3))
# per-kernel breakdown of the decode for every radix (n = 2 and 3 take the ESC instantiations on Zipf data)
import ctypes as C
L = dc.lib()
for n_ary in (2, 3, 4, 16):
    t=dc.huff_build(dc.histogram(data),n_ary); out=torch.empty(n+n//2,dtype=torch.uint8,device=dev)
    res=dc.huff_encode(data,t,out=out); nb=res.bits(); dec=torch.empty(n,dtype=torch.uint8,device=dev)
    dc.huff_decode(res.payload,nb,t,n,out=dec); torch.cuda.synchronize()
    L.dc_profile_reset(); L.dc_profile_enable(1)
    for _ in range(3):
        dc.huff_encode(data,t,out=out); dc.huff_decode(res.payload,nb,t,n,out=dec)
    torch.cuda.synchronize(); L.dc_profile_enable(0)
    row=[]
    for kid in range(40):
        ms, cnt = C.c_double(0), C.c_uint64(0)
        L.dc_profile_kernel(kid, C.byref(ms), C.byref(cnt))
        if cnt.value: row.append(f"{L.dc_profile_kernel_name(kid).decode()}={ms.value/cnt.value:.3f}")
    print("n", n_ary, "max_bits", t.download().max_bits, "bits/sym", round(nb/n,3), " ".join(row))
