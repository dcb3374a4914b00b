import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from oxli_b200 import _capi as capi
n,L,k,G=10_000_000,150,31,5_000_000
tb=n*L
d_bases=capi.device_alloc(tb+64); d_offs=capi.device_alloc((n+1)*8)
capi.synth_reads_device(d_bases,n,L,G,0xC20001)
capi.h2d(d_offs, np.arange(n+1,dtype=np.uint64)*np.uint64(L))
nw=tb-k+1
d_h=capi.device_alloc(nw*8)
t=capi.Table(k,capacity_hint=G)
for it in range(3):
    t.timer_start(); t.hash_batch_device(d_bases,d_offs,n,tb,d_h); ms=t.timer_stop()
    print('K1 hash-only: %.2f ms  %.1f G windows/s'%(ms, nw/ms/1e6))
for it in range(3):
    t.clear()
    t.timer_start(); c=t.count_hashes_device(d_h,nw,True); ms=t.timer_stop()
    print('K2 count list: %.2f ms  %.1f G kmers/s (counted %d) kernel_ms=%s'%(ms, c/ms/1e6, c, t.last_consume_kernel_ms()))
print(len(t), t.capacity)
t2=capi.Table(k,capacity_hint=G); st,tot,_,_=t2.consume_batch_device(d_bases,d_offs,n,tb,True)
print(tot, t2.digest()==t.digest())
