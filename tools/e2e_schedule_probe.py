import sys, os, time, ctypes as C
sys.path.insert(0, '/root/repo'); sys.argv=['bench.py']
import numpy as np, torch
import bench
from bitar_b200 import _capi as capi, synth
from bitar_b200.engine import CompressDevice, Configuration
from collections import deque
L=capi.lib(); seg=59460; U=1<<30; n=(U+seg-1)//seg; qps=int(os.environ.get('QPS','8')); K=int(os.environ.get('K','2')); STAG=int(os.environ.get('STAG','1'))
data=synth.lineitem_like(U)
dev=CompressDevice(0,qps).Initialize(Configuration(decompressed_seg_size=seg,max_preallocate_memzones=n+64,slot_mem_kind=capi.MEM_PINNED))
h_in,h_out=C.c_void_p(),C.c_void_p()
capi.check(L.bitar_mem_alloc(capi.MEM_PINNED,0,U,64,C.byref(h_in))); capi.check(L.bitar_mem_alloc(capi.MEM_PINNED,0,n*seg,64,C.byref(h_out)))
C.memmove(h_in.value,data.ctypes.data,U); torch.cuda.synchronize()
per=(n+qps-1)//qps; parts=[(q*per,min(n,(q+1)*per)) for q in range(qps) if q*per<n]
def part_ops(a,b): return dev.compress_ops(h_in.value+a*seg, min(U,b*seg)-a*seg)
log=[]
def step():
    t0=time.perf_counter()
    todo=[]
    for a,b in parts:
        m=(b-a+K-1)//K; todo.append(deque((a+k*m,min(b,a+(k+1)*m)) for k in range(K) if a+k*m<b))
    state=['idle']*len(parts); cur=[None]*len(parts); odd=(len(parts)<2) or not STAG
    def sc(q):
        a,b=todo[q].popleft(); o,s=part_ops(a,b); cur[q]=(a,b,s,dev.enqueue('deflate',q,o)); state[q]='c'; log.append((time.perf_counter()-t0,q,'c-start'))
    for q in range(0,len(parts),1 if odd else 2): sc(q)
    done=0
    while done<len(parts):
        for q in range(len(parts)):
            if state[q] in('c','d') and not dev.busy(q):
                a,b,s,r=cur[q]
                if state[q]=='c':
                    log.append((time.perf_counter()-t0,q,'c-done'))
                    cur[q]=(a,b,s,dev.enqueue('inflate',q,dev.decompress_ops(s,r['produced'],h_out.value+a*seg))); state[q]='d'
                    if not odd:
                        odd=True
                        for o in range(1,len(parts),2): sc(o)
                else:
                    log.append((time.perf_counter()-t0,q,'d-done'))
                    dev.put_slots(s)
                    if todo[q]: sc(q)
                    else: state[q]='idle'; done+=1
    return time.perf_counter()-t0
for i in range(4):
    log.clear(); t=step()
print(f"QPS={qps} K={K} STAG={STAG}: step {t*1e3:.1f} ms")
for e in log: print(f"  {e[0]*1e3:7.2f} ms qp{e[1]} {e[2]}")
