"""In-kernel %globaltimer stamps of the last LM solve of a registration (k_lm_solve_cluster): evaluation, block
reduction + DSMEM push, cluster barrier, controller, per pass."""
import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
import bench
from lmono_b200 import api
_, cm, sm, sweeps = bench.make_workload(0, n_sweeps=4)
dev=torch.device('cuda',0)
st=torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx=api.Context(device=0, stream=st.cuda_stream)
ctx.map_import(0,cm); ctx.map_import(1,sm)
for i in range(6):
    c,s,q,t,qp,tp=sweeps[i%4]
    ctx.map_set_state([0,0,0,1],[0,0,0])
    ctx.map_step(c,s,qp,tp)
out=(C.c_uint64*64)()
ctx.L.lmono_debug_stamps(ctx._h,out,64)
a=np.array(out[:],dtype=np.int64)
print("total solve kernel ns", a[2]-a[0], "arm", a[1]-a[0])
for e in range(5):
    b=a[8+8*e:12+8*e]
    prev = a[1] if e==0 else a[11+8*(e-1)]
    print(e, "eval", b[0]-prev, "reduce", b[1]-b[0], "cluster", b[2]-b[1], "control", b[3]-b[2])
