#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
import bench
from ginfinity_b200.encoder import DeviceShard, Ginfinity
state,_=bench.load_weights()
shard,_=bench.build_workload(20000, seed=0)
for lib in ('prev','new'):
    pass
enc=Ginfinity.from_state(state, device='cuda:0', full_precision=True)
ds=DeviceShard.from_shard(shard,'cuda:0')
from ginfinity_b200 import _native as nat
out=torch.empty((shard.node_count,128),dtype=torch.float32,device='cuda:0')
go=lambda: enc.encode_device_shard(ds,max_batch_nodes=bench.MAX_BATCH_NODES,max_batch_edges=bench.MAX_BATCH_EDGES,out_dtype=nat.GFX_F32,out=out)
for _ in range(2): go()
torch.cuda.synchronize()
a,b=torch.cuda.Event(True),torch.cuda.Event(True)
a.record()
for _ in range(3): go()
b.record(); torch.cuda.synchronize()
print('fp32 path: %.3e nt/s'%(shard.node_count*3/(a.elapsed_time(b)*1e-3)))
PY
