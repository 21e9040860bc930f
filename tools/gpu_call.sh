#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
import sys, os, torch, numpy as np
sys.path.insert(0, '.')
from ginfinity_b200 import _native as nat
from ginfinity_b200.encoder import DeviceShard, Ginfinity
from ginfinity_b200.synthetic import synthetic_shard
from bench import load_weights
state, label = load_weights()
shard = synthetic_shard(0, 20000)
ds = DeviceShard.from_shard(shard, 'cuda:0')
enc = Ginfinity.from_state(state, device='cuda:0', full_precision=True)
out = torch.empty((shard.node_count, 128), dtype=torch.float32, device='cuda:0')
go = lambda: enc.encode_device_shard(ds, out_dtype=nat.GFX_F32, out=out)
go(); torch.cuda.synchronize()
nat.profile_enable(*nat.STAGES)
for s in nat.STAGES: nat.profile_read(s)
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record(); go(); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print('fp32 split + pipelined K1: %.1f M nt/s, %.2f ms for %d nt' % (shard.node_count / ms / 1e3, ms, shard.node_count))
print({s: round(nat.profile_read(s)[0], 3) for s in nat.STAGES})
ref = out.clone()
PY
GFX_K1_F32_GENERIC=1 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from ginfinity_b200 import _native as nat
from ginfinity_b200.encoder import DeviceShard, Ginfinity
from ginfinity_b200.synthetic import synthetic_shard
from bench import load_weights
state, label = load_weights()
shard = synthetic_shard(0, 20000)
ds = DeviceShard.from_shard(shard, 'cuda:0')
enc = Ginfinity.from_state(state, device='cuda:0', full_precision=True)
out = torch.empty((shard.node_count, 128), dtype=torch.float32, device='cuda:0')
go = lambda: enc.encode_device_shard(ds, out_dtype=nat.GFX_F32, out=out)
go(); torch.cuda.synchronize()
nat.profile_enable(*nat.STAGES)
for s in nat.STAGES: nat.profile_read(s)
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record(); go(); b.record(); torch.cuda.synchronize()
print('fp32 split + generic K1: %.1f M nt/s' % (shard.node_count / a.elapsed_time(b) / 1e3), {s: round(nat.profile_read(s)[0], 3) for s in ('aggregate','mlp','head')})
PY
