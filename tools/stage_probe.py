#!/usr/bin/env python
"""Developer tool: per-stage milliseconds of one device-resident encode pass of the bench
shard (CUDA events recorded inside libgfx, gfx_profile_enable).  With GFX_LIBRARY=<other build>
the same script times another build on the same board (A/B).
    python tools/stage_probe.py [records] [fp32]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ginfinity_b200 import _native as nat  # noqa: E402
from ginfinity_b200.encoder import DeviceShard, Ginfinity  # noqa: E402

records = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
fp32 = "fp32" in sys.argv[2:]
state, _ = bench.load_weights()
shard, _ = bench.build_workload(records, seed=0)
enc = Ginfinity.from_state(state, device="cuda:0", full_precision=fp32)
ds = DeviceShard.from_shard(shard, "cuda:0")
out = torch.empty((shard.node_count, 128), dtype=torch.float32 if fp32 else torch.float16,
                  device="cuda:0")
kw = dict(max_batch_nodes=bench.MAX_BATCH_NODES, max_batch_edges=bench.MAX_BATCH_EDGES, out=out)
if fp32:
    kw["out_dtype"] = nat.GFX_F32
go = lambda: enc.encode_device_shard(ds, **kw)  # noqa: E731
for _ in range(3):
    go()
torch.cuda.synchronize()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(5):
    go()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print("%s: %.3e nt/s, %.3f ms per pass" % ("fp32" if fp32 else "fp16", shard.node_count / ms * 1e3, ms))
nat.profile_enable(*nat.STAGES)
go()
torch.cuda.synchronize()
stages = {name: round(nat.profile_read(name)[0], 3) for name in nat.STAGES}
print({k: v for k, v in stages.items() if v > 0})
