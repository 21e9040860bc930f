#!/usr/bin/env python
"""Developer tool: where the host-side time of one device-resident encode pass goes before and
between the chunk launches (perf_counter stamps; the device is idle until the first chunk)."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ginfinity_b200.encoder import DeviceShard, Ginfinity  # noqa: E402

device = "cuda:0"
torch.cuda.set_device(0)
state, _ = bench.load_weights()
shard, _ = bench.build_workload(100_000, seed=0)
enc = Ginfinity.from_state(state, device=device)
ds = DeviceShard.from_shard(shard, device)
out = torch.empty((shard.node_count, 128), dtype=torch.float16, device=device)
stamps = []
inner = enc._run_chunk


def stamped(*a, **kw):
    t = time.perf_counter()
    inner(*a, **kw)
    stamps.append((t, time.perf_counter()))


enc._run_chunk = stamped
step = lambda: enc.encode_device_shard(ds, max_batch_nodes=bench.MAX_BATCH_NODES,  # noqa: E731
                                       max_batch_edges=bench.MAX_BATCH_EDGES, out=out)
for _ in range(3):
    step()
torch.cuda.synchronize()
for rep in range(3):
    stamps.clear()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("pass %d: first chunk enqueued after %.0f us; chunks take %s us of host time; call returns at %.0f us; "
          "device done at %.0f us" % (rep, (stamps[0][0] - t0) * 1e6,
                                      [int((b - a) * 1e6) for a, b in stamps], (t1 - t0) * 1e6, (t2 - t0) * 1e6))
