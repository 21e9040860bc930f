#!/usr/bin/env python
"""Developer tool: where the time of one device-resident encode pass goes before and between the
chunk launches: perf_counter stamps on the host, CUDA events on the stream."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ginfinity_b200 import _native as nat  # noqa: E402
from ginfinity_b200.encoder import DeviceShard, Ginfinity  # noqa: E402

device = "cuda:0"
torch.cuda.set_device(0)
state, _ = bench.load_weights()
shard, _ = bench.build_workload(100_000, seed=0)
enc = Ginfinity.from_state(state, device=device)
ds = DeviceShard.from_shard(shard, device)
out = torch.empty((shard.node_count, 128), dtype=torch.float16, device=device)
stamps, events = [], []
inner = enc._run_chunk


def stamped(*a, **kw):
    t = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    inner(*a, **kw)
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    stamps.append((t, time.perf_counter()))
    events.append((e0, e1))


enc._run_chunk = stamped
step = lambda: enc.encode_device_shard(ds, max_batch_nodes=bench.MAX_BATCH_NODES,  # noqa: E731
                                       max_batch_edges=bench.MAX_BATCH_EDGES, out=out)
for _ in range(3):
    step()
torch.cuda.synchronize()
for rep in range(3):
    stamps.clear()
    events.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    step()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("pass %d: first chunk enqueued after %.0f us of host time; call returns at %.0f us; device done at %.0f us"
          % (rep, (stamps[0][0] - t0) * 1e6, (t1 - t0) * 1e6, (t2 - t0) * 1e6))
    print("   stream: pass %.3f ms; before the first chunk %.3f ms; chunks %s ms; between chunks %s ms"
          % (a.elapsed_time(b), a.elapsed_time(events[0][0]),
             ["%.3f" % e0.elapsed_time(e1) for e0, e1 in events],
             ["%.3f" % events[i][1].elapsed_time(events[i + 1][0]) for i in range(len(events) - 1)]))
# the stages of one chunk, bracketed by the library's own events
nat.profile_enable(*nat.STAGES)
step()
torch.cuda.synchronize()
print({name: round(nat.profile_read(name)[0], 3) for name in nat.STAGES})
nat.profile_enable()
