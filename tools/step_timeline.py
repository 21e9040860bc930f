#!/usr/bin/env python
"""Developer tool: one device-resident encode pass of the bench shard under torch.profiler
(CUPTI sees the library's own launches too); prints per-kernel totals, the busy time of the
stream and the idle gaps between consecutive kernels.  Numbers under a profiler are for
ATTRIBUTION only, never bench values."""
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ginfinity_b200.encoder import DeviceShard, Ginfinity  # noqa: E402

records = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
device = "cuda:0"
torch.cuda.set_device(0)
state, _ = bench.load_weights()
shard, _ = bench.build_workload(records, seed=0)
enc = Ginfinity.from_state(state, device=device)
ds = DeviceShard.from_shard(shard, device)
out = torch.empty((shard.node_count, 128), dtype=torch.float16, device=device)
step = lambda: enc.encode_device_shard(ds, max_batch_nodes=bench.MAX_BATCH_NODES,  # noqa: E731
                                       max_batch_edges=bench.MAX_BATCH_EDGES, out=out)
for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
step()
b.record()
torch.cuda.synchronize()
print("plain pass %.3f ms for %d nt" % (a.elapsed_time(b), shard.node_count))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
tot = defaultdict(lambda: [0.0, 0])
for e in ev:
    tot[e.name[:60]][0] += e.time_range.end - e.time_range.start
    tot[e.name[:60]][1] += 1
span = ev[-1].time_range.end - ev[0].time_range.start
busy = sum(v[0] for v in tot.values())
print("span %.3f ms, kernel+memop time %.3f ms, %d device events" % (span / 1e3, busy / 1e3, len(ev)))
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%10.1f us %5d  %s" % (us, n, name))
gaps = []
end = ev[0].time_range.end
for e in ev[1:]:
    if e.time_range.start > end:
        gaps.append((e.time_range.start - end, e.name[:50]))
    end = max(end, e.time_range.end)
print("idle between device events: %.1f us in %d gaps" % (sum(g for g, _ in gaps), len(gaps)))
by = defaultdict(lambda: [0.0, 0])
for g, name in gaps:
    by[name][0] += g
    by[name][1] += 1
for name, (us, n) in sorted(by.items(), key=lambda kv: -kv[1][0])[:12]:
    print("   gap before %-50s %8.1f us in %d" % (name, us, n))
