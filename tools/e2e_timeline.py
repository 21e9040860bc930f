#!/usr/bin/env python
"""Developer tool: where one end-to-end `encode_graphs` call (page-locked shard in, host
table out) spends its time against the raw copies it is bounded by.

  1. wall clock of the call (three runs);
  2. the same call under torch.profiler: busy time and idle gaps of the device->host copy
     engine (the bottleneck: 256 B/nt out against 69 B/nt in), first copy start, last copy end,
     the device events before the first large copy, and per copy what ran beside it;
  3. raw device->host copies of the same bytes with no kernels, into (a) ONE reused 256 MiB
     page-locked buffer (what bench.py's copy ceiling does) and (b) the caller's whole
     [nodes, 128] table in the pipeline's piece size, alone and with the host->device
     copies of the step running beside them.

Numbers under the profiler are for ATTRIBUTION only, never bench values."""
import json
import sys
import time
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ginfinity_b200.encoder import Ginfinity, pin_shard  # noqa: E402

records = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
device = torch.device("cuda:0")
torch.cuda.set_device(0)
state, _ = bench.load_weights()
shard, _ = bench.build_workload(records, seed=0)
enc = Ginfinity.from_state(state, device="cuda:0")
pinned = pin_shard(shard)
nodes = shard.node_count
table = Ginfinity.pinned_table(nodes)
run = lambda: enc.encode_graphs(pinned, max_batch_nodes=bench.MAX_BATCH_NODES,  # noqa: E731
                                max_batch_edges=bench.MAX_BATCH_EDGES, out=table)
report = {"nodes": nodes}
for _ in range(2):
    run()
torch.cuda.synchronize()
times = []
for _ in range(3):
    t0 = time.perf_counter()
    res = run()
    times.append(time.perf_counter() - t0)
    del res
report["call_ms"] = [round(t * 1e3, 2) for t in times]
print("encode_graphs(out=table): %s ms" % report["call_ms"], flush=True)

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    prof_wall = time.perf_counter() - t0
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
first = ev[0].time_range.start
d2h = [e for e in ev if "Memcpy DtoH" in e.name and e.time_range.end - e.time_range.start > 200]
h2d = [e for e in ev if "Memcpy HtoD" in e.name]
busy = sum(e.time_range.end - e.time_range.start for e in d2h)
gaps = [b.time_range.start - a.time_range.end for a, b in zip(d2h[:-1], d2h[1:])]
report["profiled"] = {
    "wall_ms": prof_wall * 1e3,
    "device_span_ms": (ev[-1].time_range.end - first) / 1e3,
    "d2h_copies": len(d2h), "d2h_busy_ms": busy / 1e3,
    "d2h_first_start_ms": (d2h[0].time_range.start - first) / 1e3,
    "d2h_last_end_ms": (d2h[-1].time_range.end - first) / 1e3,
    "d2h_gap_ms_total": sum(g for g in gaps if g > 0) / 1e3,
    "d2h_gap_ms_max": max(gaps) / 1e3,
    "d2h_gb_per_s_while_busy": nodes * 256 / (busy * 1e-6) / 1e9,
    "h2d_copies": len(h2d),
    "h2d_busy_ms": sum(e.time_range.end - e.time_range.start for e in h2d) / 1e3,
}
tot = defaultdict(float)
for e in ev:
    if "Memcpy" not in e.name:
        tot["kernels"] += e.time_range.end - e.time_range.start
report["profiled"]["kernel_busy_ms"] = tot["kernels"] / 1e3
print(json.dumps(report["profiled"], indent=1), flush=True)
print("device events up to the first large device->host copy (ms from the first event):")
for e in ev:
    if e.time_range.start > d2h[0].time_range.start:
        break
    print("  %8.3f .. %8.3f  %s" % ((e.time_range.start - first) / 1e3,
                                   (e.time_range.end - first) / 1e3, e.name[:70]))


def overlap(e, others):
    a, b = e.time_range.start, e.time_range.end
    return sum(max(0, min(b, o.time_range.end) - max(a, o.time_range.start)) for o in others)


kern = [e for e in ev if "Memcpy" not in e.name and "Memset" not in e.name]
print("per device->host copy: ms, GB/s (bytes from the duration-weighted share), "
      "ms of host->device copies beside it, ms of kernels beside it")
total = sum(e.time_range.end - e.time_range.start for e in d2h)
for e in d2h:
    dur = e.time_range.end - e.time_range.start
    print("  %7.3f ms   h2d beside %6.3f   kernels beside %6.3f" %
          (dur / 1e3, overlap(e, h2d) / 1e3, overlap(e, kern) / 1e3))
print("h2d copies: " + " ".join("%.2f" % ((e.time_range.end - e.time_range.start) / 1e3)
                                for e in h2d))

# ---------------- raw copies, no kernels ----------------
d2h_bytes = nodes * 256
h2d_bytes = sum(getattr(shard, n).nbytes for n in
                ("node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr"))
piece = 960_000 * 256
dbuf = torch.empty(256 << 20, dtype=torch.uint8, device=device)
small = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
h_in = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(256 << 20, dtype=torch.uint8, device=device)
big = torch.from_numpy(table).view(torch.uint8).reshape(-1)
s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)


big_in = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)   # like the pinned shard
d_big_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=device)


def raw(name, into_table, with_h2d, step, h2d_step=None):
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if with_h2d:
            with torch.cuda.stream(s1):
                left = h2d_bytes
                while left > 0 and h2d_step is None:
                    n = min(left, h_in.numel())
                    d_in[:n].copy_(h_in[:n], non_blocking=True)
                    left -= n
                off = 0
                while off < h2d_bytes and h2d_step is not None:
                    n = min(h2d_step, h2d_bytes - off)
                    d_big_in[off:off + n].copy_(big_in[off:off + n], non_blocking=True)
                    off += n
        with torch.cuda.stream(s2):
            off = 0
            while off < d2h_bytes:
                n = min(step, d2h_bytes - off, dbuf.numel())
                dst = big[off:off + n] if into_table else small[:n]
                dst.copy_(dbuf[:n], non_blocking=True)
                off += n
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    report[name] = {"ms": best * 1e3, "d2h_gb_per_s": d2h_bytes / best / 1e9}
    print("%-44s %7.2f ms  (%.1f GB/s device->host)" % (name, best * 1e3, d2h_bytes / best / 1e9),
          flush=True)


raw("d2h alone, reused 256 MiB buffer", False, False, 256 << 20)
raw("d2h alone, whole table, 246 MB pieces", True, False, piece)
raw("d2h alone, whole table, 61 MB pieces", True, False, piece // 4)
raw("d2h + h2d, reused 256 MiB buffer (ceiling)", False, True, 256 << 20)
raw("d2h + h2d, whole table, 246 MB pieces", True, True, piece)
raw("d2h + h2d of the whole shard in 256 MiB pieces", True, True, piece, 256 << 20)
raw("d2h + h2d of the whole shard in 16 MiB pieces", True, True, piece, 16 << 20)
raw("d2h + h2d of the whole shard in one piece", True, True, piece, 1 << 40)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "e2e_timeline.json").write_text(json.dumps(report, indent=1))
