#!/usr/bin/env python
"""Developer tool: how fast do results leave the device WHILE the encoder's kernels run?

The end-to-end call is bounded by the device->host copy of the embeddings (256 B/nt).  In the
pipeline that copy runs at ~50 GB/s against 56.7 GB/s alone (tools/e2e_timeline.py).  This
tool times the same copy (a) by the copy engine (cudaMemcpyAsync) and (b) by a kernel that
stores straight into the page-locked table through its device address, each alone and with
encode passes of a resident shard running on another stream."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ginfinity_b200.encoder import DeviceShard, Ginfinity  # noqa: E402

device = torch.device("cuda:0")
torch.cuda.set_device(0)
state, _ = bench.load_weights()
shard, _ = bench.build_workload(100_000, seed=0)
enc = Ginfinity.from_state(state, device="cuda:0")
ds = DeviceShard.from_shard(shard, "cuda:0")
out = torch.empty((shard.node_count, 128), dtype=torch.float16, device=device)
load = lambda: enc.encode_device_shard(ds, max_batch_nodes=bench.MAX_BATCH_NODES,  # noqa: E731
                                       max_batch_edges=bench.MAX_BATCH_EDGES, out=out)
for _ in range(2):
    load()
torch.cuda.synchronize()

PIECE = 246 << 20
PIECES = 8
dbuf = torch.empty(PIECE, dtype=torch.uint8, device=device)
host = torch.empty(PIECE * PIECES, dtype=torch.uint8, pin_memory=True)


class Mapped:
    """The page-locked table seen from the device (unified addressing: same pointer)."""
    def __init__(self, t):
        self.__cuda_array_interface__ = {"data": (t.data_ptr(), False), "shape": tuple(t.shape),
                                         "typestr": "|u1", "version": 2}
        self.keep = t


mapped = torch.as_tensor(Mapped(host), device=device)
assert mapped.data_ptr() == host.data_ptr()
s_copy, s_load = torch.cuda.Stream(device), torch.cuda.Stream(device)


def copies(kind):
    for p in range(PIECES):
        if kind == "engine":
            host[p * PIECE:(p + 1) * PIECE].copy_(dbuf, non_blocking=True)
        else:
            mapped[p * PIECE:(p + 1) * PIECE].view(torch.int64).copy_(dbuf.view(torch.int64))


def measure(kind, loaded):
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        if loaded:
            with torch.cuda.stream(s_load):
                for _ in range(4):
                    load()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        with torch.cuda.stream(s_copy):
            a.record()
            copies(kind)
            b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    print("%-7s %-28s %7.2f ms  %.1f GB/s" % (kind, "with encode passes running" if loaded
                                              else "alone", best, PIECE * PIECES / best / 1e6),
          flush=True)


for kind in ("engine", "kernel"):
    for loaded in (False, True):
        measure(kind, loaded)
# the load alone, to know how long 4 passes take (the copies should be shorter than that)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    load()
torch.cuda.synchronize()
print("4 encode passes alone: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
