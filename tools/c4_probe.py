#!/usr/bin/env python
"""Developer probe of the shard-file pipeline (BASELINE configs[3]): where the time of
encode_shard_files goes -- file read + pin (ShardPrefetcher alone), encode of pinned shards alone,
and both together."""
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import ginfinity_b200 as g  # noqa: E402
from bench import load_weights  # noqa: E402
from ginfinity_b200.multi_gpu import ShardPrefetcher, encode_shard_files  # noqa: E402
from ginfinity_b200.synthetic import synthetic_shard  # noqa: E402

files, records = int(sys.argv[1]) if len(sys.argv) > 1 else 10, 10_000
root = Path(tempfile.mkdtemp(prefix="gfx_c4_", dir="/dev/shm"))
paths, nodes = [], 0
for k in range(files):
    shard = synthetic_shard(1000 + k, records, prefix=f"f{k}_")
    g.save_graph_shard(shard, root / f"s{k}.safetensors")
    paths.append(root / f"s{k}.safetensors")
    nodes += shard.node_count
state, _ = load_weights()
enc = g.Ginfinity.from_state(state, device="cuda:0")
enc.encode_graphs(shard)
torch.cuda.synchronize()
for workers in (1, 4, 8):
    t0 = time.perf_counter()
    kept = 0
    for path, s in ShardPrefetcher(paths, device_index=0, workers=workers):
        kept += s.node_count
    dt = time.perf_counter() - t0
    print(f"prefetch only, {workers} workers: {dt:.3f} s = {nodes / dt / 1e6:.1f} M nt/s", flush=True)
pinned = [g.load_graph_shard(p) for p in paths[:3]]
from ginfinity_b200.encoder import pin_shard  # noqa: E402
pinned = [pin_shard(s) for s in pinned]
t0 = time.perf_counter()
for s in pinned:
    enc.encode_graphs(s)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"encode only (3 pinned shards): {dt:.3f} s = {sum(s.node_count for s in pinned) / dt / 1e6:.1f} M nt/s", flush=True)
for label, kw in (("kept", {}), ("consumed", {"consume": lambda path, arrays: None}),
                  ("consumed again", {"consume": lambda path, arrays: None})):
    t0 = time.perf_counter()
    out = encode_shard_files(enc, paths, rank=0, world_size=1, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"encode_shard_files ({label}): {dt:.3f} s = {nodes / dt / 1e6:.1f} M nt/s", flush=True)
import shutil
shutil.rmtree(root, ignore_errors=True)
