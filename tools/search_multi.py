#!/usr/bin/env python
"""Sharded similarity search under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port 29511 tools/search_multi.py

1. correctness: the merged result of N row shards equals a single-GPU search
   over the whole database (indices and scores bit-identical);
2. timing of BASELINE configs[4] scaled to N GPUs: Q = 1e5 queries against
   1.25e7 rows per GPU, k = 10 -- local scan, all-gather, merge, each timed
   with CUDA events, max over ranks.  Rank 0 prints one JSON line and writes
   gpurun_out/search_multi_N.json."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ginfinity_b200 import search as S  # noqa: E402


def unit(n, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    a = torch.randn(n, 128, generator=g, device=device)
    return (a / a.norm(dim=1, keepdim=True)).half()


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in
                          (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("GFX_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    rows_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
    Q, k = 100_000, 10
    report = {"n_gpus": world}

    # ---- 1. correctness --------------------------------------------------------
    full = unit(300_000 * world + 17, 5, dev)           # identical on every rank (same seed)
    q = unit(3000, 6, dev)
    lo, hi = S.shard_bounds(full.shape[0], world, rank)
    index = S.EmbeddingIndex(full[lo:hi], device=dev, index_base=lo, total_rows=full.shape[0])
    for metric in ("cosine", "l2"):
        got_s, got_i = index.search_sharded(q, k, metric)
        want_s, want_i = S.EmbeddingIndex(full, device=dev).search(q, k, metric)
        ok = bool(torch.equal(got_i, want_i) and torch.equal(got_s, want_s))
        flag = torch.tensor([1 if ok else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        report[f"sharded_equals_single_{metric}"] = bool(flag.item())
    del full, index

    # ---- 2. timing -------------------------------------------------------------
    db = unit(rows_per_gpu, 100 + rank, dev)
    q = unit(Q, 7, dev)
    index = S.EmbeddingIndex(db, device=dev, index_base=rank * rows_per_gpu,
                             total_rows=world * rows_per_gpu)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for it in range(2):                                   # first pass warms up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        s, i = index.search(q, k, "cosine")
        ev[1].record()
        if world > 1:
            all_s, all_i = S.gather_lists(s, i, world)
        else:
            all_s, all_i = s[None], i[None]
        ev[2].record()
        out_s, out_i = S.merge_lists(all_s, all_i)
        ev[3].record()
        torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]),
                      ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3])], device=dev,
                     dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    scan, gather, merge, total = (v * 1e-3 for v in t.tolist())
    flop = 2.0 * 128 * Q * rows_per_gpu * world
    report.update({
        "queries": Q, "rows_per_gpu": rows_per_gpu, "k": k,
        "scan_s": scan, "allgather_s": gather, "merge_s": merge, "total_s": total,
        "allgather_bytes_per_rank": Q * k * 12,
        "tflops_whole_job": flop / total / 1e12,
        "db_rows_per_s_whole_job": rows_per_gpu * world / total,
        "query_db_pairs_per_s": Q * rows_per_gpu * world / total,
    })
    if rank == 0:
        print(json.dumps(report), flush=True)
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"search_multi_{world}.json").write_text(json.dumps(report, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
