#!/usr/bin/env python
"""Reproduce the K5 outliers of round 1 (profiles/r01_d/f_search_timing.json: cosine, first metric
of the last shape, 14.6 s / 41.8 s against 0.06 / 0.35 s): the same call sequence as
tools/search_bench.py, but every call timed on its own, on the device (CUDA events) AND on the
host (perf_counter around the call and around the synchronise), with the SM clock sampled.
Writes gpurun_out/search_repro.json."""
import json
import subprocess
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ginfinity_b200.search import EmbeddingIndex  # noqa: E402


def unit(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(n, 128, generator=g, device="cuda")
    return (a / a.norm(dim=1, keepdim=True)).half()


def smi():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active",
                              "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout
        return out.strip().splitlines()[0]
    except Exception as exc:  # noqa: BLE001
        return str(exc)


def main():
    shapes = [(4096, 1_000_000), (100_000, 1_000_000), (100_000, 12_500_000)]
    rows = []
    for Q, D in shapes:
        t0 = time.perf_counter()
        q, db = unit(Q, 1), unit(D, 2)
        torch.cuda.synchronize()
        gen_s = time.perf_counter() - t0
        index = EmbeddingIndex(db, device="cuda")
        for metric in ("cosine", "l2", "cosine"):
            for call in range(4):
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                h0 = time.perf_counter()
                a.record()
                index.search(q, 10, metric)
                b.record()
                h1 = time.perf_counter()
                torch.cuda.synchronize()
                h2 = time.perf_counter()
                row = {"Q": Q, "D": D, "metric": metric, "call": call, "gen_s": round(gen_s, 3),
                       "device_s": a.elapsed_time(b) * 1e-3, "host_enqueue_s": h1 - h0,
                       "host_total_s": h2 - h0, "smi": smi()}
                print(row, flush=True)
                rows.append(row)
        del index, db
        torch.cuda.empty_cache()
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "search_repro.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
