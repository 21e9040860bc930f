#!/usr/bin/env python
"""Host<->device copy bandwidth on this box (pinned memory, one direction and
both at once).  The end-to-end encode is bounded by these: 69 B/nt in and
256 B/nt out (fp16).  Developer tool; prints and writes gpurun_out/pcie.json."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def main():
    dev = torch.device("cuda:0")
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out = {}

    def run(name, h2d, d2h, chunk=None, iters=5):
        torch.cuda.synchronize()
        best = None
        for _ in range(iters):
            t0 = time.perf_counter()
            step = chunk or n
            for off in range(0, n, step):
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in[off:off + step].copy_(h_in[off:off + step], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out[off:off + step].copy_(d_out[off:off + step], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        gbs = n / best / 1e9
        out[name] = gbs
        print(f"{name}: {gbs:.1f} GB/s per direction ({mb} MiB, best of {iters})", flush=True)

    run("h2d", True, False)
    run("d2h", False, True)
    run("both", True, True)
    run("d2h_chunk128M", False, True, chunk=128 << 20)
    run("d2h_chunk16M", False, True, chunk=16 << 20)
    run("both_chunk128M", True, True, chunk=128 << 20)
    # first-touch cost of a fresh pinned allocation (what encode_graphs pays per call)
    t0 = time.perf_counter()
    fresh = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    out["pin_alloc_s_per_GiB"] = (time.perf_counter() - t0) / (n / 2**30)
    print(f"fresh pinned allocation: {out['pin_alloc_s_per_GiB']:.3f} s per GiB")
    del fresh
    t0 = time.perf_counter()
    fresh = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    out["pin_realloc_s_per_GiB"] = (time.perf_counter() - t0) / (n / 2**30)
    print(f"re-allocation from torch's pinned cache: {out['pin_realloc_s_per_GiB']:.4f} s per GiB")
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "pcie.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
