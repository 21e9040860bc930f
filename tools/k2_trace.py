#!/usr/bin/env python
"""Developer timeline of K2 (umma6_mlp_kernel, or with argument 7 the CTA-pair kernel): CTA 0 stamps
clock64() at the hand-over points."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ginfinity_b200 import _native as nat  # noqa: E402
from ginfinity_b200.weights import fold, synthetic_state  # noqa: E402

dev = torch.device("cuda:0")
lib = nat.lib
S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
handle = nat.model_create(fold(synthetic_state(seed=7)))
n = 1 << 20
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 6
z = torch.randn(n, 128, device=dev).half()
h = torch.randn(n, 128, device=dev).half()
o = torch.empty_like(h)
run = lambda: nat.check(lib.gfx_mlp_ln_residual(handle, 0, z.data_ptr(), h.data_ptr(), n,  # noqa: E731
                                                o.data_ptr(), 0, impl, S()))
for _ in range(3):
    run()
torch.cuda.synchronize()
trace = torch.zeros(64 * 16 + 2 * 160, dtype=torch.int64, device=dev)
raw = ctypes.CDLL(str(ROOT / "ginfinity_b200" / "libgfx.so"))
raw.gfx_debug_k2_trace.argtypes = [ctypes.c_void_p]
raw.gfx_debug_k2_trace(trace.data_ptr())
run()
torch.cuda.synchronize()
raw.gfx_debug_k2_trace(None)
full = trace.cpu().numpy()
t = full[:1024].reshape(64, 16)
spans = full[1024:].reshape(-1, 2)
t0 = t[t > 0].min()
names = {0: "zIssue", 2: "A1full", 3: "mma1iss", 1: "resIssue", 4: "A2aFull", 5: "D2empty", 8: "mma2iss",
         6: "epiAa", 7: "epiAb", 11: "epiB_S", 9: "epiB_E", 10: "stored"}
order = [0, 2, 3, 1, 6, 4, 5, 7, 8, 11, 9, 10]
print("cycles since first stamp (CTA 0)")
print("it   " + " ".join(f"{names[k]:>8}" for k in order))
for it in range(24):
    if not t[it].any():
        break
    print(f"{it:3d}  " + " ".join(f"{(t[it, k] - t0) if t[it, k] else -1:8d}" for k in order))
rows = [i for i in range(64) if t[i, 14] and t[i, 2]]
if len(rows) > 8:
    a, b = rows[4], rows[-1]
    cyc, ns = t[b, 2] - t[a, 2], t[b, 14] - t[a, 14]
    print(f"iterations {a}..{b}: {cyc} cycles in {ns} ns = {cyc / ns:.3f} GHz effective SM clock; "
          f"{cyc / (b - a):.0f} cycles = {ns / (b - a):.0f} ns per iteration")
if impl == 7 and spans.any():
    import numpy as np
    live = spans[spans[:, 0] > 0]
    s0 = live[:, 0].min()
    start, end = live[:, 0] - s0, live[:, 1] - s0
    print(f"{len(live)} CTAs: start {start.min()}..{start.max()} ns, end {end.min()}..{end.max()} ns, "
          f"duration median {np.median(end - start):.0f} ns (min {np.min(end - start)}, max {np.max(end - start)})")
    print("CTA 0: start", start[0], "end", end[0])
    d = (end - start) // 1000
    print("duration (us) by CTA:", " ".join(str(int(x)) for x in d))
