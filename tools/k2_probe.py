#!/usr/bin/env python
"""Developer probe: K1 and K2 per-node time against the working-set size (is the kernel bound by
HBM or by its own pipeline?  below ~30 MB per buffer everything stays in the 126 MB L2)."""
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


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


impls = [int(a) for a in sys.argv[1:]] or [5]
for impl, n in ((i, n) for i in impls for n in (1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21)):
    z = torch.randn(n, 128, device=dev).half()
    h = torch.randn(n, 128, device=dev).half()
    o = torch.empty_like(h)
    t = timeit(lambda: nat.check(lib.gfx_mlp_ln_residual(handle, 0, z.data_ptr(), h.data_ptr(), n,
                                                         o.data_ptr(), 0, impl, S())))
    print(f"K2[{impl}] n={n:8d}  {t * 1e6:8.1f} us  {t / n * 1e9:.4f} ns/node  {n * 131072 / t / 1e12:7.1f} TFLOP/s"
          f"  {n * 768 / t / 1e12:.2f} TB/s   working set {3 * n * 256 / 1e6:.0f} MB")
