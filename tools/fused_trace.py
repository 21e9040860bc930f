#!/usr/bin/env python
"""Developer timeline of the fused pair kernel: CTA 0 stamps clock64() at the hand-over points
of every tile iteration (gfx_debug_fused_trace); this prints them relative to the first stamp."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from ginfinity_b200 import _native as nat  # noqa: E402
from ginfinity_b200.weights import fold, synthetic_state  # noqa: E402
import ginfinity_b200 as gb  # noqa: E402
from helpers import random_records  # noqa: E402

dev = torch.device("cuda:0")
lib = nat.lib
S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
handle = nat.model_create(fold(synthetic_state(seed=7)))
shard = gb.GraphBuilder().build_shard(random_records(0, 3000))
N, E = shard.node_count, shard.edge_count
ei, et = torch.from_numpy(shard.edge_index).to(dev), torch.from_numpy(shard.edge_types).to(dev)
row_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
col_src = torch.empty(E, dtype=torch.int32, device=dev)
col_type = torch.empty(E, dtype=torch.uint8, device=dev)
need = lib.gfx_csr_workspace_bytes(N, E)
ws = torch.empty(need, dtype=torch.uint8, device=dev)
nat.check(lib.gfx_csr_build(ei[0].data_ptr(), ei[1].data_ptr(), et.data_ptr(), N, E, 0,
                            row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(),
                            ws.data_ptr(), need, S()))
h = torch.randn(N, 128, device=dev).half()
out = torch.empty_like(h)
banded = len(sys.argv) > 1 and sys.argv[1] == "banded"
desc = torch.empty(N, dtype=torch.int32, device=dev)
nat.check(lib.gfx_row_describe(row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(), N,
                               desc.data_ptr(), S()))


def run():
    if banded:
        nat.check(lib.gfx_layer_fused_banded(handle, 0, h.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(),
                                             col_type.data_ptr(), desc.data_ptr(), N, out.data_ptr(), S()))
    else:
        nat.check(lib.gfx_layer_fused_pair(handle, 0, h.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(),
                                           col_type.data_ptr(), N, out.data_ptr(), S()))



for _ in range(3):
    run()
torch.cuda.synchronize()
trace = torch.zeros(64 * 16 + 2 * 160, dtype=torch.int64, device=dev)
raw = ctypes.CDLL(str(ROOT / "ginfinity_b200" / "libgfx.so"))
raw.gfx_debug_fused_trace.argtypes = [ctypes.c_void_p]
raw.gfx_debug_fused_trace(trace.data_ptr())
run()
torch.cuda.synchronize()
raw.gfx_debug_fused_trace(None)
full = trace.cpu().numpy()
t = full[:1024].reshape(64, 16)
spans = full[1024:].reshape(-1, 2)
t0 = t[t > 0].min()
names = ["prodS", "prodE", "A1full", "mma1", "A2+D2e", "mma2", "epiA_S", "epiA_E", "epiB_S", "epiB_E",
         "stO", "stE", "load", "w15S", "w15E", "w12E"]
print("cycles since first stamp (CTA 0); tile iterations down")
print("it   " + " ".join(f"{n:>7}" for n in names))
for it in range(40):
    if not t[it].any():
        break
    print(f"{it:3d}  " + " ".join(f"{(t[it, k] - t0) if t[it, k] else -1:7d}" for k in range(16)))
live = spans[spans[:, 0] > 0]
if len(live):
    s0 = live[:, 0].min()
    start, end = live[:, 0] - s0, live[:, 1] - s0
    print(f"{len(live)} CTAs: start {start.min()}..{start.max()} ns, end {end.min()}..{end.max()} ns, "
          f"duration median {np.median(end - start):.0f} ns (min {np.min(end - start)}, max {np.max(end - start)})")
    print("duration (us) by CTA:", " ".join(str(int(x)) for x in (end - start) // 1000))
