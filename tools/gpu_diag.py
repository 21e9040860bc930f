#!/usr/bin/env python
"""Developer diagnostics on a B200: per-kernel error patterns and timings.

Prints where the tcgen05 path differs from the SIMT path (by row/column
block) and times every stage kernel with CUDA events.  Not a test, not the
bench; output goes to stdout and gpurun_out/diag.json.
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from ginfinity_b200 import _native as nat  # noqa: E402
from ginfinity_b200.weights import fold, synthetic_state  # noqa: E402

dev = torch.device("cuda:0")
lib = nat.lib
S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
report = {}


def timeit(fn, iters=10, warm=3):
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    state = synthetic_state(seed=7)
    handle = nat.model_create(fold(state))
    g = torch.Generator(device="cpu").manual_seed(0)
    z = (torch.randn(n, 128, generator=g) * 3).to(dev).half()
    h = (torch.randn(n, 128, generator=g) * 2).to(dev).half()
    o_simt = torch.empty_like(h)
    o_umma = torch.empty_like(h)
    small = min(n, 4096 + 77)
    nat.check(lib.gfx_mlp_ln_residual(handle, 0, z.data_ptr(), h.data_ptr(), small,
                                      o_simt.data_ptr(), 0, 1, S()))
    torch.cuda.synchronize()
    print("simt ok", flush=True)
    nat.check(lib.gfx_mlp_ln_residual(handle, 0, z.data_ptr(), h.data_ptr(), small,
                                      o_umma.data_ptr(), 0, 2, S()))
    torch.cuda.synchronize()
    print("umma ran", flush=True)
    d = (o_simt[:small].float() - o_umma[:small].float()).abs().cpu().numpy()
    print("umma vs simt: max", d.max(), "mean", d.mean())
    report["umma_vs_simt_max"] = float(d.max())
    if d.max() > 0.05:
        blk = d[:4096].reshape(32, 128, 4, 32).max(axis=(1, 3))
        print("max err by (tile of 128 rows, 32-col block):\n", np.round(blk[:6], 3))
        rows = d[:128].max(axis=1)
        print("tile0 row err (every 8th):", np.round(rows[::8], 3))
    # head
    for out_code in (0, 1):
        tdt = torch.float16 if out_code == 0 else torch.float32
        a = torch.empty(small, 128, dtype=tdt, device=dev)
        b = torch.empty(small, 128, dtype=tdt, device=dev)
        nat.check(lib.gfx_head_l2norm(handle, h.data_ptr(), None, small, a.data_ptr(), 0,
                                      out_code, 1, S()))
        nat.check(lib.gfx_head_l2norm(handle, h.data_ptr(), None, small, b.data_ptr(), 0,
                                      out_code, 2, S()))
        torch.cuda.synchronize()
        dd = (a.float() - b.float()).abs().max().item()
        print(f"head out_code={out_code}: umma vs simt max {dd}")
        report[f"head_umma_vs_simt_{out_code}"] = dd

    # ---- timings ---------------------------------------------------------
    import ginfinity_b200 as gb
    from helpers import random_records
    recs = random_records(0, 3000)
    shard = gb.GraphBuilder().build_shard(recs)
    reps = max(1, n // shard.node_count)
    # tile the shard `reps` times to reach ~n nodes
    N1, E1 = shard.node_count, shard.edge_count
    ei = np.concatenate([shard.edge_index + np.int32(r * N1) for r in range(reps)], axis=1)
    et = np.tile(shard.edge_types, reps)
    x = np.tile(shard.node_features, (reps, 1))
    N, E = N1 * reps, E1 * reps
    print(f"timing on N={N} E={E} (E/N={E / N:.2f})")
    ei_d, et_d, x_d = (torch.from_numpy(a).to(dev) for a in (ei, et, x))
    row_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    col_src = torch.empty(E, dtype=torch.int32, device=dev)
    col_type = torch.empty(E, dtype=torch.uint8, device=dev)
    need = lib.gfx_csr_workspace_bytes(N, E)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)

    def csr():
        nat.check(lib.gfx_csr_build(ei_d[0].data_ptr(), ei_d[1].data_ptr(), et_d.data_ptr(), N, E,
                                    0, row_ptr.data_ptr(), col_src.data_ptr(),
                                    col_type.data_ptr(), ws.data_ptr(), need, S()))
    t = timeit(csr)
    report["csr_s"] = t
    print(f"csr_build: {t * 1e3:.3f} ms  ({E / t / 1e9:.2f} Gedge/s)")
    for code, name in (((0, "f16"),) if os.environ.get("GFX_DIAG_FAST") else ((0, "f16"), (1, "f32"))):
        tdt = torch.float16 if code == 0 else torch.float32
        hh = torch.randn(N, 128, device=dev).to(tdt)
        zz = torch.empty_like(hh)
        h2 = torch.empty_like(hh)
        t = timeit(lambda: nat.check(lib.gfx_input_linear(handle, x_d.data_ptr(), N, hh.data_ptr(), code, S())))
        print(f"input_linear[{name}]: {t * 1e3:.3f} ms  {N * (28 + 128 * hh.element_size()) / t / 1e9:.0f} GB/s")
        report[f"input_{name}_s"] = t
        t = timeit(lambda: nat.check(lib.gfx_aggregate(handle, 0, hh.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(), N, zz.data_ptr(), code, S())))
        alg = N * (2 * 128 * hh.element_size() + 4) + E * 5
        print(f"aggregate[{name}]: {t * 1e3:.3f} ms  {alg / t / 1e9:.0f} GB/s algorithmic ({alg / N:.0f} B/node)")
        report[f"aggregate_{name}_s"] = t
        report[f"aggregate_{name}_gbs"] = alg / t / 1e9
        impls = ((1, "simt"), (2, "umma"), (5, "umma_lean")) if code == 0 else ((1, "simt"),)
        if os.environ.get("GFX_DIAG_FAST"):
            impls = ((5, "umma_lean"),) if code == 0 else ()
        for impl, iname in impls:
            t = timeit(lambda: nat.check(lib.gfx_mlp_ln_residual(handle, 0, zz.data_ptr(), hh.data_ptr(), N, h2.data_ptr(), code, impl, S())), iters=5, warm=2)
            fl = N * 131072.0
            print(f"mlp[{name},{iname}]: {t * 1e3:.3f} ms  {fl / t / 1e12:.1f} TFLOP/s")
            report[f"mlp_{name}_{iname}_s"] = t
            report[f"mlp_{name}_{iname}_tflops"] = fl / t / 1e12
            t = timeit(lambda: nat.check(lib.gfx_head_l2norm(handle, hh.data_ptr(), None, N, h2.data_ptr(), code, code, impl, S())), iters=5, warm=2)
            print(f"head[{name},{iname}]: {t * 1e3:.3f} ms  {N * 65536.0 / t / 1e12:.1f} TFLOP/s")
            report[f"head_{name}_{iname}_s"] = t
        need_e = lib.gfx_encode_workspace_bytes(N, code)
        wse = torch.empty(need_e, dtype=torch.uint8, device=dev)
        out = torch.empty(N, 128, dtype=tdt, device=dev)
        for impl, iname in impls:
            t = timeit(lambda: nat.check(lib.gfx_encode(handle, x_d.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(), None, N, out.data_ptr(), code, code, impl, 0, wse.data_ptr(), need_e, S())), iters=3, warm=1)
            print(f"encode[{name},{iname}]: {t * 1e3:.3f} ms  {N / t / 1e6:.1f} M nt/s")
            report[f"encode_{name}_{iname}_nts"] = N / t
        if code == 0:
            t = timeit(lambda: nat.check(lib.gfx_layer_fused_pair(handle, 0, hh.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(), N, h2.data_ptr(), S())), iters=5, warm=2)
            print(f"fused_layer[f16]: {t * 1e3:.3f} ms  {N * 131072.0 / t / 1e12:.1f} TFLOP/s  {N / t / 1e9:.2f} G node-layers/s")
            report["fused_layer_s"] = t
            t = timeit(lambda: nat.check(lib.gfx_encode(handle, x_d.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(), None, N, out.data_ptr(), 0, 0, 2, 1, wse.data_ptr(), need_e, S())), iters=3, warm=1)
            print(f"encode[f16,fused]: {t * 1e3:.3f} ms  {N / t / 1e6:.1f} M nt/s")
            report["encode_f16_fused_nts"] = N / t
            t = timeit(lambda: nat.check(lib.gfx_layer_fused_pair(handle, 0, hh.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(), N, h2.data_ptr(), S())), iters=5, warm=2)
            print(f"fused_pair_layer[f16]: {t * 1e3:.3f} ms  {N * 131072.0 / t / 1e12:.1f} TFLOP/s  {N / t / 1e9:.2f} G node-layers/s")
            report["fused_pair_layer_s"] = t
            t = timeit(lambda: nat.check(lib.gfx_encode(handle, x_d.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(), None, N, out.data_ptr(), 0, 0, 5, 2, wse.data_ptr(), need_e, S())), iters=3, warm=1)
            print(f"encode[f16,fused_pair]: {t * 1e3:.3f} ms  {N / t / 1e6:.1f} M nt/s")
            report["encode_f16_fused_pair_nts"] = N / t
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "diag.json").write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
