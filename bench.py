#!/usr/bin/env python
"""Headline benchmark: nucleotides/s embedded (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): a synthetic shard of 100,000 RNAs x ~200
nt per GPU, `encode_graphs` with the fp16 model at max_batch_nodes=60,000 /
max_batch_edges=300,000.  One step = one pass of the hot path over the whole
shard: on-device microbatch packing, destination-CSR build, row descriptors,
input projection, 4 x fused GINE layer, head + L2 normalise.

  value     nt/s with the shard's arrays already resident in HBM (device
            timed with CUDA events; inputs 1.4 GB > L2 so no flush needed)
  e2e       the same through the public `Ginfinity.encode_graphs` call with
            HOST buffers: pinned host -> device copies of every input array
            and device -> host copy of every embedding inside the timed region;
            `e2e.copy_ceiling` is the same bytes moved with no kernels at all
            (all ranks at once), `e2e_retained` the same call when the caller
            KEEPS every step's result (caller-owned tables, `out=`)
  roofline  the dominant kernel of the step against the measured B200 peak
  parity_check  >= 200 records of the timed workload against the REAL reference
            (oracle/_ref) on the host CPU
  cpu_baseline / --impl reference
            the unmodified reference (`ginfinity.Ginfinity.load(device="cpu")`
            from oracle/_ref; `kind: "reference"`) on this box's host cores, on
            a bounded sample of the same shard; the torch-CPU port
            (oracle/cpu_port.py, `kind: "port"`) only when nothing is staged
  reference_cuda_eager  the reference's own eager `device="cuda"` path on this
            B200 (cuBLAS / ATen library kernels; SURVEY 8d's second bar)
  search    K5 at the C5 shape per GPU (1e5 queries x 1.25e7 rows, k = 10) with
            the NCCL all-gather + merge at N > 1
  c1 / c3 / c4  the other BASELINE configs as extra keys
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "nucleotides/sec embedded"
UNIT = "nt/s"
MAX_BATCH_NODES, MAX_BATCH_EDGES = 60_000, 300_000
FLOP_PER_NODE_MLP = 2 * 128 * 256 * 2            # K2, per layer  
BYTES_PER_NODE_AGG = 539.0                        # K1 fp16, per layer
BYTES_PER_NODE_FUSED = 539.0                      # fused layer: h in 256, h' out 256, CSR 26.7
BYTES_PER_NODE_BANDED = 516.0                     # banded fused layer: h in 256, h' out 256, row descriptor 4
SEARCH_Q, SEARCH_ROWS_PER_GPU, SEARCH_K = 100_000, 12_500_000, 10


def measured_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.is_file():
        p = json.loads(path.read_text())
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json; sustained bf16 figure: "
                           "kernel timed inside a long step)")
    return dict(hbm_gbs=6650.0, tflops=1400.0,
                source="fallback (B200_PROFILING.md: 6.65 TB/s, ~1.4 PFLOP/s sustained)")


class ClockSampler:
    """One `nvidia-smi -lms 50` process for the whole timed span (device-timed
    steps and end-to-end steps): SM clock and throttle reasons every 50 ms."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)                      # let the first samples arrive
        except OSError:
            self.proc = None

    def finish(self) -> dict:
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                out, _ = self.proc.communicate()
            for line in out.splitlines():
                cells = [c.strip() for c in line.split(",")]
                if len(cells) == 6:
                    try:
                        float(cells[0])
                    except ValueError:
                        continue
                    rows.append(cells)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names)
                   if any(r[2 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(rows[0][1]),
                "reasons": reasons, "samples": len(sm)}


def ncu_traffic(stage: str, nodes_per_launch: float):
    """DRAM bytes per launch of a stage's kernel, from the committed ncu
    capture (profiles/ncu_traffic.json: bytes per node of that capture, scaled
    to this run's nodes per launch); None when no capture is recorded."""
    path = ROOT / "profiles" / "ncu_traffic.json"
    if not path.is_file():
        return None
    entry = json.loads(path.read_text()).get(stage)
    return None if entry is None else entry["bytes_per_node"] * nodes_per_launch


def build_workload(records: int, seed: int):
    from ginfinity_b200.synthetic import synthetic_shard
    t0 = time.time()
    shard = synthetic_shard(seed, records)
    return shard, time.time() - t0


def load_weights():
    """(state, label): the reference checkpoint when staged, else seeded
    random weights of the same architecture."""
    from ginfinity_b200.weights import default_model_dir, load_checkpoint, synthetic_state
    root = default_model_dir()
    if root is not None:
        return load_checkpoint(root)[0], "reference checkpoint (sha256-verified)"
    return synthetic_state(seed=0), "random-init weights of the bundled architecture"


# --------------------------------------------------------------------------------------------
# the reference on the host CPU (real package from oracle/_ref, else the torch-CPU port)
# --------------------------------------------------------------------------------------------
class CpuReference:
    """`encode(shard)` = the reference's `encode_graphs` at the bench limits on the host CPU."""

    def __init__(self, state, full_precision: bool = False):
        from oracle import ref_loader
        self.kind, self.ref = "port", None
        if ref_loader.reference_root() is not None:
            self.ref, _api = ref_loader.import_reference()
            self.encoder = self.ref.Ginfinity.load(device="cpu", full_precision=full_precision)
            self.kind = "reference"
            self.what = (f"unmodified ginfinity {self.ref.__version__} from oracle/_ref, "
                         f"Ginfinity.load(device='cpu'{', full_precision=True' if full_precision else ''})"
                         ".encode_graphs")
        else:
            from oracle.cpu_port import CpuPort
            self.encoder = CpuPort(state, full_precision=full_precision)
            self.what = "oracle/cpu_port.py (the reference's torch ops in the same order; nothing staged)"

    def shard(self, shard):
        if self.ref is None:
            return shard
        from oracle import ref_loader
        return ref_loader.to_reference_shard(self.ref, shard)

    def encode(self, shard, dtype=np.float16):
        return self.encoder.encode_graphs(shard, max_batch_nodes=MAX_BATCH_NODES,
                                          max_batch_edges=MAX_BATCH_EDGES, embedding_dtype=dtype)


def cpu_reference_rate(state, shard, sample_records: int):
    """nt/s of the reference CPU path on the first `sample_records` records of the workload,
    all host threads; also returns its outputs (the parity check reuses them)."""
    import torch
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    sub = shard.slice(0, min(sample_records, shard.record_count))
    cpu = CpuReference(state)
    ref_shard = cpu.shard(sub)
    t0 = time.perf_counter()
    outputs = cpu.encode(ref_shard)
    secs = time.perf_counter() - t0
    return dict(rate=sub.node_count / secs, sub=sub, threads=torch.get_num_threads(), secs=secs,
                kind=cpu.kind, what=cpu.what, outputs=outputs)


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is given
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    state, label = load_weights()
    shard, _ = build_workload(args.sample_records, seed=0)
    cpu = CpuReference(state)
    ref_shard = cpu.shard(shard)
    run = lambda: cpu.encode(ref_shard)  # noqa: E731
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = shard.node_count * args.steps / dt
    sample = (f"first {shard.record_count} records ({shard.node_count} nt) of the synthetic "
              f"100k x ~200 nt shard per step, fp16 model (package default); {cpu.what}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": f"synthetic; {label}",
        "config": workload_config(args, shard.record_count),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(),
                         "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cpus": os.cpu_count(),
    }
    emit(line)


def workload_config(args, records):
    return {"workload": "BASELINE configs[1]: synthetic shard of 100k RNAs x ~200 nt, "
                        "encode_graphs fp16 model",
            "records_per_gpu": records, "max_batch_nodes": MAX_BATCH_NODES,
            "max_batch_edges": MAX_BATCH_EDGES, "parallelism": f"shard-per-gpu x{args.gpus}",
            "l2_policy": "inputs larger than L2 (1.4 GB of shard arrays per pass)"}


_JSON_OUT = None


def protect_stdout() -> None:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner on stdout at every NCCL_DEBUG level from VERSION up, WARN included), so file descriptor 1
    is pointed at stderr for the whole run and the JSON line goes to a private copy of the
    original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# --------------------------------------------------------------------------------------------
# extra measurements (each returns a JSON-able dict; never raises into the headline)
# --------------------------------------------------------------------------------------------
def guarded(fn, *a, **kw):
    try:
        return fn(*a, **kw)
    except Exception as exc:  # noqa: BLE001 -- an extra must not cost the headline line
        return {"error": f"{type(exc).__name__}: {exc}"[:400]}


def copy_ceiling(device, h2d_bytes: int, d2h_bytes: int, barrier, world, repeats: int = 3):
    """The bytes one end-to-end step moves, with NO kernels: pinned host -> device on one
    stream, device -> pinned host on another, all ranks at once (barrier before), best of
    `repeats`, max over ranks.  This is what bounds `e2e` from above on this box."""
    import torch
    import torch.distributed as dist
    chunk = 256 << 20
    h_in = torch.empty(min(h2d_bytes, chunk), dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(min(d2h_bytes, chunk), dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(h_in.numel(), dtype=torch.uint8, device=device)
    d_out = torch.empty(h_out.numel(), dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)
    best = None
    for _ in range(repeats):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            left = h2d_bytes
            while left > 0:
                n = min(left, h_in.numel())
                d_in[:n].copy_(h_in[:n], non_blocking=True)
                left -= n
        with torch.cuda.stream(s2):
            left = d2h_bytes
            while left > 0:
                n = min(left, h_out.numel())
                h_out[:n].copy_(d_out[:n], non_blocking=True)
                left -= n
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = dt.item() if best is None else min(best, dt.item())
    return best


def parity_check(result, shard, cpu_outputs, count: int, state):
    """First `count` records of the timed end-to-end pass against the reference CPU path
    (fp16 model, package default) and, for the cosine bar, its fp32 model."""
    sub = shard.slice(0, count)
    if cpu_outputs is None:
        cpu = CpuReference(state)
        cpu_outputs = cpu.encode(cpu.shard(sub))
        kind = cpu.kind
    else:
        kind = "reference"
    got = np.concatenate(result[:count]).astype(np.float32)
    want16 = np.concatenate(cpu_outputs[:count]).astype(np.float32)
    cpu32 = CpuReference(state, full_precision=True)
    want32 = np.concatenate(cpu32.encode(cpu32.shard(sub), dtype=np.float32)).astype(np.float64)
    g = got.astype(np.float64)
    cos = (g * want32).sum(1) / (np.linalg.norm(g, axis=1) * np.linalg.norm(want32, axis=1))
    return {"records": count, "nucleotides": int(got.shape[0]),
            "against": f"{kind} CPU path (fp16 model for max_abs, fp32 model for min_cos)",
            "max_abs": float(np.abs(got - want16).max()), "max_abs_bound": 4e-3,
            "min_cos": float(cos.min()), "min_cos_bound": 0.999,
            "ok": bool(np.abs(got - want16).max() <= 4e-3 and cos.min() >= 0.999)}


def reference_cuda_eager(shard, sample_records: int, device: str):
    """The reference's own eager CUDA path on this GPU (api.py:69-76, 110-112, 232-260):
    end to end through its public API, and its forward alone on one microbatch (CUDA events)."""
    import torch
    from oracle import ref_loader
    if ref_loader.reference_root() is None:
        return {"unavailable": "reference not staged (python -m oracle.stage_ref)"}
    ref, _api = ref_loader.import_reference()
    sub = shard.slice(0, min(sample_records, shard.record_count))
    ref_shard = ref_loader.to_reference_shard(ref, sub)
    enc = ref.Ginfinity.load(device=device, allow_nondeterministic_cuda=True)
    run = lambda: enc.encode_graphs(ref_shard, max_batch_nodes=MAX_BATCH_NODES,  # noqa: E731
                                    max_batch_edges=MAX_BATCH_EDGES)
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # forward only, one ~60k-node microbatch, device time
    stop = int(np.searchsorted(sub.node_ptr, MAX_BATCH_NODES, side="right")) - 1
    micro = ref_shard.slice(0, max(1, stop))
    dev = torch.device(device)
    with torch.inference_mode():
        node = torch.from_numpy(micro.node_features).to(dev, torch.float16)
        ei = torch.from_numpy(micro.edge_index).to(dev, torch.long)
        attr = torch.nn.functional.one_hot(torch.from_numpy(micro.edge_types).to(dev, torch.long),
                                           num_classes=10).to(torch.float16)
        for _ in range(2):
            enc._model(node, ei, attr)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            enc._model(node, ei, attr)
        b.record()
        torch.cuda.synchronize()
    fwd_s = a.elapsed_time(b) * 1e-3 / 5
    return {"value": sub.node_count / e2e_s, "unit": UNIT,
            "what": "unmodified reference, Ginfinity.load(device='cuda', "
                    "allow_nondeterministic_cuda=True).encode_graphs, host arrays in and out",
            "sample": f"first {sub.record_count} records ({sub.node_count} nt) of the same shard",
            "forward_only": {"value": micro.node_count / fwd_s, "unit": UNIT,
                             "nodes": int(micro.node_count), "ms": fwd_s * 1e3,
                             "what": "GINEEncoder.forward on one microbatch, inputs resident, "
                                     "CUDA events (cuBLAS / ATen kernels, fp16 model)"}}


def search_bench(device, rank, world, barrier, peaks):
    """K5 at the C5 shape per GPU; at N > 1 the database is row-sharded over the ranks and the
    per-query lists are all-gathered (NCCL) and merged.  Five consecutive runs."""
    import torch
    import torch.distributed as dist
    from ginfinity_b200.search import EmbeddingIndex, gather_lists, merge_lists

    def unit(n, seed):
        g = torch.Generator(device=device).manual_seed(seed)
        a = torch.randn(n, 128, generator=g, device=device)
        return (a / a.norm(dim=1, keepdim=True)).half()

    q = unit(SEARCH_Q, 1)
    index = EmbeddingIndex(unit(SEARCH_ROWS_PER_GPU, 100 + rank), device=device,
                           index_base=rank * SEARCH_ROWS_PER_GPU,
                           total_rows=world * SEARCH_ROWS_PER_GPU)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    runs, scan, gather, merge = [], [], [], []
    for it in range(6):                                    # first run is the warm-up
        barrier()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        s, i = index.search(q, SEARCH_K, "cosine")
        e1.record()
        if world > 1:
            all_s, all_i = gather_lists(s, i, world)
            e2.record()
            s, i = merge_lists(all_s, all_i)
        else:
            e2.record()
        e3.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e3), e0.elapsed_time(e1), e1.elapsed_time(e2),
                          e2.elapsed_time(e3)], device=device, dtype=torch.float64) * 1e-3
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it:
            runs.append(t[0].item()); scan.append(t[1].item())
            gather.append(t[2].item()); merge.append(t[3].item())
    del index
    torch.cuda.empty_cache()
    # sharded == single GPU at a reduced size: every rank builds the same small database
    rows = 65_536 * world
    small_db, small_q = unit(rows, 7), unit(2048, 8)
    full = EmbeddingIndex(small_db, device=device)
    want_s, want_i = full.search(small_q, SEARCH_K, "cosine")
    agree = True
    if world > 1:
        part = EmbeddingIndex.shard(small_db, rank=rank, world_size=world, device=device)
        got_s, got_i = part.search_sharded(small_q, SEARCH_K, "cosine")
        agree = bool(torch.equal(got_i, want_i) and torch.equal(got_s, want_s))
        flag = torch.tensor([1.0 if agree else 0.0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        agree = bool(flag.item() == 1.0)
    total = float(np.median(runs))
    flop = 2.0 * 128 * SEARCH_Q * SEARCH_ROWS_PER_GPU * world
    scan_tflops = 2.0 * 128 * SEARCH_Q * SEARCH_ROWS_PER_GPU / float(np.median(scan)) / 1e12
    return {"shape": {"queries": SEARCH_Q, "rows_per_gpu": SEARCH_ROWS_PER_GPU, "k": SEARCH_K,
                      "metric": "cosine", "dim": 128},
            "seconds": total, "runs_s": [round(v, 5) for v in runs],
            "spread": (max(runs) - min(runs)) / min(runs),
            "scan_s": float(np.median(scan)), "allgather_s": float(np.median(gather)),
            "merge_s": float(np.median(merge)),
            "tflops_whole_job": flop / total / 1e12,
            "roofline": {"bound": "tensor", "achieved": scan_tflops, "peak": peaks["tflops"],
                         "unit": "TFLOP/s", "frac": scan_tflops / peaks["tflops"],
                         "what": "gfx_topk scan + finish on one rank: 256 FLOP per (query, row)"},
            "collective": ("all_gather_into_tensor of [Q,k] fp32 scores + int64 indices (NCCL), "
                           "then gfx_topk_merge") if world > 1 else "none at N = 1",
            "sharded_equals_single_gpu": agree,
            "parity": "unpinned (the reference has no search); checked against this repo's contract"}


def c1_embed(device):
    """BASELINE configs[0]: the reference's 5,840-RNA table through `embed` semantics
    (table in, NPZ + manifest out), wall time including the table read and the archive."""
    from oracle import ref_loader
    table = ref_loader.rouskin_table()
    if table is None:
        return {"unavailable": "tests/rouskin_sample_6k.tsv not staged (python -m oracle.stage_ref)"}
    import ginfinity_b200 as g
    from ginfinity_b200 import cli
    with tempfile.TemporaryDirectory() as tmp:
        out = Path(tmp) / "c1.npz"
        argv = ["embed", "--input", str(table), "--output", str(out), "--device", device,
                "--no-slices"]
        t0 = time.perf_counter()
        rc = cli.main(argv)
        wall = time.perf_counter() - t0
        manifest = json.loads(out.with_suffix(".manifest.json").read_text())
    t0 = time.perf_counter()
    records = g.read_rna_table(table)
    read_s = time.perf_counter() - t0
    enc = g.Ginfinity.load(device)
    enc.encode_many(records)
    t0 = time.perf_counter()
    arrays = enc.encode_many(records)
    enc_s = time.perf_counter() - t0
    nt = int(sum(a.shape[0] for a in arrays))
    return {"workload": "BASELINE configs[0]: rouskin_sample_6k.tsv, 5,840 RNAs, fp16 model on the GPU "
                        "(the config's CPU fp32 reference path is the cpu_baseline / parity side)",
            "rc": rc, "records": len(records), "nucleotides": nt,
            "embed_cli_wall_s": wall, "embed_cli_nt_per_s": nt / wall,
            "manifest_elapsed_s": manifest.get("elapsed_seconds"),
            "table_read_s": read_s, "encode_many_s": enc_s, "encode_many_nt_per_s": nt / enc_s,
            "note": "the CLI's wall time is the compressed NPZ archive (the reference's format; "
                    "members deflated on every host core, ginfinity_b200/npz.py) + table parsing + "
                    "model load + process start; encode_many is strings in, host arrays out"}


def windows_bench(encoder, count: int = 20_000):
    """Windowed records (the query side of BASELINE configs[4]: "query windows"; reference
    graph.py:599-695): the middle third of every RNA is the core, paired neighbours and two hops
    of context are kept; `encode_many` selects, builds the induced subgraphs (K7), encodes them
    with the CSR-walking pair kernel and returns only the core rows."""
    import torch
    import ginfinity_b200 as g
    from ginfinity_b200.synthetic import synthetic_records
    recs = [g.RNA(r.identifier, r.sequence, r.structure, start=len(r.sequence) // 3,
                  end=2 * len(r.sequence) // 3) for r in synthetic_records(11, count)]
    core = sum(r.end - r.start for r in recs)
    run = lambda: encoder.encode_many(recs, max_batch_nodes=MAX_BATCH_NODES,  # noqa: E731
                                      max_batch_edges=MAX_BATCH_EDGES, keep_paired_neighbours=True,
                                      context_hops=2)
    res = run()
    assert len(res) == len(recs) and sum(a.shape[0] for a in res) == core
    del res
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    secs = (time.perf_counter() - t0) / 2
    return {"workload": "windowed records: core = middle third of each synthetic RNA, "
                        "keep_paired_neighbours, context_hops = 2; encode_many, strings in, "
                        "host arrays (core rows only) out", "windows": len(recs),
            "core_nucleotides": core, "seconds": secs, "core_nt_per_s": core / secs,
            "windows_per_s": len(recs) / secs}


def c3_long(encoder, device, count: int):
    """BASELINE configs[2]: long RNAs (1-10 knt, long-range pairs), device-resident encode at
    three microbatch limits; GENERIC-row fraction and time per node against C2."""
    import torch
    from ginfinity_b200 import _native as nat
    from ginfinity_b200.encoder import DeviceShard
    from ginfinity_b200.synthetic import synthetic_shard
    t0 = time.perf_counter()
    shard = synthetic_shard(1, count, log_uniform=True, lo=1000, hi=10_000, prefix="long")
    gen_s = time.perf_counter() - t0
    ds = DeviceShard.from_shard(shard, device)
    n, e = shard.node_count, shard.edge_count
    # row descriptors of the whole shard: how many rows leave the banded fast path
    lib = nat.lib
    st = torch.cuda.current_stream().cuda_stream
    u8 = lambda nbytes: torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)  # noqa: E731
    row_ptr, col_src, col_type = u8(4 * (n + 1)), u8(4 * e), u8(e)
    ws = u8(lib.gfx_csr_workspace_bytes(n, e))
    nat.check(lib.gfx_csr_build(ds.edge_index[0].data_ptr(), ds.edge_index[1].data_ptr(),
                                ds.edge_types.data_ptr(), n, e, 0, row_ptr.data_ptr(),
                                col_src.data_ptr(), col_type.data_ptr(), ws.data_ptr(), ws.numel(), st))
    desc = torch.empty(n, dtype=torch.int32, device=device)
    nat.check(lib.gfx_row_describe(row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(), n,
                                   desc.data_ptr(), st))
    generic = int((desc < 0).sum().item())
    partner = ((desc >> 6) & ((1 << 25) - 1)).long()
    has_pair = (desc & 4) != 0
    idx = torch.arange(n, device=device)
    far = int((has_pair & ((partner >> 7) != (idx >> 7))).sum().item())
    del row_ptr, col_src, col_type, ws, desc, partner
    out = torch.empty((n, 128), dtype=torch.float16, device=device)
    sweeps = {}
    longest = int(np.diff(shard.node_ptr).max())
    for limit in (60_000, 240_000, 1_000_000):
        if limit < longest:
            continue
        run = lambda: encoder.encode_device_shard(ds, max_batch_nodes=limit,  # noqa: E731
                                                  max_batch_edges=5 * limit, out=out)
        run()
        torch.cuda.synchronize()
        nat.profile_enable("fused_layer")
        nat.profile_read("fused_layer")
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run()
        b.record()
        torch.cuda.synchronize()
        layer_ms, calls = nat.profile_read("fused_layer")
        nat.profile_enable()
        ms = a.elapsed_time(b) / 3
        sweeps[str(limit)] = {"nt_per_s": n / (ms * 1e-3), "ms_per_pass": ms,
                              "microbatches": len(encoder.last_microbatch_bounds) - 1,
                              "layer_ns_per_node_layer": layer_ms * 1e6 / (3 * 4 * n) if calls else None}
    return {"workload": "BASELINE configs[2]: synthetic long RNAs, log-uniform 1-10 knt, "
                        "domains enclosed by long-range stems", "records": shard.record_count,
            "nucleotides": n, "edges": e, "longest": longest, "generation_s": gen_s,
            "generic_row_fraction": generic / n,
            "out_of_tile_partner_fraction": far / n,
            "max_batch_nodes_sweep": sweeps}


def c4_files(encoder, device, rank, world, barrier, files: int, records_per_file: int):
    """BASELINE configs[3] as this rank's share of a corpus of shard FILES: `files` safetensors
    shards per GPU read back through encode_shard_files (next file loaded and page-locked on a
    background thread while the current one is encoded); sustained nt/s including file load."""
    import torch
    import torch.distributed as dist
    import ginfinity_b200 as g
    from ginfinity_b200.multi_gpu import encode_shard_files
    from ginfinity_b200.synthetic import synthetic_shard
    import shutil
    need = files * records_per_file * 15_000 * world          # ~14 KB of shard file per record
    base = None
    for candidate in ("/dev/shm", tempfile.gettempdir()):
        try:
            if Path(candidate).is_dir() and shutil.disk_usage(candidate).free > 2 * need:
                base = candidate
                break
        except OSError:
            continue
    root = Path(tempfile.mkdtemp(prefix=f"gfx_c4_r{rank}_", dir=base))
    try:
        paths, nodes, t0, failure = [], 0, time.perf_counter(), None
        try:
            for k in range(files):
                shard = synthetic_shard(1000 * (rank + 1) + k, records_per_file, prefix=f"r{rank}f{k}_")
                path = root / f"shard_{k:03d}.safetensors"
                g.save_graph_shard(shard, path)
                paths.append(path)
                nodes += shard.node_count
        except Exception as exc:  # noqa: BLE001 -- every rank must reach the collective below
            failure = f"{type(exc).__name__}: {exc}"[:200]
        ok = torch.tensor([0.0 if failure else 1.0], device=device)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() != 1.0:
            return {"error": failure or "another rank could not write its shard files"}
        write_s = time.perf_counter() - t0
        nbytes = sum(p.stat().st_size for p in paths)
        barrier()
        # every shard's embeddings are consumed as they arrive (the reference's per-shard job writes
        # its archive and moves on, cli.py:139-197); here: a checksum over every record's first row
        stamps, check = [], [0.0]

        def consume(path, arrays):
            check[0] += float(sum(float(a[0, 0]) for a in arrays[::997]))
            stamps.append(time.perf_counter())

        t0 = time.perf_counter()
        out = encode_shard_files(encoder, paths, rank=0, world_size=1, consume=consume,
                                 max_batch_nodes=MAX_BATCH_NODES, max_batch_edges=MAX_BATCH_EDGES)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        secs = torch.tensor([t1 - t0], device=device, dtype=torch.float64)
        total = torch.tensor([float(nodes)], device=device, dtype=torch.float64)
        # steady state: from the third shard's completion to the last (pinned buffers exist by then)
        steady = torch.tensor([(stamps[-1] - stamps[2]) / max(files - 3, 1)], device=device,
                              dtype=torch.float64)
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
            dist.all_reduce(steady, op=dist.ReduceOp.MAX)
            dist.all_reduce(total, op=dist.ReduceOp.SUM)
        assert sum(out.values()) == files * records_per_file
        return {"workload": "BASELINE configs[3] slice: graph-shard files streamed through "
                            "encode_shard_files (files memory-mapped, copied into reusable "
                            "page-locked buffers and validated on background threads; one reused "
                            "page-locked result table; results consumed per shard), no collective",
                "files_per_gpu": files, "records_per_file": records_per_file,
                "nucleotides_all_gpus": int(total.item()), "file_bytes_per_gpu": nbytes,
                "where": str(root.parent), "seconds": secs.item(),
                "nt_per_s": total.item() / secs.item(),
                "steady_state_nt_per_s": total.item() / files / steady.item(),
                "steady_state_what": "per-shard period from the 3rd shard on (the first shards "
                                     "pay for page-locking the buffer sets, ~0.6 s per GiB)",
                "generation_and_write_s": write_s}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=100_000)
    ap.add_argument("--sample-records", type=int, default=3_000,
                    help="records per step of the CPU reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-records-e2e", dest="records_e2e", action="store_false",
                    help="skip the encode_many (records -> embeddings) measurement")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip search, C1/C3/C4, the reference eager-CUDA bar and the fp32 pass")
    ap.add_argument("--chunk-nodes", type=int, default=0,
                    help="nodes per device chunk (0 = the encoder's default)")
    ap.add_argument("--c3-records", type=int, default=600)
    ap.add_argument("--c4-files", type=int, default=10)
    args = ap.parse_args()
    t_start = time.perf_counter()
    protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ginfinity_b200 import _native as nat
    from ginfinity_b200.encoder import DeviceShard, Ginfinity, pin_shard

    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    from ginfinity_b200.multi_gpu import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank)       # before any pinned allocation
    if world > 1:
        # NCCL's log goes to stderr (protect_stdout() is the backstop); the caller's NCCL_DEBUG
        # is respected so that the driver can read the communicator lines
        os.environ.setdefault("NCCL_DEBUG", os.environ.get("GFX_NCCL_DEBUG", "WARN"))
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(device))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state, label = load_weights()
    shard, gen_s = build_workload(args.records, seed=rank)   # weak scaling: a shard per GPU
    encoder = Ginfinity.from_state(state, device=device)
    if args.chunk_nodes > 0:
        encoder.chunk_nodes = encoder.resident_chunk_nodes = args.chunk_nodes
    nodes, edges = shard.node_count, shard.edge_count

    # ---------------- device-resident throughput (`value`) --------------------
    dshard = DeviceShard.from_shard(shard, device)
    out = torch.empty((nodes, 128), dtype=torch.float16, device=device)
    step = lambda: encoder.encode_device_shard(  # noqa: E731
        dshard, max_batch_nodes=MAX_BATCH_NODES, max_batch_edges=MAX_BATCH_EDGES, out=out)
    sampler = ClockSampler(local_rank)
    sampler.start()                               # samples cover warm-up + timed steps (GPU busy)
    for _ in range(args.warmup):
        step()
    barrier()
    nat.launch_counts(reset=True)
    nat.profile_enable("mlp", "aggregate", "fused_layer")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.finish()
    # rows of the TIMED device-resident passes, kept to be compared bit for bit with what the
    # end-to-end call (other chunk sizes, host path) returns for the same records
    sample_ids = sorted(set(range(min(200, shard.record_count))) | set(range(0, shard.record_count, 997)))
    nptr = shard.node_ptr
    resident_sample = {i: out[int(nptr[i]):int(nptr[i + 1])].cpu().numpy() for i in sample_ids}
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    total_nodes = torch.tensor([float(nodes)], device=device, dtype=torch.float64)
    launches = nat.launch_counts(reset=True)
    launches = {k: int(v) for k, v in launches.items()}
    mlp_ms, mlp_calls = nat.profile_read("mlp")
    agg_ms, agg_calls = nat.profile_read("aggregate")
    fused_ms, fused_calls = nat.profile_read("fused_layer")
    # one more pass, untimed, with every stage bracketed: where the step goes
    nat.profile_enable(*nat.STAGES)
    step()
    torch.cuda.synchronize()
    stage_ms = {name: round(nat.profile_read(name)[0], 4) for name in nat.STAGES}
    stage_ms = {k: v for k, v in stage_ms.items() if v > 0}
    # the two-kernel form of a layer (K1 aggregation, K2 MLP) on the same shard, untimed: the
    # north-star's per-kernel roofline targets are stated for these
    split = None
    if getattr(encoder, "fused", 0) > 0:
        keep_fused, encoder.fused = encoder.fused, 0
        nat.profile_enable()
        step()                                     # warm
        torch.cuda.synchronize()
        nat.profile_enable("mlp", "aggregate")
        nat.profile_read("mlp")                    # reset the accumulators
        nat.profile_read("aggregate")
        step()
        torch.cuda.synchronize()
        split = (nat.profile_read("mlp"), nat.profile_read("aggregate"))
        encoder.fused = keep_fused
    nat.profile_enable()
    nat.launch_counts(reset=True)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_nodes, op=dist.ReduceOp.SUM)
    ms_per_step = elapsed_ms.item() / args.steps
    value = total_nodes.item() / (ms_per_step * 1e-3)
    microbatches = len(encoder.last_microbatch_bounds) - 1

    # ---------------- end to end through the public API (`e2e`) ----------------
    pinned = pin_shard(shard)
    h2d = sum(getattr(shard, n).nbytes for n in
              ("node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr"))
    d2h = nodes * 128 * 2
    run = lambda **kw: encoder.encode_graphs(  # noqa: E731
        pinned, max_batch_nodes=MAX_BATCH_NODES, max_batch_edges=MAX_BATCH_EDGES, **kw)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(1, args.warmup - 1)):
        result = run()
    assert len(result) == shard.record_count and result[0].dtype == np.float16
    check_records = min(200, shard.record_count)
    kept_for_parity = [np.array(a) for a in result[:check_records]]
    resident_equal = all(np.array_equal(result[i], rows) for i, rows in resident_sample.items())
    resident_checked = (len(resident_sample), int(sum(r.shape[0] for r in resident_sample.values())))
    del resident_sample
    del result
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    clocks_e2e = sampler.finish()
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = total_nodes.item() * args.steps / e2e_s.item()
    ceiling_s = copy_ceiling(device, h2d, d2h, barrier, world)
    ceiling_value = total_nodes.item() / ceiling_s

    # -------- the caller KEEPS every step's result (the reference contract: caller owns them) ----
    retained_steps = min(args.steps, 3)
    tables = [Ginfinity.pinned_table(nodes) for _ in range(retained_steps)]   # allocated up front
    kept = []
    barrier()
    t0 = time.perf_counter()
    for k in range(retained_steps):
        kept.append(run(out=tables[k]))
    torch.cuda.synchronize()
    ret_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    # ... and what it costs when the encoder has to page-lock a fresh table inside the call
    barrier()
    t0 = time.perf_counter()
    kept.append(run())                              # previous results alive: nothing to recycle
    torch.cuda.synchronize()
    fresh_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    same = all(np.array_equal(kept[0][i], kept[-1][i]) for i in range(0, shard.record_count, 997))
    del kept, tables
    if world > 1:
        dist.all_reduce(ret_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(fresh_s, op=dist.ReduceOp.MAX)
    e2e_retained = {"value": total_nodes.item() * retained_steps / ret_s.item(), "unit": UNIT,
                    "steps": retained_steps, "ms_per_step": ret_s.item() / retained_steps * 1e3,
                    "what": "encode_graphs(out=table_k): every step's embeddings are kept by the "
                            "caller in its own page-locked table (allocated before the timed region)",
                    "fresh_table_inside_the_call": {
                        "value": total_nodes.item() / fresh_s.item(), "unit": UNIT,
                        "ms": fresh_s.item() * 1e3,
                        "what": "one call with out=None while earlier results are alive: the "
                                "encoder page-locks a new 5 GB table (cudaHostAlloc, ~0.6 s/GiB)"},
                    "bit_identical_across_steps": bool(same)}

    # ------------- records -> embeddings (`encode_many`, graphs built on the GPU) ---------
    from_records = None
    if args.records_e2e:
        from ginfinity_b200.synthetic import synthetic_records
        recs = synthetic_records(rank, args.records)          # the same RNAs the shard was built from
        assert sum(len(r.sequence) for r in recs) == nodes
        run_many = lambda: encoder.encode_many(  # noqa: E731
            recs, max_batch_nodes=MAX_BATCH_NODES, max_batch_edges=MAX_BATCH_EDGES)
        res = run_many()
        assert len(res) == len(recs)
        del res
        barrier()
        nat.launch_counts(reset=True)
        t0 = time.perf_counter()
        for _ in range(2):
            run_many()
        torch.cuda.synchronize()
        many_s = torch.tensor([(time.perf_counter() - t0) / 2], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(many_s, op=dist.ReduceOp.MAX)
        from_records = {"value": total_nodes.item() / many_s.item(), "unit": UNIT,
                        "ms_per_step": many_s.item() * 1e3,
                        "h2d_bytes_per_step": 2 * nodes + 8 * (shard.record_count + 1),
                        "d2h_bytes_per_step": d2h,
                        "what": "Ginfinity.encode_many(records): dot-bracket strings in, host "
                                "arrays out; graphs built by gfx_graph_count/fill on the GPU"}
        del recs

    # ---------------- extras: fp32 path, search, the other configs, the reference on CUDA --------
    peaks = measured_peaks()
    extras = {}
    if args.extras:
        def fp32_pass():
            enc32 = Ginfinity.from_state(state, device=device, full_precision=True)
            o32 = torch.empty((nodes, 128), dtype=torch.float32, device=device)
            go = lambda: enc32.encode_device_shard(  # noqa: E731
                dshard, max_batch_nodes=MAX_BATCH_NODES, max_batch_edges=MAX_BATCH_EDGES,
                out_dtype=nat.GFX_F32, out=o32)
            for _ in range(2):
                go()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                go()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3
            nat.profile_enable(*nat.STAGES)
            go()
            torch.cuda.synchronize()
            stages = {name: round(nat.profile_read(name)[0], 3) for name in nat.STAGES}
            nat.profile_enable()
            nat.launch_counts(reset=True)
            k1_gbs = nodes * enc32._weights.cfg.layers * 1028 / (stages["aggregate"] * 1e-3) / 1e9
            return {"value": nodes / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "n_gpus": 1,
                    "stage_ms": {k: v for k, v in stages.items() if v > 0},
                    "k1_roofline": {"kernel": "aggregate_f32_band_kernel (K1 for fp32 storage from row "
                                              "descriptors: 64-row tiles by bulk copy, register window)",
                                    "bound": "hbm", "achieved": k1_gbs, "peak": peaks["hbm_gbs"],
                                    "unit": "GB/s", "frac": k1_gbs / peaks["hbm_gbs"],
                                    "algorithmic_bytes_per_node_layer": 1028},
                    "what": "full_precision=True (fp32 storage end to end, reference api.py:110-112) "
                            "on this rank's shard, device-resident, mean of three timed passes"}
        extras["value_fp32"] = guarded(fp32_pass)
        del dshard, out
        torch.cuda.empty_cache()
        extras["search"] = guarded(search_bench, device, rank, world, barrier, peaks)
        torch.cuda.empty_cache()
        extras["c4_shard_files"] = guarded(c4_files, encoder, device, rank, world, barrier,
                                           args.c4_files, max(1000, args.records // 10))
        if rank == 0:
            extras["c1_embed"] = guarded(c1_embed, device)
            extras["c3_long_rnas"] = guarded(c3_long, encoder, device, args.c3_records)
            extras["windows"] = guarded(windows_bench, encoder)
            extras["reference_cuda_eager"] = guarded(reference_cuda_eager, shard,
                                                     args.sample_records, device)
        barrier()

    if rank == 0:
        # per-launch algorithmic work / average launch duration == totals ratio
        node_layers = nodes * 4 * args.steps          # 4 layers per step (this rank)
        mlp_tflops = node_layers * FLOP_PER_NODE_MLP / (mlp_ms * 1e-3) / 1e12 if mlp_ms else 0.0
        agg_gbs = node_layers * BYTES_PER_NODE_AGG / (agg_ms * 1e-3) / 1e9 if agg_ms else 0.0
        step_ms_rank = ev0.elapsed_time(ev1) / args.steps
        roofline_mlp = {"kernel": "umma4_mlp_kernel (K2: MLP + LayerNorm + residual, tcgen05 + TMA)",
                        "bound": "tensor", "achieved": mlp_tflops, "peak": peaks["tflops"],
                        "unit": "TFLOP/s", "frac": mlp_tflops / peaks["tflops"],
                        "traffic": ncu_traffic("mlp", node_layers / max(mlp_calls, 1)),
                        "launches": mlp_calls,
                        "avg_launch_ms": mlp_ms / max(mlp_calls, 1),
                        "share_of_step": mlp_ms / args.steps / step_ms_rank,
                        "peak_source": peaks["source"]}
        roofline_agg = {"kernel": "aggregate_f16_q_kernel (K1: CSR gather + table + ReLU + self term, quarter-warp per node)",
                        "bound": "hbm", "achieved": agg_gbs, "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": agg_gbs / peaks["hbm_gbs"],
                        "traffic": ncu_traffic("aggregate", node_layers / max(agg_calls, 1)),
                        "launches": agg_calls, "avg_launch_ms": agg_ms / max(agg_calls, 1),
                        "share_of_step": agg_ms / args.steps / step_ms_rank,
                        "peak_source": peaks["source"]}
        dominant, other = ((roofline_mlp, roofline_agg) if mlp_ms >= agg_ms
                           else (roofline_agg, roofline_mlp))
        if fused_ms:
            # K1 + K2 as one kernel on CTA pairs: the layer's dense FLOPs against the tensor peak
            # (its HBM side -- h in, h' out, row descriptors: 516 B per node-layer -- is the "other")
            fl = node_layers * FLOP_PER_NODE_MLP / (fused_ms * 1e-3) / 1e12
            banded = getattr(encoder, "fused", 0) == 3
            gb = node_layers * (BYTES_PER_NODE_BANDED if banded else BYTES_PER_NODE_FUSED) / (fused_ms * 1e-3) / 1e9
            common = {"launches": fused_calls, "avg_launch_ms": fused_ms / max(fused_calls, 1),
                      "share_of_step": fused_ms / args.steps / step_ms_rank,
                      "peak_source": peaks["source"]}
            name = (("fused_banded_kernel" if banded else "fused_pair_kernel") +
                    " (K1 + K2 in one kernel: aggregation producing the tcgen05 "
                    "cta_group::2 A operand, MLP + LayerNorm + residual)")
            dominant = {"kernel": name, "bound": "tensor", "achieved": fl, "peak": peaks["tflops"],
                        "unit": "TFLOP/s", "frac": fl / peaks["tflops"],
                        "traffic": ncu_traffic("fused_banded" if banded else "fused_layer",
                                               node_layers / max(fused_calls, 1)),
                        **common}
            other = {"kernel": name + " -- HBM side", "bound": "hbm", "achieved": gb,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gb / peaks["hbm_gbs"],
                     "traffic": ncu_traffic("fused_banded" if banded else "fused_layer",
                                            node_layers / max(fused_calls, 1)),
                     **common}
        split_line = None
        if split is not None:
            (s_mlp_ms, s_mlp_calls), (s_agg_ms, s_agg_calls) = split
            nl = nodes * 4
            split_line = {
                "what": "one untimed pass with the layer as two kernels (GFX_FUSED=0)",
                "k1_aggregate": {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"],
                                 "achieved": nl * BYTES_PER_NODE_AGG / (s_agg_ms * 1e-3) / 1e9,
                                 "frac": nl * BYTES_PER_NODE_AGG / (s_agg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                 "launches": s_agg_calls, "ms_per_pass": s_agg_ms},
                "k2_mlp": {"bound": "tensor", "unit": "TFLOP/s", "peak": peaks["tflops"],
                           "achieved": nl * FLOP_PER_NODE_MLP / (s_mlp_ms * 1e-3) / 1e12,
                           "frac": nl * FLOP_PER_NODE_MLP / (s_mlp_ms * 1e-3) / 1e12 / peaks["tflops"],
                           "launches": s_mlp_calls, "ms_per_pass": s_mlp_ms}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = guarded(cpu_reference_rate, state, shard, args.sample_records)
        cpu_outputs = cpu["outputs"] if cpu and "outputs" in cpu and cpu["kind"] == "reference" else None
        parity = guarded(parity_check, kept_for_parity, shard, cpu_outputs, check_records, state)
        if isinstance(parity, dict):
            parity["timed_resident_pass_equals_e2e_bit_for_bit"] = {
                "ok": bool(resident_equal), "records": resident_checked[0],
                "nucleotides": resident_checked[1],
                "what": "rows written by the timed device-resident passes (`value`) against the "
                        "end-to-end call's host arrays (`e2e`) for the same records"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": f"synthetic; {label}",
            "config": {**workload_config(args, shard.record_count),
                       "nodes_per_gpu": nodes, "edges_per_gpu": edges,
                       "microbatches_per_step": microbatches,
                       "chunk_nodes": encoder._chunk_limit(MAX_BATCH_NODES, resident=True),
                       "chunk_nodes_e2e": encoder._chunk_limit(MAX_BATCH_NODES)},
            "clocks": clocks, "clocks_e2e": clocks_e2e,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s.item() / args.steps * 1e3,
                    "copy_ceiling": {"value": ceiling_value, "unit": UNIT, "ms": ceiling_s * 1e3,
                                     "gb_per_s_all_gpus": world * (h2d + d2h) / ceiling_s / 1e9,
                                     "what": "the same bytes per step as raw pinned H2D + D2H copies "
                                             "on two streams, no kernels, all ranks at once"},
                    "frac_of_copy_ceiling": e2e_value / ceiling_value},
            "e2e_retained": e2e_retained,
            "parity_check": parity,
            "gpu_launches": int(sum(launches.values())),
            "gpu_launches_by_stage": launches,
            "stage_ms_per_step": stage_ms,
            "e2e_from_records": from_records,
            "roofline": dominant, "roofline_other": other, "roofline_two_kernel_layer": split_line,
            "layer_kernel": {"chosen": {3: "fused_banded_kernel", 2: "fused_pair_kernel"}.get(
                                 getattr(encoder, "fused", 0), "K1 + K2"),
                             "how": "fixed (banded fused layer; the bit-identical pair kernel for "
                                    "shards with context nodes); GFX_FUSED=auto times the forms, "
                                    "GFX_FUSED=0|2|3 pins one"},
            "setup": {"workload_generation_s": gen_s, "host_cpus": os.cpu_count(), "numa_binding": numa,
                      "whole_run_wall_s": round(time.perf_counter() - t_start, 1)},
            **extras,
        }
        if cpu is not None:
            if "error" in cpu:
                line["cpu_baseline"] = cpu
            else:
                line["cpu_baseline"] = {
                    "value": cpu["rate"], "unit": UNIT, "cores": cpu["threads"], "kind": cpu["kind"],
                    "sample": f"first {cpu['sub'].record_count} records ({cpu['sub'].node_count} nt) "
                              f"of the same shard, fp16 model, {cpu['secs']:.1f} s of CPU work; "
                              f"{cpu['what']}"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
