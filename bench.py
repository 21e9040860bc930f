#!/usr/bin/env python
"""Headline benchmark: nucleotides/s embedded (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): a synthetic shard of 100,000 RNAs x ~200
nt per GPU, `encode_graphs` with the fp16 model at max_batch_nodes=60,000 /
max_batch_edges=300,000.  One step = one pass of the hot path over the whole
shard: on-device microbatch packing, destination-CSR build, input projection,
4 x (aggregation + MLP/LayerNorm/residual), head + L2 normalise.

  value     nt/s with the shard's arrays already resident in HBM (device
            timed with CUDA events; inputs 1.4 GB > L2 so no flush needed)
  e2e       the same through the public `Ginfinity.encode_graphs` call with
            HOST buffers: pinned host -> device copies of every input array
            and device -> host copy of every embedding inside the timed region
  roofline  the dominant kernel of the step against the measured B200 peak
  cpu_baseline / --impl reference
            the reference's CPU algorithm (oracle/cpu_port.py: the same torch
            ops in the same order; the reference itself cannot travel to the
            GPU box) on this box's host cores, on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
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


def cpu_port_rate(state, shard, sample_records: int, repeats: int = 1):
    """nt/s of the reference's CPU algorithm on the first `sample_records`
    records of the workload, all host threads."""
    import torch
    from oracle.cpu_port import CpuPort
    sub = shard.slice(0, min(sample_records, shard.record_count))
    port = CpuPort(state)                           # fp16 model = package default
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        port.encode_graphs(sub, max_batch_nodes=MAX_BATCH_NODES,
                           max_batch_edges=MAX_BATCH_EDGES)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sub.node_count / best, sub, torch.get_num_threads(), best


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is given
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    state, label = load_weights()
    shard, _ = build_workload(args.sample_records, seed=0)
    from oracle.cpu_port import CpuPort
    port = CpuPort(state)
    run = lambda: port.encode_graphs(shard, max_batch_nodes=MAX_BATCH_NODES,  # noqa: E731
                                     max_batch_edges=MAX_BATCH_EDGES)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = shard.node_count * args.steps / dt
    sample = (f"first {shard.record_count} records ({shard.node_count} nt) of the synthetic "
              f"100k x ~200 nt shard per step, fp16 model, torch CPU ops")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": f"synthetic; {label}",
        "config": workload_config(args, shard.record_count),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
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
    ap.add_argument("--chunk-nodes", type=int, default=0,
                    help="nodes per device chunk (0 = the encoder's default)")
    args = ap.parse_args()
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
        # NCCL's log (version banner included) belongs on stderr; protect_stdout() is the backstop
        os.environ["NCCL_DEBUG"] = os.environ.get("GFX_NCCL_DEBUG", "WARN")
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
    run = lambda: encoder.encode_graphs(  # noqa: E731
        pinned, max_batch_nodes=MAX_BATCH_NODES, max_batch_edges=MAX_BATCH_EDGES)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(1, args.warmup - 1)):
        result = run()
    assert len(result) == shard.record_count and result[0].dtype == np.float16
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

    if rank == 0:
        peaks = measured_peaks()
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
            # (its HBM side -- h in, h' out, CSR entries: 539 B per node-layer -- is the "other")
            fl = node_layers * FLOP_PER_NODE_MLP / (fused_ms * 1e-3) / 1e12
            banded = getattr(encoder, "fused", 0) == 3
            gb = node_layers * (BYTES_PER_NODE_BANDED if banded else BYTES_PER_NODE_FUSED) / (fused_ms * 1e-3) / 1e9
            common = {"launches": fused_calls, "avg_launch_ms": fused_ms / max(fused_calls, 1),
                      "share_of_step": fused_ms / args.steps / step_ms_rank,
                      "peak_source": peaks["source"]}
            name = (("fused_banded_kernel" if getattr(encoder, "fused", 0) == 3 else "fused_pair_kernel") +
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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": f"synthetic; {label}",
            "config": {**workload_config(args, shard.record_count),
                       "nodes_per_gpu": nodes, "edges_per_gpu": edges,
                       "microbatches_per_step": microbatches,
                       "chunk_nodes": max(encoder.chunk_nodes, encoder.resident_chunk_nodes),
                       "chunk_nodes_e2e": encoder.chunk_nodes},
            "clocks": clocks, "clocks_e2e": clocks_e2e,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s.item() / args.steps * 1e3},
            "gpu_launches": int(sum(launches.values())),
            "gpu_launches_by_stage": launches,
            "stage_ms_per_step": stage_ms,
            "e2e_from_records": from_records,
            "roofline": dominant, "roofline_other": other, "roofline_two_kernel_layer": split_line,
            "layer_kernel": {"chosen": {3: "fused_banded_kernel", 2: "fused_pair_kernel"}.get(
                                 getattr(encoder, "fused", 0), "K1 + K2"),
                             "tuning_ms_2e19_nodes": {{3: "fused_banded", 2: "fused_pair"}.get(k, "k1_k2"): round(v, 4)
                                                      for k, v in (encoder.layer_kernel_times or {}).items()},
                             "how": "the forms of the layer timed once per device on a fixed "
                                    "synthetic chunk; GFX_FUSED pins the choice"},
            "setup": {"workload_generation_s": gen_s, "host_cpus": os.cpu_count(), "numa_binding": numa},
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, sub, threads, secs = cpu_port_rate(state, shard, args.sample_records)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"first {sub.record_count} records ({sub.node_count} nt) of the same "
                          f"shard, fp16 model, {secs:.1f} s of CPU work"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
