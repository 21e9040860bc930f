"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: shard assignment for
the collective-free encode, and the one exchange step of the similarity search
(row sharding -> local lists -> all-gather -> merge).  The local search and
the merge are GPU kernels in the product; here the ORACLE stands in for them so
the plumbing (row ranges, index bases, gather layout, rank order) is what is
tested.  The bench's max-over-ranks / sum-over-ranks reduction is covered too."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from ginfinity_b200 import multi_gpu, search
    from oracle import gine_oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert multi_gpu.rank_info() == (rank, rank, world)
        rng = np.random.default_rng(0)                       # same data on every rank
        db = rng.normal(size=(1001, 128)).astype(np.float16)
        q = rng.normal(size=(37, 128)).astype(np.float16)
        k = 6
        lo, hi = search.shard_bounds(len(db), world, rank)
        s, i = O.topk_exact(q, db[lo:hi], k, "cosine", index_base=lo)
        all_s, all_i = search.gather_lists(torch.from_numpy(s), torch.from_numpy(i), world)
        assert tuple(all_s.shape) == (world, 37, k)
        # slice r of the gathered tensor is rank r's list
        assert torch.equal(all_s[rank], torch.from_numpy(s))
        got_s, got_i = O.merge_topk(all_s.numpy(), all_i.numpy(), k)
        want_s, want_i = O.topk_exact(q, db, k, "cosine")
        np.testing.assert_array_equal(got_i, want_i)
        np.testing.assert_array_equal(got_s, want_s)
        # bench.py's reduction: time = max over ranks, work = sum over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        n = torch.tensor([100.0 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        assert t.item() == 10.0 + world - 1 and n.item() == 100.0 * world * (world + 1) / 2
        # collective-free encode: every shard is taken by exactly one rank
        mine = multi_gpu.assign_shards([500, 100, 400, 300, 50, 250], world)[rank]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        assert sorted(sum(gathered, [])) == list(range(6))
        (Path(tmp) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_search_exchange_and_shard_assignment(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").is_file() for r in range(world))


def test_assign_shards_is_balanced_and_deterministic():
    sys.path.insert(0, str(ROOT))
    from ginfinity_b200.multi_gpu import assign_shards
    rng = np.random.default_rng(1)
    counts = rng.integers(1_000_000, 30_000_000, size=50).tolist()
    for world in (1, 2, 4, 8):
        parts = assign_shards(counts, world)
        assert sorted(sum(parts, [])) == list(range(50))
        loads = [sum(counts[i] for i in p) for p in parts]
        assert max(loads) <= 1.1 * (sum(counts) / world) + max(counts) * (world > 1) * 0.2
        assert parts == assign_shards(counts, world)
    assert assign_shards([], 4) == [[], [], [], []]
    with pytest.raises(ValueError):
        assign_shards([1], 0)


def test_shard_bounds_cover_the_database_exactly():
    sys.path.insert(0, str(ROOT))
    from ginfinity_b200.search import shard_bounds
    for rows in (0, 1, 7, 1000, 10**8):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(rows, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_numa_binding_helper_never_raises_and_reports():
    """On a box without a GPU (this runner) the helper must say why it did
    nothing instead of failing; the CPU-list parser is exercised directly."""
    sys.path.insert(0, str(ROOT))
    from ginfinity_b200.multi_gpu import _cpu_list, bind_to_gpu_numa_node
    assert _cpu_list("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _cpu_list("") == set()
    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa_node(0)
    assert set(info) == {"node", "cpus", "reason"}
    if info["node"] is None:
        assert info["reason"] and os.sched_getaffinity(0) == before


def test_shard_prefetcher_order_errors_and_overlap():
    """The prefetcher yields every file once, in order, loads ahead of the
    consumer, and re-raises a loader failure on the consumer's thread."""
    import time
    sys.path.insert(0, str(ROOT))
    from ginfinity_b200.multi_gpu import ShardPrefetcher
    loaded = []

    def load(path):
        loaded.append(path)
        time.sleep(0.05)
        return {"path": path}

    paths = [f"shard-{i}" for i in range(5)]
    seen = []
    t0 = time.perf_counter()
    for path, shard in ShardPrefetcher(paths, pin=False, load=load):
        if len(seen) == 1:
            time.sleep(0.15)
            assert len(loaded) >= 3          # the loader ran ahead while the consumer was busy
        time.sleep(0.05)
        seen.append((path, shard["path"]))
    assert seen == [(p, p) for p in paths]
    assert time.perf_counter() - t0 < 5 * 0.1 + 0.15 + 0.2   # overlapped, not serial

    def broken(path):
        if path.endswith("2"):
            raise OSError("truncated file")
        return path

    got = []
    with pytest.raises(OSError, match="truncated"):
        for path, _ in ShardPrefetcher(paths, pin=False, load=broken):
            got.append(path)
    assert got == paths[:2]
    assert list(ShardPrefetcher([], pin=False, load=load)) == []


def test_bench_keeps_stdout_to_one_json_line():
    """bench.py's contract is ONE JSON line on stdout; NCCL writes its version banner to file
    descriptor 1 at every NCCL_DEBUG level from VERSION up.  protect_stdout() points fd 1 at stderr
    and emit() writes the line to a private copy of the original stdout."""
    import json
    import subprocess
    import sys
    import textwrap
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    code = textwrap.dedent('''
        import importlib.util, os, sys
        spec = importlib.util.spec_from_file_location("bench", "bench.py")
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
        bench.protect_stdout()
        os.write(1, b"NCCL version 2.28.9+cuda12.9\\n")     # what a C library does
        print("a stray Python print")
        bench.emit({"metric": "m", "value": 1.5})
    ''')
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root)
    assert run.returncode == 0, run.stderr
    assert json.loads(run.stdout) == {"metric": "m", "value": 1.5}
    assert run.stdout.count("\n") == 1
    assert "NCCL version" in run.stderr and "stray" in run.stderr


def test_chunks_cover_every_microbatch_and_ramp_up():
    """Host logic of the streaming encode: consecutive microbatches are grouped into device
    chunks of at most `limit` nodes (always at least one microbatch); with `ramp` the first
    chunks hold 1/8, 1/4, 1/2 of the limit so the device->host copies start early."""
    from ginfinity_b200.encoder import Ginfinity
    rng = np.random.default_rng(0)
    sizes = rng.integers(40_000, 60_000, 335)
    node_at = np.concatenate([[0], np.cumsum(sizes)])
    plan = np.stack([np.arange(336), node_at, node_at * 5])
    limit = 960_000
    for ramp in (False, True):
        chunks = Ginfinity._chunks(None, plan, limit, ramp=ramp)
        assert chunks[0][0] == 0 and chunks[-1][1] == 335
        assert all(a[1] == b[0] for a, b in zip(chunks[:-1], chunks[1:]))
        nodes = [int(node_at[b] - node_at[a]) for a, b in chunks]
        assert all(b > a for a, b in chunks) and max(nodes) <= limit
        if ramp:
            assert nodes[0] <= limit // 8 and nodes[1] <= limit // 4 and nodes[2] <= limit // 2
            assert min(nodes[3:-1]) > limit - 60_000
        else:
            assert min(nodes[:-1]) > limit - 60_000
    # a microbatch larger than the (ramped) cap still gets its own chunk
    one = Ginfinity._chunks(None, plan, 10_000, ramp=True)
    assert one == [(i, i + 1) for i in range(335)]
