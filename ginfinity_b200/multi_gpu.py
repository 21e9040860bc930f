"""Host-side plumbing for one-process-per-GPU runs (torchrun / torch.distributed).

Encoding partitions by graph shard with NO collective: graphs never interact
(reference docs/GRAPH_PIPELINE.md:22-24 hands independent shard files to an
external scheduler), so each rank takes whole shards, balanced by node count.
Similarity search shards database rows (search.shard_bounds) and has exactly
one exchange step, the all-gather of per-query top-k lists (search.py).
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple


def rank_info() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def _cpu_list(text: str) -> set:
    """'0-15,64-79' -> {0..15, 64..79}"""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off.

    The end-to-end encode moves 325 B per nucleotide between pinned host memory
    and the GPU.  Pinned pages are placed on the node of the allocating thread;
    if that is the other socket every copy crosses the inter-socket link, which
    the ranks of that socket then share (measured on an 8-GPU box: 41 instead
    of ~190 M nt/s per GPU).  Call this before the first pinned allocation.
    Returns what was done ({"node": .., "cpus": ..}) or {"node": None, ...} when
    the topology cannot be read; never raises."""
    info = {"node": None, "cpus": None, "reason": None}
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        if hasattr(props, "pci_bus_id"):
            bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        else:                                      # older torch: ask NVML
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device_index]) if visible else device_index
            raw = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
            raw = raw.decode() if isinstance(raw, bytes) else raw
            bus = raw.lower()[-12:]                # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            info["reason"] = "the platform reports no NUMA node for the GPU"
            return info
        cpus = _cpu_list(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            info["reason"] = "no allowed CPU on the GPU's node"
            return info
        os.sched_setaffinity(0, allowed)
        info.update(node=node, cpus=len(allowed))
    except Exception as exc:                       # topology files missing, permissions, ...
        info["reason"] = f"{type(exc).__name__}: {exc}"
    return info


def assign_shards(node_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Deterministic longest-first greedy assignment of shards to ranks,
    balanced by node count (ties: lower shard index first, lower rank first).
    Returns, per rank, the shard indices in ascending order."""
    if world_size < 1:
        raise ValueError("world_size must be positive")
    order = sorted(range(len(node_counts)), key=lambda i: (-int(node_counts[i]), i))
    load = [0] * world_size
    mine: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        load[r] += int(node_counts[i])
        mine[r].append(i)
    return [sorted(m) for m in mine]


class _PinnedSet:
    """One shard's worth of page-locked buffers, grown on demand and reused file after file
    (page-locking costs ~0.6 s per GiB: it must not be paid per shard)."""

    def __init__(self):
        self._buffers = {}

    def allocate(self, name, shape, dtype):
        import numpy as np
        import torch
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize if shape else dtype.itemsize
        have = self._buffers.get(name)
        if have is None or have.numel() < nbytes:
            grow = max(nbytes + nbytes // 8, 4096)
            have = torch.empty((grow + 63) & ~63, dtype=torch.uint8, pin_memory=True)
            self._buffers[name] = have
        return have.numpy()[:nbytes].view(dtype).reshape(shape)


class ShardPrefetcher:
    """Iterate over shard files in order while `workers` background threads read the NEXT files:
    each file is memory-mapped and its tensors are copied once into a reusable set of page-locked
    buffers (`pin`), validated, and handed over in file order (shard I/O that
    keeps a GPU fed).  At most `depth` loaded shards wait; a shard's buffers are recycled when the
    consumer asks for the next one, so a yielded shard must not be kept across iterations."""

    def __init__(self, paths: Sequence, *, pin: bool = True, depth: int = 2, load=None,
                 device_index: int = None, workers: int = 4):
        import threading
        self._paths = [str(p) for p in paths]
        self._pin, self._load, self._device_index = pin, load, device_index
        self._cond = threading.Condition()
        self._next = 0                     # next file index to hand to a worker
        self._done = {}                    # index -> (shard, buffer set, exception)
        workers = max(1, min(int(workers), len(self._paths) or 1))
        self._free = [_PinnedSet() for _ in range(workers + max(1, int(depth)) + 1)] if pin else None
        self._threads = [threading.Thread(target=self._work, name=f"gfx-shard-prefetch-{k}",
                                          daemon=True) for k in range(workers)]
        for t in self._threads:
            t.start()

    def _work(self) -> None:
        load = self._load
        if load is None:
            from .graph import load_graph_shard as load
        if self._pin and self._device_index is not None:
            import torch
            torch.cuda.set_device(self._device_index)
        while True:
            with self._cond:
                # indices and buffer sets are handed out together, in order: the lowest outstanding
                # file always owns a set, so the consumer can always make progress
                while self._next < len(self._paths) and self._pin and not self._free:
                    self._cond.wait()
                if self._next >= len(self._paths):
                    return
                index = self._next
                self._next += 1
                buffers = self._free.pop() if self._pin else None
            try:
                if buffers is not None and self._load is None:
                    shard = load(self._paths[index], allocate=buffers.allocate)
                else:
                    shard = load(self._paths[index])
                    if self._pin:
                        from .encoder import pin_shard
                        shard = pin_shard(shard)
                result = (shard, buffers, None)
            except BaseException as exc:              # surfaced on the consumer's thread
                result = (None, buffers, exc)
            with self._cond:
                self._done[index] = result
                self._cond.notify_all()

    def __iter__(self):
        held = None
        for index, path in enumerate(self._paths):
            with self._cond:
                if held is not None:                  # the previous shard's buffers go back
                    self._free.append(held)
                    held = None
                    self._cond.notify_all()
                while index not in self._done:
                    self._cond.wait()
                shard, held, exc = self._done.pop(index)
            if exc is not None:
                raise exc
            yield path, shard
        with self._cond:
            if held is not None:
                self._free.append(held)
                self._cond.notify_all()


def shard_file_node_count(path) -> int:
    """Node total of a graph-shard file from the safetensors header alone (8-byte length + a
    small JSON header): the row count of `node_features`.  The JSON sidecar also has it, but
    carries every sequence and structure string of the shard (megabytes to parse)."""
    import json
    with open(path, "rb") as fh:
        size = int.from_bytes(fh.read(8), "little")
        header = json.loads(fh.read(size).decode("utf-8"))
    return int(header["node_features"]["shape"][0])


def encode_shard_files(encoder, paths: Sequence, *, rank: int, world_size: int,
                       node_counts: Sequence[int] = None, prefetch: bool = True,
                       consume=None, workers: int = 2, **encode_kwargs) -> Dict[str, list]:
    """Encode this rank's share of a list of graph-shard files (the
    reference's `embed-graphs` unit of work, cli.py:139-197).  Every rank must
    pass the same `paths`; when `node_counts` is not given the safetensors
    headers are read for the node totals.  With `prefetch` the next files are
    memory-mapped, copied into reusable page-locked buffers and validated on
    background threads while the current one is on the GPU.

    Results.  Every shard's embeddings land in ONE page-locked table that is
    reused from shard to shard (page-locking a fresh table per shard costs
    ~0.6 s per GiB -- more than the encode).  `consume(path, arrays)`, when
    given, is called with row-range views of that table and must be done with
    them when it returns (write them out, reduce them, send them on); the
    function then returns {path: record count}.  Without `consume` the arrays
    are copied out into ordinary memory and returned as {path: [embeddings per
    record]}."""
    import numpy as np

    from .encoder import Ginfinity, split_rows
    from .graph import load_graph_shard
    paths = [str(p) for p in paths]
    if node_counts is None:
        node_counts = [shard_file_node_count(p) for p in paths]
    share = assign_shards(node_counts, world_size)[rank]
    mine = [paths[i] for i in share]
    dtype = np.dtype(encode_kwargs.get("embedding_dtype", np.float16))
    table = None
    out = {}

    def run(path, shard):
        nonlocal table
        if dtype.itemsize > 4:                      # float64: the encoder converts on the host
            arrays = encoder.encode_graphs(shard, **encode_kwargs)
        else:
            rows = shard.node_count if shard.all_core else int(shard.core_count_array().sum())
            if table is None or table.shape[0] < rows:
                table = Ginfinity.pinned_table(max(rows, max(node_counts[i] for i in share)), dtype)
            arrays = encoder.encode_graphs(shard, out=table[:rows], **encode_kwargs)
        if consume is not None:
            consume(path, arrays)
            out[path] = len(arrays)
        elif dtype.itemsize > 4:
            out[path] = arrays
        else:
            kept = np.array(table[:rows])           # one copy out of the reused pinned table
            ptr = np.zeros(len(arrays) + 1, np.int64)
            np.cumsum([a.shape[0] for a in arrays], out=ptr[1:])
            out[path] = split_rows(kept, ptr)

    if prefetch:
        index = getattr(getattr(encoder, "_torch_device", None), "index", None)
        for path, shard in ShardPrefetcher(mine, device_index=index, workers=workers, depth=1):
            run(path, shard)
    else:
        for path in mine:
            run(path, load_graph_shard(path))
    return out
