"""Graph construction on the GPU for full-molecule records.

`build_device_shard(records, device)` is `GraphBuilder.build_shard(records)`
(reference: src/ginfinity/graph.py:494-567 + GraphShard.from_graphs,
graph.py:376-412) with the arrays born in HBM: the host sends 2 bytes per
nucleotide (sequence and structure characters) instead of 69 bytes of
features, edge indices and edge types, and one thread per record / per
nucleotide does the pairing and the fill (gfx_graph_count, gfx_graph_fill).
Every array is bit-identical to the host builder's and therefore to the
reference's.  The sin/cos position columns are tabulated on the host with
NumPy once per distinct record length (NumPy's float32 sin/cos is not
correctly rounded, so only NumPy reproduces it) and gathered on the device.

Windowed (sliced) records go through `gfx_slice_select` / `gfx_slice_fill`
(reference: graph.py:599-695): the full molecules' strings
go up, the window, its pairing partners and the breadth-first context are
selected per record on the device and only the induced subgraphs are written.
Any graph specification other than the bundled one keeps the host path
(`GraphBuilder`).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _native as nat
from .encoder import DeviceShard
from .graph import GraphSpec, GraphValidationError

_BUNDLED = dict(struct_feature="A", positional=True)


def supports(spec: GraphSpec, records: Sequence) -> bool:
    """True when the device builder covers this request: the bundled feature
    layout (full-molecule and windowed records alike)."""
    if spec.node_feature_dim != 7 or spec.struct_feature != "A" or not spec.positional:
        return False
    if any(e not in ("skip2",) for e in spec.extra_edges):
        return False
    return True


def any_sliced(records: Sequence) -> bool:
    return any(getattr(r, "sliced", False) for r in records)


def position_table(lengths: np.ndarray):
    """(table float32 [T,2], offset int64 [B]): rows of (sin, cos) of
    float32(pi) * float32(i) / float32(max(L-1, 1)) for every distinct length,
    the reference's exact float32 expression (graph.py:510-514)."""
    uniq, inverse = np.unique(lengths, return_inverse=True)
    starts = np.zeros(uniq.shape[0] + 1, np.int64)
    np.cumsum(uniq, out=starts[1:])
    which = np.repeat(np.arange(uniq.shape[0], dtype=np.int64), uniq)
    local = np.arange(int(starts[-1]), dtype=np.int64) - starts[:-1][which]
    denom = np.maximum(uniq - 1, 1)[which]
    relative = local.astype(np.float32) / denom.astype(np.float32)
    angle = np.float32(np.pi) * relative
    table = np.empty((int(starts[-1]), 2), np.float32)
    table[:, 0] = np.sin(angle)
    table[:, 1] = np.cos(angle)
    return table, starts[:-1][inverse]


def _pinned_bytes(text: str) -> torch.Tensor:
    """ASCII characters of `text` in page-locked memory (so the copy to the
    device is a real asynchronous DMA)."""
    raw = text.encode("ascii")
    t = torch.empty(len(raw), dtype=torch.uint8, pin_memory=True)
    t.numpy()[:] = np.frombuffer(raw, np.uint8)
    return t


def build_device_shard(records: Sequence, device, spec: GraphSpec = None,
                       *, with_node_metadata: bool = False, lengths=None,
                       keep_paired_neighbours: bool = False,
                       context_hops: int = 1) -> DeviceShard:
    """Device-resident shard of the graphs of `records` (objects with
    `.sequence`, `.structure` and, for windowed records, `.start` / `.end`),
    i.e. `GraphBuilder(spec, keep_paired_neighbours=..., context_hops=...)
    .build_shard(records)` with the arrays born in HBM.  Raises
    GraphValidationError for characters outside ACGU / ().  and for
    unbalanced structures."""
    spec = spec if spec is not None else GraphSpec()
    records = list(records)
    if not records:
        raise GraphValidationError("a graph shard needs at least one record")
    if context_hops < 1:
        raise ValueError("context_hops must be >= 1")
    if not supports(spec, records):
        raise GraphValidationError("the device builder covers the bundled graph "
                                   "specification only")
    if any_sliced(records):
        return _build_sliced(records, device, spec, bool(keep_paired_neighbours),
                             int(context_hops))
    dev = torch.device(device)
    sequences = [r.sequence for r in records]
    structures = [r.structure for r in records]
    B = len(records)
    if lengths is None:
        lengths = np.fromiter(map(len, sequences), np.int64, B)
    if np.any(lengths < 1) or list(map(len, structures)) != lengths.tolist():
        raise GraphValidationError("sequence and structure lengths must match and be positive")
    node_ptr = np.zeros(B + 1, np.int64)
    np.cumsum(lengths, out=node_ptr[1:])
    N = int(node_ptr[-1])
    if N > 1 << 30:
        raise GraphValidationError("shard exceeds 2^30 nucleotides")
    try:
        seq, dbn = _pinned_bytes("".join(sequences)), _pinned_bytes("".join(structures))
    except UnicodeEncodeError as exc:
        raise GraphValidationError("sequence contains characters outside ACGU") from exc
    table, offset = position_table(lengths)
    skip2 = 1 if "skip2" in spec.extra_edges else 0

    lib = nat.lib
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        up = lambda a: torch.from_numpy(a).to(dev, non_blocking=True)  # noqa: E731
        seq_d, dbn_d = seq.to(dev, non_blocking=True), dbn.to(dev, non_blocking=True)
        node_ptr_d = up(node_ptr)
        table_d, offset_d = up(table), up(offset)
        edge_ptr_d = torch.empty(B + 1, dtype=torch.int64, device=dev)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        need = lib.gfx_graph_workspace_bytes(N, B)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        nat.check(lib.gfx_graph_count(dbn_d.data_ptr(), node_ptr_d.data_ptr(), B, N, skip2,
                                      edge_ptr_d.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                      need, stream))
        edge_ptr = edge_ptr_d.cpu().numpy()                 # one small read-back: E and limits
        E = int(edge_ptr[-1])
        feats = torch.empty((N, 7), dtype=torch.float32, device=dev)
        edge_index = torch.empty((2, E), dtype=torch.int32, device=dev)
        edge_types = torch.empty(E, dtype=torch.uint8, device=dev)
        residue = roles = None
        if with_node_metadata:
            residue = torch.empty(N, dtype=torch.int32, device=dev)
            roles = torch.empty(N, dtype=torch.uint8, device=dev)
        nat.check(lib.gfx_graph_fill(
            seq_d.data_ptr(), dbn_d.data_ptr(), node_ptr_d.data_ptr(), edge_ptr_d.data_ptr(),
            B, N, E, skip2, table_d.data_ptr(), offset_d.data_ptr(), feats.data_ptr(),
            edge_index.data_ptr() if E else None, edge_types.data_ptr() if E else None,
            None if residue is None else residue.data_ptr(),
            None if roles is None else roles.data_ptr(), status.data_ptr(), ws.data_ptr(), need,
            stream))
        flags = int(status.item())
    if flags & 1:
        raise GraphValidationError("sequence contains characters outside ACGU")
    if flags & 6:
        raise GraphValidationError("unbalanced or malformed dot-bracket structure")
    shard = DeviceShard(feats, edge_index, edge_types, node_ptr_d, edge_ptr_d, None,
                        max_nodes_per_record=int(lengths.max()),
                        max_edges_per_record=int(np.diff(edge_ptr).max()),
                        core_count=N, spec=spec, core_ptr_host=node_ptr)
    shard.residue_index, shard.node_roles_full = residue, roles
    shard.edge_ptr_host = edge_ptr
    shard.edges_checked = True          # built here: every edge stays inside its record
    return shard


def _build_sliced(records: list, device, spec: GraphSpec, keep_paired: bool,
                  hops: int) -> DeviceShard:
    """Windowed records: selection and induced subgraphs on the device."""
    dev = torch.device(device)
    sequences = [r.sequence for r in records]
    structures = [r.structure for r in records]
    B = len(records)
    lengths = np.fromiter(map(len, sequences), np.int64, B)
    if np.any(lengths < 1) or list(map(len, structures)) != lengths.tolist():
        raise GraphValidationError("sequence and structure lengths must match and be positive")
    start = np.zeros(B, np.int32)
    end = lengths.astype(np.int32)
    for k, r in enumerate(records):
        if getattr(r, "sliced", False):
            if r.start is None or r.end is None:
                raise GraphValidationError("sliced graph is missing start/end")
            start[k], end[k] = r.start, r.end
    if np.any(start < 0) or np.any(end > lengths) or np.any(start >= end):
        raise GraphValidationError("slice window outside the molecule")
    full_ptr = np.zeros(B + 1, np.int64)
    np.cumsum(lengths, out=full_ptr[1:])
    NF = int(full_ptr[-1])
    if NF > 1 << 30:
        raise GraphValidationError("shard exceeds 2^30 nucleotides")
    try:
        seq, dbn = _pinned_bytes("".join(sequences)), _pinned_bytes("".join(structures))
    except UnicodeEncodeError as exc:
        raise GraphValidationError("sequence contains characters outside ACGU") from exc
    table, offset = position_table(lengths)
    skip2 = 1 if "skip2" in spec.extra_edges else 0
    lib = nat.lib
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        up = lambda a: torch.from_numpy(a).to(dev, non_blocking=True)  # noqa: E731
        seq_d, dbn_d = seq.to(dev, non_blocking=True), dbn.to(dev, non_blocking=True)
        full_ptr_d, start_d, end_d = up(full_ptr), up(start), up(end)
        table_d, offset_d = up(table), up(offset)
        ptrs = torch.empty((2, B + 1), dtype=torch.int64, device=dev)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        need = lib.gfx_slice_workspace_bytes(NF, B)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        nat.check(lib.gfx_slice_select(
            dbn_d.data_ptr(), full_ptr_d.data_ptr(), start_d.data_ptr(), end_d.data_ptr(), B, NF,
            skip2, 1 if keep_paired else 0, hops, ptrs[0].data_ptr(), ptrs[1].data_ptr(),
            status.data_ptr(), ws.data_ptr(), need, stream))
        ptrs_h = ptrs.cpu().numpy()                         # one read-back: N, E and the limits
        node_ptr, edge_ptr = ptrs_h[0], ptrs_h[1]
        N, E = int(node_ptr[-1]), int(edge_ptr[-1])
        feats = torch.empty((N, 7), dtype=torch.float32, device=dev)
        edge_index = torch.empty((2, E), dtype=torch.int32, device=dev)
        edge_types = torch.empty(E, dtype=torch.uint8, device=dev)
        residue = torch.empty(N, dtype=torch.int32, device=dev)
        roles = torch.empty(N, dtype=torch.uint8, device=dev)
        nat.check(lib.gfx_slice_fill(
            seq_d.data_ptr(), dbn_d.data_ptr(), full_ptr_d.data_ptr(), start_d.data_ptr(),
            end_d.data_ptr(), ptrs[0].data_ptr(), ptrs[1].data_ptr(), B, NF, N, E, skip2,
            table_d.data_ptr(), offset_d.data_ptr(), feats.data_ptr(),
            edge_index.data_ptr() if E else None, edge_types.data_ptr() if E else None,
            residue.data_ptr(), roles.data_ptr(), status.data_ptr(), ws.data_ptr(), need, stream))
        flags = int(status.item())
    if flags & 1:
        raise GraphValidationError("sequence contains characters outside ACGU")
    if flags & 6:
        raise GraphValidationError("unbalanced or malformed dot-bracket structure")
    if flags & 8:
        raise GraphValidationError("slice window outside the molecule")
    core_ptr = np.zeros(B + 1, np.int64)
    np.cumsum((end - start).astype(np.int64), out=core_ptr[1:])
    all_core = int(core_ptr[-1]) == N
    shard = DeviceShard(feats, edge_index, edge_types, ptrs[0], ptrs[1],
                        None if all_core else roles,
                        max_nodes_per_record=int(np.diff(node_ptr).max()),
                        max_edges_per_record=int(np.diff(edge_ptr).max()),
                        core_count=int(core_ptr[-1]), spec=spec, core_ptr_host=core_ptr)
    shard.residue_index, shard.node_roles_full = residue, roles
    shard.node_ptr_host, shard.edge_ptr_host = node_ptr, edge_ptr
    shard.edges_checked = True
    return shard
