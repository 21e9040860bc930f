"""Graph data plane: GraphSpec, Graph, GraphShard, GraphBuilder, shard files.

Host-side mirror of the reference graph module (reference:
src/ginfinity/graph.py).  The array contract is byte-identical to the
reference's (docs/GRAPH_PIPELINE.md:30-38):

    node_features f32 [N,7]   edge_index i32 [2,E]   edge_types u8 [E]
    node_ptr i64 [B+1]        edge_ptr i64 [B+1]
    residue_index i32 [N]     node_roles u8 [N]

and so is the graph-spec fingerprint (graph.py:46-49,87-88).  What differs
is how the arrays are produced: the reference walks every nucleotide and
every record in Python; here a whole list of records is built with a
handful of NumPy passes over the concatenated characters (bracket matching
by a stable sort on nesting depth), and shard validation is vectorised, so
that a 100k-record shard is checked in milliseconds rather than seconds.
The integer arrays and feature columns 0-4 are bit-exact by construction;
columns 5-6 are produced by the same float32 NumPy expression as the
reference (graph.py:510-514) because NumPy's float32 sin/cos is not
correctly rounded and only the same call reproduces it.
"""
from __future__ import annotations

import hashlib
import json
import os
import tempfile
from dataclasses import dataclass, field
from pathlib import Path
from typing import Iterable, Iterator, Mapping, Optional, Sequence

import numpy as np

from .records import RNA

GRAPH_SHARD_FORMAT = "ginfinity-graph-shard"
GRAPH_SHARD_FORMAT_VERSION = 1

# reference: graph.py:20-27
EDGE_TYPE_CODES = {
    "backbone_forward": 0, "backbone_reverse": 1,
    "base_pair_forward": 2, "base_pair_reverse": 3,
    "skip2_forward": 4, "skip2_reverse": 5,
}
NODE_ROLE_CORE = np.uint8(0)
NODE_ROLE_CONTEXT = np.uint8(1)

_REQUIRED_TENSORS = frozenset(
    ("node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr"))
_OPTIONAL_TENSORS = frozenset(("residue_index", "node_roles"))


class GraphValidationError(ValueError):
    """A graph or graph shard violates the public interchange contract."""


class GraphCompatibilityError(GraphValidationError):
    """A graph was built with a specification incompatible with the encoder."""


def _file_sha256(path: Path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        while True:
            block = fh.read(1 << 20)
            if not block:
                break
            h.update(block)
    return h.hexdigest()


# --------------------------------------------------------------------------
# GraphSpec
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class GraphSpec:
    """Model-versioned feature/edge contract (reference: graph.py:60-161)."""

    format_version: int = 1
    struct_feature: str = "A"
    positional: bool = True
    edge_dim: int = 10
    extra_edges: tuple = ("skip2",)
    _fingerprint: str = field(default="", init=False, repr=False,
                              compare=False)

    def __post_init__(self):
        object.__setattr__(self, "extra_edges", tuple(self.extra_edges))
        if self.format_version != GRAPH_SHARD_FORMAT_VERSION:
            raise GraphValidationError(
                f"unsupported graph specification version {self.format_version}")
        if self.struct_feature not in ("A", "B"):
            raise GraphValidationError(
                f"unsupported structure feature {self.struct_feature!r}")
        unknown = sorted(set(self.extra_edges) - {"skip2"})
        if unknown:
            raise GraphValidationError(
                "unsupported extra edge type(s): " + ", ".join(unknown))
        if self.edge_dim < (6 if "skip2" in self.extra_edges else 4):
            raise GraphValidationError(
                f"edge_dim={self.edge_dim} cannot represent all configured edges")
        # canonical JSON: sorted keys, no whitespace, ASCII (graph.py:46-49)
        blob = json.dumps(self.to_dict(), sort_keys=True,
                          separators=(",", ":"), ensure_ascii=True)
        object.__setattr__(self, "_fingerprint",
                           hashlib.sha256(blob.encode("utf-8")).hexdigest())

    @property
    def node_feature_dim(self) -> int:
        return 4 + (1 if self.struct_feature == "A" else 3) + (
            2 if self.positional else 0)

    @property
    def edge_types(self) -> dict:
        names = ["backbone_forward", "backbone_reverse",
                 "base_pair_forward", "base_pair_reverse"]
        if "skip2" in self.extra_edges:
            names += ["skip2_forward", "skip2_reverse"]
        return {n: EDGE_TYPE_CODES[n] for n in names}

    def to_dict(self) -> dict:
        return {
            "format_version": self.format_version,
            "struct_feature": self.struct_feature,
            "positional": self.positional,
            "node_feature_dimension": self.node_feature_dim,
            "edge_feature_dimension": self.edge_dim,
            "edge_types": self.edge_types,
            "extra_edges": list(self.extra_edges),
        }

    @property
    def sha256(self) -> str:
        return self._fingerprint

    @classmethod
    def from_dict(cls, value: Mapping) -> "GraphSpec":
        spec = cls(
            format_version=int(value.get("format_version", 1)),
            struct_feature=str(value["struct_feature"]),
            positional=bool(value["positional"]),
            edge_dim=int(value.get("edge_feature_dimension",
                                   value.get("edge_dim", 10))),
            extra_edges=tuple(value.get("extra_edges", ())))
        if ("node_feature_dimension" in value and
                int(value["node_feature_dimension"]) != spec.node_feature_dim):
            raise GraphValidationError("node feature dimension is inconsistent")
        if "edge_types" in value and dict(value["edge_types"]) != spec.edge_types:
            raise GraphValidationError("edge type mapping is inconsistent")
        return spec

    @classmethod
    def from_encoder_config(cls, value) -> "GraphSpec":
        get = (value.__getitem__ if isinstance(value, Mapping)
               else lambda name: getattr(value, name))
        return cls(struct_feature=str(get("struct_feature")),
                   positional=bool(get("positional")),
                   edge_dim=int(get("edge_dim")),
                   extra_edges=tuple(get("extra_edges")))

    @classmethod
    def bundled(cls) -> "GraphSpec":
        """The contract of the bundled model.

        The reference reads it from its packaged model.json
        (graph.py:148-161).  The contract of model 1.0.0 is fixed, and the
        fingerprint below is the one recorded at
        src/ginfinity/data/model.json:42; it is checked, not trusted.
        """
        spec = cls()
        if spec.sha256 != BUNDLED_GRAPH_SPEC_SHA256:
            raise GraphValidationError(
                "bundled graph specification fingerprint mismatch")
        return spec


BUNDLED_GRAPH_SPEC_SHA256 = (
    "da2e670e377e47667fec8a8ebb1c90c6e506b9cdd8a5555a6bfab50b202fb9bd")


# --------------------------------------------------------------------------
# Graph
# --------------------------------------------------------------------------
def _check_roles(node_roles: np.ndarray) -> None:
    if node_roles.size and int(node_roles.max()) > int(NODE_ROLE_CONTEXT):
        raise GraphValidationError("unknown node role")


@dataclass(frozen=True)
class Graph:
    """One RNA graph with local, zero-based edge indices
    (reference: graph.py:164-258)."""

    identifier: str
    sequence: str
    structure: str
    node_features: np.ndarray
    edge_index: np.ndarray
    edge_types: np.ndarray
    spec: GraphSpec
    residue_index: np.ndarray
    node_roles: np.ndarray

    def __post_init__(self):
        meta_ok = all(isinstance(v, str) for v in
                      (self.identifier, self.sequence, self.structure))
        if (not meta_ok or not self.identifier or not self.sequence
                or len(self.structure) != len(self.sequence)):
            raise GraphValidationError("invalid graph record metadata")
        ri, roles = self.residue_index, self.node_roles
        if ri.dtype != np.int32 or ri.ndim != 1 or ri.size == 0:
            raise GraphValidationError(
                "residue_index must be a non-empty int32 vector")
        n = int(ri.shape[0])
        if roles.dtype != np.uint8 or roles.shape != (n,):
            raise GraphValidationError("node_roles must match residue_index")
        if (self.node_features.dtype != np.float32 or
                self.node_features.shape != (n, self.spec.node_feature_dim)):
            raise GraphValidationError("invalid node feature array")
        ei, et = self.edge_index, self.edge_types
        if ei.dtype != np.int32 or ei.ndim != 2 or ei.shape[0] != 2:
            raise GraphValidationError(
                "edge_index must have shape (2, E) and int32 dtype")
        if et.dtype != np.uint8 or et.ndim != 1 or et.shape[0] != ei.shape[1]:
            raise GraphValidationError(
                "edge_types must have shape (E,) and uint8 dtype")
        if ei.size and (int(ei.min()) < 0 or int(ei.max()) >= n):
            raise GraphValidationError("edge index outside graph node range")
        if et.size and int(et.max()) >= self.spec.edge_dim:
            raise GraphValidationError("edge type outside graph feature range")
        if int(ri.min()) < 0 or int(ri.max()) >= len(self.sequence):
            raise GraphValidationError("residue index outside source sequence")
        if n > 1 and not bool(np.all(ri[1:] > ri[:-1])):
            raise GraphValidationError(
                "residue_index must be strictly increasing")
        _check_roles(roles)
        if not bool(np.any(roles == NODE_ROLE_CORE)):
            raise GraphValidationError("graph has no core nodes")

    @property
    def length(self) -> int:
        return len(self.sequence)

    @property
    def node_count(self) -> int:
        return int(self.node_features.shape[0])

    @property
    def edge_count(self) -> int:
        return int(self.edge_index.shape[1])

    @property
    def core_mask(self) -> np.ndarray:
        return self.node_roles == NODE_ROLE_CORE

    @property
    def core_count(self) -> int:
        return int(np.count_nonzero(self.core_mask))

    @property
    def core_positions(self) -> np.ndarray:
        return self.residue_index[self.core_mask]

    @property
    def core_span(self) -> tuple:
        core = self.core_positions
        return int(core[0]), int(core[-1]) + 1


# --------------------------------------------------------------------------
# GraphShard
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class GraphShard:
    """A persistent scheduling unit of one or more graphs
    (reference: graph.py:261-457).  Validation is vectorised."""

    identifiers: tuple
    sequences: tuple
    structures: tuple
    node_features: np.ndarray
    edge_index: np.ndarray
    edge_types: np.ndarray
    node_ptr: np.ndarray
    edge_ptr: np.ndarray
    spec: GraphSpec
    residue_index: np.ndarray
    node_roles: np.ndarray

    def __post_init__(self):
        self._validate()

    def _validate(self) -> None:
        b = len(self.identifiers)
        if b == 0:
            raise GraphValidationError("a graph shard cannot be empty")
        if (not all(isinstance(v, str) and v for v in self.identifiers)
                or not all(isinstance(v, str) and v for v in self.sequences)
                or not all(isinstance(v, str) for v in self.structures)):
            raise GraphValidationError("invalid graph shard record metadata")
        if len(set(self.identifiers)) != b:
            raise GraphValidationError("duplicate identifiers in graph shard")
        if len(self.sequences) != b or len(self.structures) != b:
            raise GraphValidationError("graph shard metadata count mismatch")
        nptr, eptr = self.node_ptr, self.edge_ptr
        if nptr.dtype != np.int64 or nptr.shape != (b + 1,):
            raise GraphValidationError(
                "node_ptr must have shape (B + 1,) and int64 dtype")
        if eptr.dtype != np.int64 or eptr.shape != (b + 1,):
            raise GraphValidationError(
                "edge_ptr must have shape (B + 1,) and int64 dtype")
        if (nptr[0] != 0 or eptr[0] != 0 or np.any(nptr[1:] <= nptr[:-1])
                or np.any(eptr[1:] < eptr[:-1])):
            raise GraphValidationError("invalid graph shard offsets")
        n, e = int(nptr[-1]), int(eptr[-1])
        if (self.node_features.dtype != np.float32 or
                self.node_features.shape != (n, self.spec.node_feature_dim)):
            raise GraphValidationError("invalid shard node feature array")
        if self.edge_index.dtype != np.int32 or self.edge_index.shape != (2, e):
            raise GraphValidationError("invalid shard edge index array")
        if self.edge_types.dtype != np.uint8 or self.edge_types.shape != (e,):
            raise GraphValidationError("invalid shard edge type array")
        if (self.residue_index.dtype != np.int32
                or self.residue_index.shape != (n,)):
            raise GraphValidationError("invalid shard residue_index array")
        if self.node_roles.dtype != np.uint8 or self.node_roles.shape != (n,):
            raise GraphValidationError("invalid shard node_roles array")
        if e and (int(self.edge_index.min()) < 0
                  or int(self.edge_index.max()) >= n):
            raise GraphValidationError("edge index outside shard node range")
        if e and int(self.edge_types.max()) >= self.spec.edge_dim:
            raise GraphValidationError("edge type outside shard feature range")
        seq_len = np.fromiter((len(s) for s in self.sequences), np.int64, b)
        dbn_len = np.fromiter((len(s) for s in self.structures), np.int64, b)
        if not np.array_equal(seq_len, dbn_len):
            raise GraphValidationError(
                "sequence/structure length mismatch in shard")
        _check_roles(self.node_roles)
        # per-record checks of graph.py:330-343, all records at once
        counts = np.diff(nptr)
        ri = self.residue_index
        limit = np.repeat(seq_len, counts)
        if bool(np.any(ri < 0)) or bool(np.any(ri >= limit)):
            raise GraphValidationError("residue index outside source sequence")
        if n > 1:
            rising = ri[1:] > ri[:-1]
            rising[nptr[1:-1] - 1] = True      # record boundaries are exempt
            if not bool(rising.all()):
                raise GraphValidationError(
                    "residue_index must be strictly increasing")
        core = np.add.reduceat(
            (self.node_roles == NODE_ROLE_CORE).astype(np.int64), nptr[:-1])
        if bool(np.any(core == 0)):
            raise GraphValidationError("graph has no core nodes")

    # -- sizes ------------------------------------------------------------
    @property
    def record_count(self) -> int:
        return len(self.identifiers)

    @property
    def node_count(self) -> int:
        return int(self.node_ptr[-1])

    @property
    def edge_count(self) -> int:
        return int(self.edge_ptr[-1])

    @property
    def lengths(self) -> tuple:
        return tuple(np.diff(self.node_ptr).tolist())

    @property
    def edge_counts(self) -> tuple:
        return tuple(np.diff(self.edge_ptr).tolist())

    @property
    def core_counts(self) -> tuple:
        return tuple(self.core_count_array().tolist())

    def core_count_array(self) -> np.ndarray:
        return np.add.reduceat(
            (self.node_roles == NODE_ROLE_CORE).astype(np.int64),
            self.node_ptr[:-1])

    @property
    def all_core(self) -> bool:
        return not bool(self.node_roles.any())

    # -- construction -----------------------------------------------------
    @classmethod
    def from_graphs(cls, graphs: Sequence[Graph]) -> "GraphShard":
        graphs = list(graphs)
        if not graphs:
            raise GraphValidationError("cannot create a shard without graphs")
        spec = graphs[0].spec
        if any(g.spec.sha256 != spec.sha256 for g in graphs):
            raise GraphCompatibilityError(
                "all graphs in a shard must use the same graph specification")
        b = len(graphs)
        node_ptr = np.zeros(b + 1, np.int64)
        edge_ptr = np.zeros(b + 1, np.int64)
        np.cumsum([g.node_count for g in graphs], out=node_ptr[1:])
        np.cumsum([g.edge_count for g in graphs], out=edge_ptr[1:])
        if int(node_ptr[-1]) > np.iinfo(np.int32).max:
            raise GraphValidationError(
                "graph shard exceeds the int32 node-index capacity; split it")
        edge_index = np.concatenate([g.edge_index for g in graphs], axis=1)
        edge_index = np.ascontiguousarray(edge_index, dtype=np.int32)
        edge_index += np.repeat(node_ptr[:-1], np.diff(edge_ptr)).astype(
            np.int32)[None, :]
        cat = lambda name: np.ascontiguousarray(  # noqa: E731
            np.concatenate([getattr(g, name) for g in graphs], axis=0))
        return cls(identifiers=tuple(g.identifier for g in graphs),
                   sequences=tuple(g.sequence for g in graphs),
                   structures=tuple(g.structure for g in graphs),
                   node_features=cat("node_features"),
                   edge_index=edge_index, edge_types=cat("edge_types"),
                   node_ptr=node_ptr, edge_ptr=edge_ptr, spec=spec,
                   residue_index=cat("residue_index"),
                   node_roles=cat("node_roles"))

    def slice(self, start: int, stop: int) -> "GraphShard":
        """Contiguous record range with rebased indices (graph.py:414-444)."""
        if not 0 <= start < stop <= self.record_count:
            raise IndexError("invalid graph shard slice")
        n0, n1 = int(self.node_ptr[start]), int(self.node_ptr[stop])
        e0, e1 = int(self.edge_ptr[start]), int(self.edge_ptr[stop])
        return GraphShard(
            identifiers=self.identifiers[start:stop],
            sequences=self.sequences[start:stop],
            structures=self.structures[start:stop],
            node_features=np.ascontiguousarray(self.node_features[n0:n1]),
            edge_index=np.ascontiguousarray(
                self.edge_index[:, e0:e1] - np.int32(n0), dtype=np.int32),
            edge_types=np.ascontiguousarray(self.edge_types[e0:e1]),
            node_ptr=np.ascontiguousarray(
                self.node_ptr[start:stop + 1] - n0, dtype=np.int64),
            edge_ptr=np.ascontiguousarray(
                self.edge_ptr[start:stop + 1] - e0, dtype=np.int64),
            spec=self.spec,
            residue_index=np.ascontiguousarray(self.residue_index[n0:n1]),
            node_roles=np.ascontiguousarray(self.node_roles[n0:n1]))

    def validate_values(self) -> None:
        """Finite features; no edge leaves its graph (graph.py:446-457)."""
        if not np.isfinite(self.node_features).all():
            raise GraphValidationError("non-finite node features in graph shard")
        if self.edge_count == 0:
            return
        per_edge = np.repeat(np.arange(self.record_count), np.diff(self.edge_ptr))
        lo = self.node_ptr[:-1][per_edge]
        hi = self.node_ptr[1:][per_edge]
        ei = self.edge_index
        if bool(np.any(ei < lo[None, :])) or bool(np.any(ei >= hi[None, :])):
            raise GraphValidationError("edge crosses graph boundaries")


# --------------------------------------------------------------------------
# Vectorised construction
# --------------------------------------------------------------------------
_BASE_COLUMN = np.full(256, -1, np.int64)
for _i, _c in enumerate("ACGU"):
    _BASE_COLUMN[ord(_c)] = _i
_OPEN, _CLOSE, _DOT = ord("("), ord(")"), ord(".")


def pair_table_flat(dbn: np.ndarray, node_ptr: np.ndarray) -> np.ndarray:
    """Partner index (local to its record) or -1 for every character of the
    concatenated dot-bracket strings ``dbn`` (uint8).

    Same result as the reference's stack matcher (graph.py:737-747).  At a
    fixed nesting level brackets alternate open/close, so a stable sort of
    bracket positions on (record, level) puts partners next to each other.
    """
    n = dbn.shape[0]
    partners = np.full(n, -1, np.int32)
    step = (dbn == _OPEN).astype(np.int64) - (dbn == _CLOSE).astype(np.int64)
    if not step.any():
        return partners
    depth = np.cumsum(step)
    b = node_ptr.shape[0] - 1
    counts = np.diff(node_ptr)
    record = np.repeat(np.arange(b, dtype=np.int64), counts)
    # depth carried in from earlier records is 0 for balanced input, but
    # subtract it anyway so an unbalanced record cannot shift its successors
    carry = np.concatenate(([0], depth[node_ptr[1:-1] - 1]))
    depth = depth - carry[record]
    pos = np.flatnonzero(step)
    level = depth[pos] + (step[pos] < 0)        # '(' : after +1, ')' : before -1
    key = record[pos] * (int(level.max()) + 2) + level
    order = np.argsort(key, kind="stable")
    ordered = pos[order]
    if ordered.shape[0] % 2:
        raise GraphValidationError("unbalanced structure")
    opens, closes = ordered[0::2], ordered[1::2]
    if (np.any(dbn[opens] != _OPEN) or np.any(dbn[closes] != _CLOSE)
            or np.any(record[opens] != record[closes])):
        raise GraphValidationError("unbalanced structure")
    base = node_ptr[:-1][record[opens]]
    partners[opens] = (closes - base).astype(np.int32)
    partners[closes] = (opens - base).astype(np.int32)
    return partners


def _pair_table(structure: str) -> np.ndarray:
    dbn = np.frombuffer(structure.encode("ascii"), np.uint8)
    return pair_table_flat(dbn, np.array([0, dbn.shape[0]], np.int64))


def _full_arrays(sequences: Sequence[str], structures: Sequence[str],
                 spec: GraphSpec):
    """Concatenated full-molecule arrays for many records at once.

    Edge order per record is the reference's (graph.py:520-542): backbone
    i->i+1, backbone i+1->i, pairs open->close by ascending opening,
    close->open, then skip-2 interleaved (i->i+2, i+2->i).  Edge indices
    returned here are LOCAL to each record.
    """
    b = len(sequences)
    lengths = np.fromiter((len(s) for s in sequences), np.int64, b)
    node_ptr = np.zeros(b + 1, np.int64)
    np.cumsum(lengths, out=node_ptr[1:])
    n = int(node_ptr[-1])
    seq = np.frombuffer("".join(sequences).encode("ascii"), np.uint8)
    dbn = np.frombuffer("".join(structures).encode("ascii"), np.uint8)
    record = np.repeat(np.arange(b, dtype=np.int64), lengths)
    local = np.arange(n, dtype=np.int64) - node_ptr[:-1][record]

    feats = np.zeros((n, spec.node_feature_dim), np.float32)
    feats[np.arange(n), _BASE_COLUMN[seq]] = 1.0
    col = 4
    if spec.struct_feature == "A":
        feats[:, 4] = dbn != _DOT
        col = 5
    else:
        state = np.where(dbn == _OPEN, 0, np.where(dbn == _DOT, 1, 2))
        feats[np.arange(n), 4 + state] = 1.0
        col = 7
    if spec.positional:
        # identical float32 expression to graph.py:511-514
        denom = np.maximum(lengths - 1, 1)[record]
        relative = local.astype(np.float32) / denom.astype(np.float32)
        angle = np.float32(np.pi) * relative
        feats[:, col] = np.sin(angle)
        feats[:, col + 1] = np.cos(angle)

    partners = pair_table_flat(dbn, node_ptr)
    skip = "skip2" in spec.extra_edges
    n_bb = np.maximum(lengths - 1, 0)
    n_sk = np.maximum(lengths - 2, 0) if skip else np.zeros(b, np.int64)
    is_open = partners > local
    n_pair = np.add.reduceat(is_open.astype(np.int64), node_ptr[:-1]) \
        if n else np.zeros(b, np.int64)
    e_per = 2 * n_bb + 2 * n_pair + 2 * n_sk
    edge_ptr = np.zeros(b + 1, np.int64)
    np.cumsum(e_per, out=edge_ptr[1:])
    e = int(edge_ptr[-1])
    src = np.empty(e, np.int32)
    dst = np.empty(e, np.int32)
    typ = np.empty(e, np.uint8)

    def seg(counts):
        """(record id, index within record) for `counts[r]` items per record."""
        rec = np.repeat(np.arange(b, dtype=np.int64), counts)
        start = np.cumsum(counts) - counts
        return rec, np.arange(int(counts.sum()), dtype=np.int64) - start[rec]

    # backbone
    rec, i = seg(n_bb)
    at = edge_ptr[:-1][rec] + i
    src[at], dst[at], typ[at] = i, i + 1, 0
    at = at + n_bb[rec]
    src[at], dst[at], typ[at] = i + 1, i, 1
    # base pairs (openings are already in ascending order per record)
    op = np.flatnonzero(is_open)
    rec = record[op]
    k = np.arange(op.shape[0], dtype=np.int64) - (np.cumsum(n_pair) - n_pair)[rec]
    at = edge_ptr[:-1][rec] + 2 * n_bb[rec] + k
    src[at], dst[at], typ[at] = local[op], partners[op], 2
    at = at + n_pair[rec]
    src[at], dst[at], typ[at] = partners[op], local[op], 3
    # skip-2, interleaved
    if skip:
        rec, i = seg(n_sk)
        at = edge_ptr[:-1][rec] + 2 * n_bb[rec] + 2 * n_pair[rec] + 2 * i
        src[at], dst[at], typ[at] = i, i + 2, 4
        src[at + 1], dst[at + 1], typ[at + 1] = i + 2, i, 5
    return feats, src, dst, typ, node_ptr, edge_ptr, local.astype(np.int32)


def _select_slice_nodes(graph: Graph, start: int, end: int, *,
                        keep_paired_neighbours: bool, context_hops: int):
    """Core window + (optionally) crossing-pair partners + BFS context.

    Reference: graph.py:608-646.  Hop 1 adds partners of core nucleotides;
    hops 2.. expand only from the nodes added by the previous hop, along
    every outgoing edge.
    """
    n = graph.node_count
    chosen = np.zeros(n, bool)
    chosen[start:end] = True
    if keep_paired_neighbours:
        partners = _pair_table(graph.structure)
        mates = partners[start:end]
        mates = mates[mates >= 0]
        frontier = np.unique(mates[~chosen[mates]])
        chosen[frontier] = True
        if context_hops > 1 and frontier.size and graph.edge_count:
            order = np.argsort(graph.edge_index[0], kind="stable")
            nbr = graph.edge_index[1][order]
            offs = np.zeros(n + 1, np.int64)
            np.cumsum(np.bincount(graph.edge_index[0], minlength=n),
                      out=offs[1:])
            for _ in range(context_hops - 1):
                if not frontier.size:
                    break
                lo, hi = offs[frontier], offs[frontier + 1]
                take = np.concatenate(
                    [nbr[a:b] for a, b in zip(lo.tolist(), hi.tolist())]
                ) if frontier.size else np.zeros(0, np.int32)
                fresh = np.unique(take[~chosen[take]])
                chosen[fresh] = True
                frontier = fresh
    residue_index = np.flatnonzero(chosen).astype(np.int32)
    roles = np.where((residue_index >= start) & (residue_index < end),
                     NODE_ROLE_CORE, NODE_ROLE_CONTEXT).astype(np.uint8)
    return residue_index, roles


def _extract_slice(graph: Graph, start, end, *, keep_paired_neighbours: bool,
                   context_hops: int) -> Graph:
    """Induced subgraph on the selected nodes, original edge order kept
    (reference: graph.py:649-695)."""
    if start is None or end is None:
        raise GraphValidationError("sliced graph is missing start/end")
    residue_index, roles = _select_slice_nodes(
        graph, start, end, keep_paired_neighbours=keep_paired_neighbours,
        context_hops=context_hops)
    remap = np.full(graph.node_count, -1, np.int32)
    remap[residue_index] = np.arange(residue_index.shape[0], dtype=np.int32)
    if graph.edge_count:
        mapped = remap[graph.edge_index]
        keep = (mapped[0] >= 0) & (mapped[1] >= 0)
        edge_index = np.ascontiguousarray(mapped[:, keep])
        edge_types = np.ascontiguousarray(graph.edge_types[keep])
    else:
        edge_index = np.zeros((2, 0), np.int32)
        edge_types = np.zeros((0,), np.uint8)
    return Graph(identifier=graph.identifier, sequence=graph.sequence,
                 structure=graph.structure,
                 node_features=np.ascontiguousarray(
                     graph.node_features[residue_index]),
                 edge_index=edge_index, edge_types=edge_types,
                 spec=graph.spec, residue_index=residue_index,
                 node_roles=roles)


class GraphBuilder:
    """Deterministic RNA -> graph conversion (reference: graph.py:460-567)."""

    def __init__(self, spec: Optional[GraphSpec] = None, *,
                 keep_paired_neighbours: bool = False, context_hops: int = 1):
        if context_hops < 1:
            raise ValueError("context_hops must be >= 1")
        self.spec = spec if spec is not None else GraphSpec.bundled()
        self.keep_paired_neighbours = bool(keep_paired_neighbours)
        self.context_hops = int(context_hops)

    def _build_full(self, record) -> Graph:
        feats, src, dst, typ, _, _, local = _full_arrays(
            [record.sequence], [record.structure], self.spec)
        return Graph(identifier=record.identifier, sequence=record.sequence,
                     structure=record.structure, node_features=feats,
                     edge_index=np.ascontiguousarray(np.stack((src, dst))),
                     edge_types=typ, spec=self.spec, residue_index=local,
                     node_roles=np.zeros(local.shape[0], np.uint8))

    def build(self, record) -> Graph:
        graph = self._build_full(record)
        if not getattr(record, "sliced", False):
            return graph
        return _extract_slice(
            graph, record.start, record.end,
            keep_paired_neighbours=self.keep_paired_neighbours,
            context_hops=self.context_hops)

    def build_many(self, records: Iterable) -> list:
        return [self.build(r) for r in records]

    def build_shard(self, records: Iterable) -> GraphShard:
        """Same arrays as ``GraphShard.from_graphs(self.build_many(records))``
        (graph.py:566-567); full-molecule records take a batched path that
        never creates per-record ``Graph`` objects."""
        records = list(records)
        if not records:
            raise GraphValidationError("cannot create a shard without graphs")
        if any(getattr(r, "sliced", False) for r in records):
            return GraphShard.from_graphs(self.build_many(records))
        seqs = tuple(r.sequence for r in records)
        dbns = tuple(r.structure for r in records)
        feats, src, dst, typ, node_ptr, edge_ptr, local = _full_arrays(
            seqs, dbns, self.spec)
        if int(node_ptr[-1]) > np.iinfo(np.int32).max:
            raise GraphValidationError(
                "graph shard exceeds the int32 node-index capacity; split it")
        shift = np.repeat(node_ptr[:-1], np.diff(edge_ptr)).astype(np.int32)
        edge_index = np.empty((2, src.shape[0]), np.int32)
        np.add(src, shift, out=edge_index[0])
        np.add(dst, shift, out=edge_index[1])
        return GraphShard(
            identifiers=tuple(r.identifier for r in records),
            sequences=seqs, structures=dbns, node_features=feats,
            edge_index=edge_index, edge_types=typ, node_ptr=node_ptr,
            edge_ptr=edge_ptr, spec=self.spec, residue_index=local,
            node_roles=np.zeros(local.shape[0], np.uint8))


def partition_records(records: Iterable, *, max_records: int,
                      max_nodes: Optional[int] = None) -> Iterator[tuple]:
    """Deterministic scheduling units (reference: graph.py:570-596)."""
    if max_records <= 0:
        raise ValueError("max_records must be positive")
    if max_nodes is not None and max_nodes <= 0:
        raise ValueError("max_nodes must be positive")
    group: list = []
    total = 0
    for record in records:
        size = record.length
        if max_nodes is not None and size > max_nodes:
            raise ValueError(
                f"record {record.identifier!r} exceeds max_nodes={max_nodes}")
        full = len(group) >= max_records or (
            max_nodes is not None and total + size > max_nodes)
        if group and full:
            yield tuple(group)
            group, total = [], 0
        group.append(record)
        total += size
    if group:
        yield tuple(group)


# --------------------------------------------------------------------------
# Shard files (reference: graph.py:750-923)
# --------------------------------------------------------------------------
def graph_metadata_path(tensor_path) -> Path:
    return Path(tensor_path).with_suffix(".json")


def _trivial_node_metadata(shard: GraphShard) -> bool:
    """True when roles/residues are exactly what full molecules imply, so the
    optional tensors may be omitted (reference: graph.py:698-714)."""
    if shard.node_roles.any():
        return False
    seq_len = np.fromiter((len(s) for s in shard.sequences), np.int64,
                          shard.record_count)
    counts = np.diff(shard.node_ptr)
    if not np.array_equal(counts, seq_len):
        return False
    expect = (np.arange(shard.node_count, dtype=np.int64)
              - np.repeat(shard.node_ptr[:-1], counts))
    return bool(np.array_equal(shard.residue_index, expect))


def _atomic_write(target: Path, writer) -> None:
    fd, name = tempfile.mkstemp(dir=target.parent, prefix=f".{target.name}.",
                                suffix=".tmp")
    os.close(fd)
    tmp = Path(name)
    try:
        writer(tmp)
        os.replace(tmp, target)
    finally:
        tmp.unlink(missing_ok=True)


def save_graph_shard(shard: GraphShard, tensor_path, *, metadata_path=None,
                     checksum: bool = False) -> tuple:
    from safetensors.numpy import save_file

    tensor_path = Path(tensor_path)
    metadata_path = (Path(metadata_path) if metadata_path is not None
                     else graph_metadata_path(tensor_path))
    if tensor_path.resolve() == metadata_path.resolve():
        raise ValueError("tensor and metadata paths must be different")
    tensor_path.parent.mkdir(parents=True, exist_ok=True)
    metadata_path.parent.mkdir(parents=True, exist_ok=True)
    tensors = {name: getattr(shard, name) for name in
               ("node_features", "edge_index", "edge_types", "node_ptr",
                "edge_ptr")}
    if not _trivial_node_metadata(shard):
        tensors["residue_index"] = shard.residue_index
        tensors["node_roles"] = shard.node_roles
    header = {"format": GRAPH_SHARD_FORMAT,
              "format_version": str(GRAPH_SHARD_FORMAT_VERSION),
              "graph_spec_sha256": shard.spec.sha256}
    _atomic_write(tensor_path,
                  lambda tmp: save_file(tensors, str(tmp), metadata=header))
    sidecar = {
        "format": GRAPH_SHARD_FORMAT,
        "format_version": GRAPH_SHARD_FORMAT_VERSION,
        "graph_spec": shard.spec.to_dict(),
        "graph_spec_sha256": shard.spec.sha256,
        "tensor_file": tensor_path.name,
        "record_count": shard.record_count,
        "node_count": shard.node_count,
        "edge_count": shard.edge_count,
        "identifiers": list(shard.identifiers),
        "sequences": list(shard.sequences),
        "structures": list(shard.structures),
    }
    if checksum:
        sidecar["tensor_sha256"] = _file_sha256(tensor_path)
    _atomic_write(metadata_path, lambda tmp: tmp.write_text(
        json.dumps(sidecar, indent=2) + "\n"))
    return tensor_path, metadata_path


_SAFETENSORS_DTYPES = {"F32": np.float32, "I32": np.int32, "U8": np.uint8, "I64": np.int64}


def _read_safetensors_into(path: Path, allocate) -> tuple:
    """(header metadata, {name: array}) of a safetensors file, every tensor copied ONCE from the
    memory-mapped file into `allocate(name, shape, dtype)` (e.g. page-locked buffers).  The format
    is an 8-byte little-endian header length, a JSON header {name: {dtype, shape, data_offsets}}
    and the raw tensor bytes; only the dtypes a graph shard uses are accepted."""
    import mmap
    with open(path, "rb") as fh:
        size = fh.seek(0, 2)
        if size < 8:
            raise ValueError("file too short for a safetensors header")
        with mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            header_len = int.from_bytes(mm[:8], "little")
            if header_len <= 0 or 8 + header_len > size:
                raise ValueError("invalid safetensors header length")
            header = json.loads(mm[8:8 + header_len].decode("utf-8"))
            base = 8 + header_len
            arrays = {}
            view = memoryview(mm)
            try:
                for name, info in header.items():
                    if name == "__metadata__":
                        continue
                    dtype = _SAFETENSORS_DTYPES.get(info["dtype"])
                    if dtype is None:
                        raise ValueError(f"unexpected dtype {info['dtype']!r} for {name}")
                    begin, end = (int(v) for v in info["data_offsets"])
                    shape = tuple(int(v) for v in info["shape"])
                    count = int(np.prod(shape, dtype=np.int64)) if shape else 1
                    if (begin < 0 or end < begin or base + end > size
                            or end - begin != count * np.dtype(dtype).itemsize):
                        raise ValueError(f"invalid data offsets for {name}")
                    source = np.frombuffer(view[base + begin:base + end], dtype=dtype).reshape(shape)
                    target = allocate(name, shape, np.dtype(dtype))
                    np.copyto(target, source)
                    del source
                    arrays[name] = target
            finally:
                view.release()
    return header.get("__metadata__") or {}, arrays


def load_graph_shard(tensor_path, *, metadata_path=None,
                     expected_spec: Optional[GraphSpec] = None,
                     verify_checksum: bool = False,
                     validation: str = "metadata", allocate=None) -> GraphShard:
    """Load and validate a graph shard (reference graph.py:826-923).

    `allocate(name, shape, dtype) -> ndarray` (extension): where each tensor is
    to live, e.g. page-locked buffers of a pool; the file is then memory-mapped
    and every tensor copied once, straight into its destination."""
    from safetensors import safe_open
    from safetensors.numpy import load_file

    tensor_path = Path(tensor_path)
    metadata_path = (Path(metadata_path) if metadata_path is not None
                     else graph_metadata_path(tensor_path))
    if validation not in ("metadata", "full"):
        raise ValueError("validation must be 'metadata' or 'full'")
    try:
        sidecar = json.loads(metadata_path.read_text())
    except (OSError, json.JSONDecodeError) as exc:
        raise GraphValidationError(
            f"cannot read graph shard metadata: {exc}") from exc
    if (sidecar.get("format") != GRAPH_SHARD_FORMAT
            or sidecar.get("format_version") != GRAPH_SHARD_FORMAT_VERSION):
        raise GraphValidationError("unsupported graph shard format")
    try:
        spec = GraphSpec.from_dict(sidecar["graph_spec"])
    except GraphValidationError:
        raise
    except (KeyError, TypeError, ValueError) as exc:
        raise GraphValidationError(
            f"invalid graph specification metadata: {exc}") from exc
    if sidecar.get("graph_spec_sha256") != spec.sha256:
        raise GraphValidationError("graph specification fingerprint mismatch")
    if expected_spec is not None and spec.sha256 != expected_spec.sha256:
        raise GraphCompatibilityError(
            "graph shard specification is incompatible with the encoder")
    if verify_checksum:
        stored = sidecar.get("tensor_sha256")
        if not stored:
            raise GraphValidationError("graph shard has no stored checksum")
        if _file_sha256(tensor_path) != stored:
            raise GraphValidationError("graph shard checksum mismatch")
    try:
        if allocate is not None:
            header, arrays = _read_safetensors_into(tensor_path, allocate)
        else:
            with safe_open(str(tensor_path), framework="np") as fh:
                header = fh.metadata() or {}
        if (header.get("format") != GRAPH_SHARD_FORMAT
                or header.get("format_version") != str(GRAPH_SHARD_FORMAT_VERSION)
                or header.get("graph_spec_sha256") != spec.sha256):
            raise GraphValidationError("tensor header metadata mismatch")
        if allocate is None:
            arrays = load_file(str(tensor_path))
    except GraphValidationError:
        raise
    except Exception as exc:
        raise GraphValidationError(
            f"cannot load graph shard tensors: {exc}") from exc
    names = set(arrays)
    if _REQUIRED_TENSORS - names:
        raise GraphValidationError("graph shard tensor set mismatch")
    extra = names - _REQUIRED_TENSORS - _OPTIONAL_TENSORS
    if extra:
        raise GraphValidationError(
            "unexpected graph shard tensor(s): " + ", ".join(sorted(extra)))
    if ("residue_index" in names) != ("node_roles" in names):
        raise GraphValidationError(
            "residue_index and node_roles must be stored together")
    try:
        sequences = tuple(sidecar["sequences"])
        node_ptr = arrays["node_ptr"]
        if "residue_index" in names:
            residue_index, node_roles = (arrays["residue_index"],
                                         arrays["node_roles"])
        else:
            # shards written before slicing existed (graph.py:717-734)
            if (node_ptr.dtype != np.int64
                    or node_ptr.shape != (len(sequences) + 1,)):
                raise GraphValidationError("invalid graph shard offsets")
            seq_len = np.fromiter((len(s) for s in sequences), np.int64,
                                  len(sequences))
            if not np.array_equal(np.diff(node_ptr), seq_len):
                raise GraphValidationError(
                    "sequence lengths do not match node offsets")
            total = int(node_ptr[-1])
            residue_index = (np.arange(total, dtype=np.int64) - np.repeat(
                node_ptr[:-1], seq_len)).astype(np.int32)
            node_roles = np.zeros(total, np.uint8)
        shard = GraphShard(
            identifiers=tuple(sidecar["identifiers"]), sequences=sequences,
            structures=tuple(sidecar["structures"]),
            node_features=arrays["node_features"],
            edge_index=arrays["edge_index"], edge_types=arrays["edge_types"],
            node_ptr=node_ptr, edge_ptr=arrays["edge_ptr"], spec=spec,
            residue_index=residue_index, node_roles=node_roles)
    except GraphValidationError:
        raise
    except (KeyError, TypeError, ValueError) as exc:
        raise GraphValidationError(
            f"invalid graph shard metadata: {exc}") from exc
    if (sidecar.get("record_count") != shard.record_count
            or sidecar.get("node_count") != shard.node_count
            or sidecar.get("edge_count") != shard.edge_count):
        raise GraphValidationError("graph shard count metadata mismatch")
    if validation == "full":
        shard.validate_values()
    return shard


__all__ = [
    "GRAPH_SHARD_FORMAT", "GRAPH_SHARD_FORMAT_VERSION", "NODE_ROLE_CONTEXT",
    "NODE_ROLE_CORE", "Graph", "GraphBuilder", "GraphCompatibilityError",
    "GraphShard", "GraphSpec", "GraphValidationError", "graph_metadata_path",
    "load_graph_shard", "partition_records", "save_graph_shard",
]
