"""Import the real reference package (TEST / BENCH INFRASTRUCTURE ONLY).

`reference_root()` is oracle/_ref (staged by `python -m oracle.stage_ref`, the
copy that travels to the GPU box); in the build container, when nothing has
been staged yet, it falls back to the read-only checkout /root/reference.
`import_reference()` returns the imported `ginfinity` package (and its `api`
module) from there.  Only tests/, __graft_entry__.smoke() and bench.py's
reference legs call this; nothing under ginfinity_b200/ does.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path
from typing import Optional

import numpy as np

HERE = Path(__file__).resolve().parent


def reference_root() -> Optional[Path]:
    for root in (HERE / "_ref", Path("/root/reference")):
        if (root / "src" / "ginfinity" / "api.py").is_file():
            return root
    return None


def import_reference():
    """(ginfinity package, ginfinity.api module) of the unmodified reference."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference has not been staged: run `python -m oracle.stage_ref` "
                          "where /root/reference exists")
    src = str(root / "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    package = importlib.import_module("ginfinity")
    if not str(Path(package.__file__).resolve()).startswith(str(root.resolve())):
        raise ImportError(f"`ginfinity` resolved to {package.__file__}, not to the staged reference")
    return package, importlib.import_module("ginfinity.api")


def rouskin_table() -> Optional[Path]:
    root = reference_root()
    path = None if root is None else root / "tests" / "rouskin_sample_6k.tsv"
    return path if path is not None and path.is_file() else None


def to_reference_shard(ref, shard):
    """The same arrays as a reference `GraphShard` (validated by the reference's
    own `__post_init__`, graph.py:277-343)."""
    spec = ref.GraphSpec.from_dict(shard.spec.to_dict())
    return ref.GraphShard(
        identifiers=tuple(shard.identifiers), sequences=tuple(shard.sequences),
        structures=tuple(shard.structures),
        node_features=np.ascontiguousarray(shard.node_features),
        edge_index=np.ascontiguousarray(shard.edge_index),
        edge_types=np.ascontiguousarray(shard.edge_types),
        node_ptr=np.ascontiguousarray(shard.node_ptr),
        edge_ptr=np.ascontiguousarray(shard.edge_ptr), spec=spec,
        residue_index=np.ascontiguousarray(shard.residue_index),
        node_roles=np.ascontiguousarray(shard.node_roles))
