"""ginfinity_b200 -- B200-native GINFINITY encoder path.

Same public names as the reference package (src/ginfinity/__init__.py:14-36).
The data plane (RNA, GraphSpec, Graph, GraphShard, GraphBuilder, shard files)
is importable anywhere; `Ginfinity` needs the built libgfx.so and a CUDA
device and fails loudly without them -- there is no CPU fallback.
"""
from .records import InputValidationError, RNA, read_rna_table
from .graph import (GRAPH_SHARD_FORMAT, GRAPH_SHARD_FORMAT_VERSION,
                    NODE_ROLE_CONTEXT, NODE_ROLE_CORE, Graph, GraphBuilder,
                    GraphCompatibilityError, GraphShard, GraphSpec,
                    GraphValidationError, graph_metadata_path,
                    load_graph_shard, partition_records, save_graph_shard)
from .weights import EncoderConfig, ModelIntegrityError

__version__ = "0.1.0"

_LAZY = {"Ginfinity": "encoder", "DeviceShard": "encoder",
         "default_alignment_parameters": "encoder",
         "EmbeddingIndex": "search"}


def __getattr__(name):
    module = _LAZY.get(name)
    if module is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    import importlib
    return getattr(importlib.import_module(f"{__name__}.{module}"), name)


__all__ = [
    "Ginfinity", "GRAPH_SHARD_FORMAT", "GRAPH_SHARD_FORMAT_VERSION",
    "NODE_ROLE_CONTEXT", "NODE_ROLE_CORE", "Graph", "GraphBuilder",
    "GraphCompatibilityError", "GraphShard", "GraphSpec",
    "GraphValidationError", "InputValidationError", "ModelIntegrityError",
    "RNA", "default_alignment_parameters", "graph_metadata_path",
    "load_graph_shard", "partition_records", "read_rna_table",
    "save_graph_shard", "DeviceShard", "EmbeddingIndex", "EncoderConfig",
    "__version__",
]
