"""ginfinity_b200 -- B200-native GINFINITY encoder path."""
from .records import InputValidationError, RNA, read_rna_table
from .graph import (GRAPH_SHARD_FORMAT, GRAPH_SHARD_FORMAT_VERSION,
                    NODE_ROLE_CONTEXT, NODE_ROLE_CORE, Graph, GraphBuilder,
                    GraphCompatibilityError, GraphShard, GraphSpec,
                    GraphValidationError, graph_metadata_path,
                    load_graph_shard, partition_records, save_graph_shard)
from .weights import EncoderConfig, ModelIntegrityError

__version__ = "0.1.0"
