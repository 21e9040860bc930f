"""pytest plugin: run the REFERENCE's own tests with libgfx.so under its API.

Loaded with `-p ref_binding_plugin` (PYTHONPATH=tests) by tests/test_gpu_reference.py.  It
imports the staged, unmodified reference (oracle/_ref), installs the C-ABI
binding of INTEGRATION.md (ginfinity_b200/reference_binding.py) over
`ginfinity.api.Ginfinity._run_graph_shard` -- the one seam of the hot path,
api.py:232-260 -- and counts how often the reference's code went through it.
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

CALLS = {"count": 0, "nodes": 0}


def pytest_configure(config):
    from oracle.ref_loader import import_reference
    _ref, api = import_reference()
    from ginfinity_b200 import reference_binding
    reference_binding.install(api)
    bound = api.Ginfinity._run_graph_shard

    def counted(self, shard, embedding_dtype):
        CALLS["count"] += 1
        CALLS["nodes"] += shard.node_count
        return bound(self, shard, embedding_dtype)

    api.Ginfinity._run_graph_shard = counted


def pytest_terminal_summary(terminalreporter):
    from ginfinity_b200 import _native
    launches = sum(_native.launch_counts().values())
    terminalreporter.write_line(
        f"gfx-binding: calls={CALLS['count']} nodes={CALLS['nodes']} libgfx_launches={launches}")
