import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def golden_meta():
    return json.loads((GOLDEN / "golden_records.json").read_text())


@pytest.fixture(scope="session")
def golden_graphs():
    return dict(np.load(GOLDEN / "golden_graphs.npz"))


@pytest.fixture(scope="session")
def golden_embeddings():
    return dict(np.load(GOLDEN / "golden_embeddings.npz"))


@pytest.fixture(scope="session")
def golden_shard(golden_meta):
    import ginfinity_b200 as g
    records = [g.RNA(*t) for t in golden_meta["full"]]
    return g.GraphBuilder().build_shard(records)


def staged_model_dir():
    from ginfinity_b200.weights import default_model_dir
    return default_model_dir()


@pytest.fixture(scope="session")
def real_state():
    """The reference checkpoint, if it has been staged (git-ignored)."""
    root = staged_model_dir()
    if root is None:
        pytest.skip("reference checkpoint not staged (python -m ginfinity_b200.stage_model)")
    from ginfinity_b200.weights import load_checkpoint
    return load_checkpoint(root)[0]


@pytest.fixture(scope="session")
def synthetic_state():
    from ginfinity_b200.weights import synthetic_state
    return synthetic_state(seed=7)
