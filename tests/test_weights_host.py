"""Checkpoint integrity checks of `Ginfinity.load` (reference api.py:77-109), restated from the
reference's tests/test_api.py:44-54, 79-91 against `weights.load_checkpoint` (what
`Ginfinity.load` calls before any device work), plus the architecture guard of the kernels."""
import json
import shutil

import numpy as np
import pytest

from conftest import staged_model_dir


@pytest.fixture()
def model_copy(tmp_path):
    source = staged_model_dir()
    if source is None:
        pytest.skip("reference checkpoint not staged (python -m ginfinity_b200.stage_model)")
    destination = tmp_path / "model"
    shutil.copytree(source, destination)
    return destination


def _rewrite_metadata(root, **changes):
    path = root / "model.json"
    data = json.loads(path.read_text())
    data.update(changes)
    path.write_text(json.dumps(data))


def test_checkpoint_tampering_is_rejected(model_copy):
    from ginfinity_b200.weights import ModelIntegrityError, load_checkpoint
    checkpoint = model_copy / "encoder.pt"
    payload = bytearray(checkpoint.read_bytes())
    payload[-1] ^= 1
    checkpoint.write_bytes(payload)
    with pytest.raises(ModelIntegrityError, match="SHA-256"):
        load_checkpoint(model_copy)


def test_unsupported_format_parameter_count_and_spec_mismatch_are_rejected(model_copy):
    from ginfinity_b200.weights import ModelIntegrityError, load_checkpoint
    original = (model_copy / "model.json").read_text()
    _rewrite_metadata(model_copy, format_version=2)
    with pytest.raises(ModelIntegrityError, match="unsupported model format"):
        load_checkpoint(model_copy)
    (model_copy / "model.json").write_text(original)
    _rewrite_metadata(model_copy, parameter_count=306_437)
    with pytest.raises(ModelIntegrityError, match="parameter-count"):
        load_checkpoint(model_copy)
    (model_copy / "model.json").write_text(original)
    _rewrite_metadata(model_copy, graph_spec_sha256="0" * 64)
    with pytest.raises(ModelIntegrityError, match="graph specification"):
        load_checkpoint(model_copy)
    (model_copy / "model.json").write_text(original)
    cfg = json.loads(original)["encoder_config"]
    _rewrite_metadata(model_copy, encoder_config={**cfg, "layers": 5})
    with pytest.raises(ModelIntegrityError, match="architecture mismatch"):
        load_checkpoint(model_copy)
    (model_copy / "model.json").write_text("{not json")
    with pytest.raises(ModelIntegrityError, match="cannot read model metadata"):
        load_checkpoint(model_copy)
    (model_copy / "model.json").write_text(original)
    (model_copy / "encoder.pt").unlink()
    with pytest.raises(ModelIntegrityError, match="missing checkpoint"):
        load_checkpoint(model_copy)


def test_checkpoint_loading_is_restricted_to_weights_only(model_copy, monkeypatch):
    import torch
    from ginfinity_b200.weights import load_checkpoint
    original, observed = torch.load, {}

    def wrapped(*args, **kwargs):
        observed.update(kwargs)
        return original(*args, **kwargs)

    monkeypatch.setattr(torch, "load", wrapped)
    state, cfg, spec, metadata = load_checkpoint(model_copy)
    assert observed["weights_only"] is True
    assert metadata["parameter_count"] == 306_436
    json.dumps(metadata)                           # api.py:116-126: info() is JSON-serialisable
    assert spec.sha256 == metadata["graph_spec_sha256"]
    assert all(v.dtype in (np.float32, np.int64) for v in state.values())


def test_architectures_the_kernels_do_not_implement_are_refused():
    """cfg.residual=False (or another feature layout) passes the reference's own checks; the
    kernels hard-code h + LayerNorm(...) over 7 features, so such a model is refused, not
    silently computed wrongly (reference honours the flag, _model.py:68-71)."""
    import dataclasses
    from ginfinity_b200.encoder import _check_architecture
    from ginfinity_b200.weights import BUNDLED_CONFIG, ModelIntegrityError
    _check_architecture(BUNDLED_CONFIG)
    for change in (dict(residual=False), dict(struct_feature="B"), dict(positional=False),
                   dict(hidden=64)):
        with pytest.raises(ModelIntegrityError, match="unsupported encoder architecture"):
            _check_architecture(dataclasses.replace(BUNDLED_CONFIG, **change))
