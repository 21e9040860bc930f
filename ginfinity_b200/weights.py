"""Checkpoint verification and weight preparation for the device encoder.

Reference behaviour mirrored here:
  * integrity and compatibility checks of ``Ginfinity.load``
    (src/ginfinity/api.py:64-114): SHA-256 of ``encoder.pt`` against
    ``model.json``, format version, architecture and graph-spec agreement,
    strict state-dict key/shape match, parameter count.
  * ``EncoderConfig`` (src/ginfinity/_model.py:11-26).

New here (no reference counterpart): ``fold`` rewrites the trained tensors
into the form the kernels consume.  The algebra is exact in real numbers:

  * ``edge_lin(one_hot(t)) = W_e[:, t] + b_e``  ->  a [edge_dim, hidden]
    lookup table per layer (_model.py:33,43 with api.py:243-245);
  * eval-mode ``BatchNorm1d`` after ``mlp.0`` is a per-channel affine map
    (_model.py:35), folded into the rows of ``W1`` and into ``b1``.

Folding is done in float64 and rounded once to float32.
"""
from __future__ import annotations

import hashlib
import json
from dataclasses import dataclass
from pathlib import Path
from typing import Mapping, Optional

import numpy as np

from .graph import GraphSpec

BN_EPS = 1e-5      # torch.nn.BatchNorm1d default, used by _model.py:35
LN_EPS = 1e-5      # torch.nn.LayerNorm default, used by _model.py:59-60


class ModelIntegrityError(RuntimeError):
    """A packaged model artifact failed compatibility or integrity checks."""


@dataclass(frozen=True)
class EncoderConfig:
    hidden: int
    layers: int
    out_dim: int
    dropout: float
    struct_feature: str
    positional: bool
    residual: bool
    train_eps: bool
    edge_dim: int
    extra_edges: tuple

    @classmethod
    def from_dict(cls, value: Mapping) -> "EncoderConfig":
        fields = dict(value)
        fields["extra_edges"] = tuple(fields.get("extra_edges", ()))
        return cls(**fields)

    @property
    def feature_dim(self) -> int:
        return 4 + (1 if self.struct_feature == "A" else 3) + (
            2 if self.positional else 0)


BUNDLED_CONFIG = EncoderConfig(
    hidden=128, layers=4, out_dim=128, dropout=0.023642890587088308,
    struct_feature="A", positional=True, residual=True, train_eps=True,
    edge_dim=10, extra_edges=("skip2",))


def expected_state_shapes(cfg: EncoderConfig) -> dict:
    """Every state-dict entry of GINEEncoder(cfg) and its shape
    (_model.py:29-63)."""
    h, d = cfg.hidden, cfg.feature_dim
    shapes = {"input.weight": (h, d), "input.bias": (h,)}
    for l in range(cfg.layers):
        p = f"convs.{l}."
        shapes.update({
            p + "eps": (1,),
            p + "edge_lin.weight": (h, cfg.edge_dim),
            p + "edge_lin.bias": (h,),
            p + "mlp.0.weight": (2 * h, h), p + "mlp.0.bias": (2 * h,),
            p + "mlp.1.weight": (2 * h,), p + "mlp.1.bias": (2 * h,),
            p + "mlp.1.running_mean": (2 * h,),
            p + "mlp.1.running_var": (2 * h,),
            p + "mlp.1.num_batches_tracked": (),
            p + "mlp.4.weight": (h, 2 * h), p + "mlp.4.bias": (h,),
            f"norms.{l}.weight": (h,), f"norms.{l}.bias": (h,),
        })
    shapes.update({"head.0.weight": (h, h), "head.0.bias": (h,),
                   "head.2.weight": (cfg.out_dim, h),
                   "head.2.bias": (cfg.out_dim,)})
    return shapes


_BUFFER_SUFFIXES = ("running_mean", "running_var", "num_batches_tracked")


def parameter_count(state: Mapping[str, np.ndarray]) -> int:
    return int(sum(v.size for k, v in state.items()
                   if not k.endswith(_BUFFER_SUFFIXES)))


def load_checkpoint(model_dir) -> tuple:
    """Verify and read ``model_dir/{model.json,encoder.pt}``.

    Returns ``(state, cfg, graph_spec, metadata)`` with ``state`` a dict of
    float32 NumPy arrays.  Same checks, same order and same messages as
    api.py:77-109.
    """
    import torch

    root = Path(model_dir)
    meta_path, ckpt_path = root / "model.json", root / "encoder.pt"
    try:
        metadata = json.loads(meta_path.read_text())
    except (OSError, json.JSONDecodeError) as exc:
        raise ModelIntegrityError(
            f"cannot read model metadata {meta_path}: {exc}") from exc
    if not ckpt_path.is_file():
        raise ModelIntegrityError(f"missing checkpoint {ckpt_path}")
    digest = hashlib.sha256(ckpt_path.read_bytes()).hexdigest()
    if digest != metadata.get("checkpoint_sha256"):
        raise ModelIntegrityError("checkpoint SHA-256 mismatch")
    if metadata.get("format_version") != 1:
        raise ModelIntegrityError("unsupported model format")
    try:
        payload = torch.load(ckpt_path, map_location="cpu", weights_only=True)
        cfg = EncoderConfig.from_dict(payload["cfg"])
        if cfg != EncoderConfig.from_dict(metadata["encoder_config"]):
            raise ModelIntegrityError(
                "checkpoint and metadata architecture mismatch")
        spec = GraphSpec.from_dict(metadata["graph_spec"])
        if (spec.sha256 != GraphSpec.from_encoder_config(cfg).sha256
                or metadata.get("graph_spec_sha256") != spec.sha256):
            raise ModelIntegrityError("model and graph specification mismatch")
        state = _strict_state(payload["state_dict"], cfg)
    except ModelIntegrityError:
        raise
    except Exception as exc:
        raise ModelIntegrityError(
            f"checkpoint could not be loaded: {exc}") from exc
    if parameter_count(state) != metadata.get("parameter_count"):
        raise ModelIntegrityError("parameter-count mismatch")
    return state, cfg, spec, metadata


def _strict_state(raw: Mapping, cfg: EncoderConfig) -> dict:
    want = expected_state_shapes(cfg)
    missing = sorted(set(want) - set(raw))
    unexpected = sorted(set(raw) - set(want))
    if missing or unexpected:
        raise KeyError(f"state_dict mismatch: missing={missing} "
                       f"unexpected={unexpected}")
    state = {}
    for name, shape in want.items():
        value = raw[name]
        array = (value.detach().cpu().numpy() if hasattr(value, "detach")
                 else np.asarray(value))
        if tuple(array.shape) != tuple(shape):
            raise ValueError(f"size mismatch for {name}: {tuple(array.shape)} "
                             f"vs {tuple(shape)}")
        if name.endswith("num_batches_tracked"):
            state[name] = array.astype(np.int64)
        else:
            state[name] = np.ascontiguousarray(array, dtype=np.float32)
    return state


def synthetic_state(cfg: EncoderConfig = BUNDLED_CONFIG, seed: int = 0) -> dict:
    """Random weights of the bundled architecture, scaled like a trained
    model (activations stay O(1..10) through 4 layers).  Used by tests and
    by bench.py when no checkpoint has been staged."""
    rng = np.random.default_rng(seed)
    state = {}
    for name, shape in expected_state_shapes(cfg).items():
        if name.endswith("num_batches_tracked"):
            state[name] = np.asarray(1000, np.int64)
        elif name.endswith("running_var"):
            state[name] = rng.uniform(0.5, 2.0, shape).astype(np.float32)
        elif name.endswith("running_mean"):
            state[name] = rng.normal(0, 0.3, shape).astype(np.float32)
        elif name.endswith(".eps"):
            state[name] = rng.uniform(-0.7, -0.2, shape).astype(np.float32)
        elif ("norms." in name or "mlp.1." in name) and name.endswith("weight"):
            state[name] = rng.uniform(0.6, 1.4, shape).astype(np.float32)
        elif name.endswith("bias"):
            state[name] = rng.normal(0, 0.1, shape).astype(np.float32)
        else:
            fan_in = shape[-1]
            state[name] = rng.normal(0, 1.0 / np.sqrt(fan_in), shape).astype(
                np.float32)
    return state


@dataclass(frozen=True)
class FoldedWeights:
    """Kernel-ready float32 tensors (all C-contiguous).

    w_in [H,F] b_in [H]; per layer l: table[l] [edge_dim,H], eps1[l] = 1+eps,
    w1[l] [2H,H] (BN folded), b1[l] [2H], w2[l] [H,2H], b2[l] [H],
    ln_g[l] [H], ln_b[l] [H]; head wa [H,H] ba [H] wb [O,H] bb [O].
    """
    cfg: EncoderConfig
    w_in: np.ndarray
    b_in: np.ndarray
    table: np.ndarray
    eps1: np.ndarray
    w1: np.ndarray
    b1: np.ndarray
    w2: np.ndarray
    b2: np.ndarray
    ln_g: np.ndarray
    ln_b: np.ndarray
    wa: np.ndarray
    ba: np.ndarray
    wb: np.ndarray
    bb: np.ndarray


def fold(state: Mapping[str, np.ndarray],
         cfg: EncoderConfig = BUNDLED_CONFIG) -> FoldedWeights:
    f64 = lambda name: np.asarray(state[name], np.float64)  # noqa: E731
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    table, eps1, w1, b1, w2, b2, g, b = ([] for _ in range(8))
    for l in range(cfg.layers):
        p = f"convs.{l}."
        # one_hot(t) @ W_e^T + b_e  ==  W_e[:, t] + b_e
        table.append(f64(p + "edge_lin.weight").T + f64(p + "edge_lin.bias"))
        eps1.append(1.0 + f64(p + "eps")[0])
        scale = f64(p + "mlp.1.weight") / np.sqrt(
            f64(p + "mlp.1.running_var") + BN_EPS)
        w1.append(f64(p + "mlp.0.weight") * scale[:, None])
        b1.append((f64(p + "mlp.0.bias") - f64(p + "mlp.1.running_mean"))
                  * scale + f64(p + "mlp.1.bias"))
        w2.append(f64(p + "mlp.4.weight"))
        b2.append(f64(p + "mlp.4.bias"))
        g.append(f64(f"norms.{l}.weight"))
        b.append(f64(f"norms.{l}.bias"))
    return FoldedWeights(
        cfg=cfg, w_in=f32(state["input.weight"]), b_in=f32(state["input.bias"]),
        table=f32(np.stack(table)), eps1=f32(np.asarray(eps1)),
        w1=f32(np.stack(w1)), b1=f32(np.stack(b1)), w2=f32(np.stack(w2)),
        b2=f32(np.stack(b2)), ln_g=f32(np.stack(g)), ln_b=f32(np.stack(b)),
        wa=f32(state["head.0.weight"]), ba=f32(state["head.0.bias"]),
        wb=f32(state["head.2.weight"]), bb=f32(state["head.2.bias"]))


def default_model_dir() -> Optional[Path]:
    """Where a verified checkpoint is looked for when ``model_dir`` is None:
    ``$GINFINITY_MODEL_DIR``, then ``ginfinity_b200/data`` (populated by
    ``python -m ginfinity_b200.stage_model``; the weights are CC BY-NC and
    are not committed to this repository)."""
    import os

    env = os.environ.get("GINFINITY_MODEL_DIR")
    if env:
        return Path(env)
    here = Path(__file__).resolve().parent / "data"
    if (here / "encoder.pt").is_file():
        return here
    return None
