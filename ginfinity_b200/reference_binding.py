"""The reference-side binding of libgfx.so: what a GINFINITY maintainer would
add as `src/ginfinity/_gfx.py` (INTEGRATION.md quotes this file).

It is SELF-CONTAINED on purpose -- ctypes, NumPy and torch only, nothing from
ginfinity_b200 -- and touches the reference in exactly one place:

    import ginfinity.api as api
    from ginfinity_b200 import reference_binding      # or a copy named ginfinity._gfx
    reference_binding.install(api)

replaces the body of `Ginfinity._run_graph_shard` (src/ginfinity/api.py:232-260)
and nothing else: `Ginfinity.load`'s integrity checks, `encode_graphs`' argument
checks, its greedy packing loop (api.py:211-229), `GraphShard.slice` and the
return contract stay the reference's own code.  tests/test_gpu_reference.py runs
the reference's own test-suite through it on a B200.

The loaded torch module stays where `Ginfinity.load` put it (its tests look at
it); the kernels get their own folded copy of the weights on the GPU the first
time an encoder is used.  The device is the encoder's when that is a CUDA
device, else cuda:0 (the reference's tests call `Ginfinity.load()` with the
default "cpu").
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from pathlib import Path

import numpy as np
import torch

_LIB_PATH = os.environ.get("GFX_LIBRARY", str(Path(__file__).resolve().parent / "libgfx.so"))
_lib = C.CDLL(_LIB_PATH)
_p, _i64, _sz = C.c_void_p, C.c_int64, C.c_size_t
_lib.gfx_last_error.restype = C.c_char_p
_lib.gfx_csr_workspace_bytes.restype = _sz
_lib.gfx_csr_workspace_bytes.argtypes = [_i64, _i64]
_lib.gfx_encode_workspace_bytes.restype = _sz
_lib.gfx_encode_workspace_bytes.argtypes = [_i64, C.c_int]
_lib.gfx_csr_build.argtypes = [_p, _p, _p, _i64, _i64, C.c_int32, _p, _p, _p, _p, _sz, _p]
_lib.gfx_encode.argtypes = [_p, _p, _p, _p, _p, _p, _i64, _p, C.c_int, C.c_int, C.c_int, C.c_int,
                            _p, _sz, _p]
_lib.gfx_model_create.argtypes = [_p, C.POINTER(_p)]
_lib.gfx_model_destroy.argtypes = [_p]
GFX_F16, GFX_F32 = 0, 1
GFX_FUSED_BANDED = 3       # gfx_encode `fused`: falls back by itself where it does not apply


def _check(rc):
    if rc:
        raise RuntimeError(_lib.gfx_last_error().decode())


class _Weights(C.Structure):          # gfx_folded_weights (include/gfx.h)
    _fields_ = [(n, C.c_int32) for n in ("hidden", "layers", "out_dim", "feature_dim", "edge_dim")] + \
               [(n, _p) for n in ("w_in", "b_in", "table", "eps1", "w1", "b1", "w2", "b2",
                                  "ln_g", "ln_b", "wa", "ba", "wb", "bb")]


def create_model(model) -> int:
    """Fold the eval-mode `GINEEncoder` into the arrays gfx_model_create takes
    (on the CURRENT CUDA device)."""
    sd = {k: v.detach().to(torch.float64).cpu().numpy() for k, v in model.state_dict().items()}
    L = len(model.convs)
    keep = []                                      # keep the arrays alive during the call

    def f32(a):
        a = np.ascontiguousarray(a, np.float32)
        keep.append(a)
        return a.ctypes.data

    w1, b1, table, eps1 = [], [], [], []
    for l in range(L):
        g, b = sd[f"convs.{l}.mlp.1.weight"], sd[f"convs.{l}.mlp.1.bias"]
        mu, var = sd[f"convs.{l}.mlp.1.running_mean"], sd[f"convs.{l}.mlp.1.running_var"]
        s = g / np.sqrt(var + 1e-5)                # BatchNorm1d eval (_model.py:35)
        w1.append(sd[f"convs.{l}.mlp.0.weight"] * s[:, None])
        b1.append((sd[f"convs.{l}.mlp.0.bias"] - mu) * s + b)
        # edge_lin(one_hot(t)) = W_e[:, t] + b_e  (_model.py:33,43; api.py:243-245)
        table.append(sd[f"convs.{l}.edge_lin.weight"].T + sd[f"convs.{l}.edge_lin.bias"])
        eps1.append(1.0 + float(sd[f"convs.{l}.eps"].reshape(-1)[0]))
    cfg = model.cfg
    w = _Weights(cfg.hidden, L, cfg.out_dim, sd["input.weight"].shape[1], cfg.edge_dim,
                 f32(sd["input.weight"]), f32(sd["input.bias"]), f32(np.stack(table)), f32(eps1),
                 f32(np.stack(w1)), f32(np.stack(b1)),
                 f32(np.stack([sd[f"convs.{l}.mlp.4.weight"] for l in range(L)])),
                 f32(np.stack([sd[f"convs.{l}.mlp.4.bias"] for l in range(L)])),
                 f32(np.stack([sd[f"norms.{l}.weight"] for l in range(L)])),
                 f32(np.stack([sd[f"norms.{l}.bias"] for l in range(L)])),
                 f32(sd["head.0.weight"]), f32(sd["head.0.bias"]),
                 f32(sd["head.2.weight"]), f32(sd["head.2.bias"]))
    handle = _p()
    _check(_lib.gfx_model_create(C.byref(w), C.byref(handle)))
    return handle.value


def _gfx_device(encoder) -> torch.device:
    device = torch.device(encoder.device)
    return device if device.type == "cuda" else torch.device("cuda:0")


@torch.inference_mode()
def _run_graph_shard(self, shard, embedding_dtype):
    """Drop-in body of Ginfinity._run_graph_shard (api.py:232-260)."""
    dev = _gfx_device(self)
    with torch.cuda.device(dev):
        if getattr(self, "_gfx_model", None) is None:
            # the fp16 module was rounded by .half(); the kernels fold the same values
            self._gfx_model = create_model(self._model)
            weakref.finalize(self, _lib.gfx_model_destroy, self._gfx_model)
        st = torch.cuda.current_stream().cuda_stream
        n, e = shard.node_count, shard.edge_count
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)  # noqa: E731
        x, ei, et = up(shard.node_features), up(shard.edge_index), up(shard.edge_types)
        u8 = lambda nbytes: torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)  # noqa: E731
        row_ptr, col_src, col_type = u8(4 * (n + 1)), u8(4 * e), u8(e)
        ws = u8(_lib.gfx_csr_workspace_bytes(n, e))
        _check(_lib.gfx_csr_build(ei[0].data_ptr() if e else None, ei[1].data_ptr() if e else None,
                                  et.data_ptr() if e else None, n, e, 0,
                                  row_ptr.data_ptr(), col_src.data_ptr(), col_type.data_ptr(),
                                  ws.data_ptr(), ws.numel(), st))
        act = GFX_F32 if self.full_precision else GFX_F16
        out = torch.empty((n, 128), dtype=torch.float32, device=dev)
        ews = u8(_lib.gfx_encode_workspace_bytes(n, act))
        all_core = not bool(np.any(shard.node_roles))
        fused = 0 if self.full_precision else (GFX_FUSED_BANDED if all_core else 2)
        _check(_lib.gfx_encode(self._gfx_model, x.data_ptr(), row_ptr.data_ptr(), col_src.data_ptr(),
                               col_type.data_ptr(), None, n, out.data_ptr(), act, GFX_F32, 0, fused,
                               ews.data_ptr(), ews.numel(), st))
        embeddings = out.cpu().numpy()             # already unit-norm; rows still per node
    outputs = []
    for index in range(shard.record_count):        # api.py:253-259 unchanged
        a, b = int(shard.node_ptr[index]), int(shard.node_ptr[index + 1])
        core = shard.node_roles[a:b] == 0
        outputs.append(np.ascontiguousarray(embeddings[a:b][core], dtype=embedding_dtype))
    return outputs


def install(api_module) -> None:
    """Patch `api_module.Ginfinity._run_graph_shard` (the one seam of the path)."""
    api_module.Ginfinity._gfx_model = None
    api_module.Ginfinity._run_graph_shard = _run_graph_shard


def uninstall(api_module, original) -> None:
    api_module.Ginfinity._run_graph_shard = original
