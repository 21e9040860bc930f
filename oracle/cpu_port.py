"""Multi-threaded CPU port of the reference encoder path.  TEST/BENCH
INFRASTRUCTURE ONLY (bench.py `cpu_baseline` and `--impl reference`).

The reference (Python + torch eager) cannot travel to the GPU box, so its
CPU path is restated here with the same torch ops in the same order and the
same data flow, so that it costs what the reference costs on the same host
cores:

  encode_graphs greedy packing + shard.slice     src/ginfinity/api.py:211-229
  int32->int64 / uint8->int64 casts, one-hot      api.py:237-245
  GINEEncoder.forward (index_select, add, relu,
    index_add_, Linear, BatchNorm1d eval, ReLU,
    Linear, LayerNorm, residual, head)            _model.py:39-46, 65-72
  float64 L2 normalise, per-record core split     api.py:250-259

Validated against the reference's recorded outputs in
tests/test_oracle_golden.py::test_cpu_port_matches_reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .gine_oracle import pack_microbatches


class CpuPort:
    def __init__(self, state: dict, *, layers: int = 4, edge_dim: int = 10,
                 full_precision: bool = False):
        self.layers, self.edge_dim = layers, edge_dim
        self.dtype = torch.float32 if full_precision else torch.float16
        self.p = {k: torch.from_numpy(np.asarray(v)).to(
            self.dtype if np.asarray(v).dtype.kind == "f" else torch.int64)
            for k, v in state.items()}

    @torch.inference_mode()
    def forward(self, x, edge_index, edge_attr):
        p = self.p
        h = F.linear(x, p["input.weight"], p["input.bias"])
        src, dst = edge_index
        for l in range(self.layers):
            c = f"convs.{l}."
            msg = F.relu(h.index_select(0, src) + F.linear(
                edge_attr, p[c + "edge_lin.weight"], p[c + "edge_lin.bias"]))
            agg = torch.zeros_like(h).index_add_(0, dst, msg)
            z = (1.0 + p[c + "eps"]) * h + agg
            a = F.linear(z, p[c + "mlp.0.weight"], p[c + "mlp.0.bias"])
            a = F.batch_norm(a, p[c + "mlp.1.running_mean"], p[c + "mlp.1.running_var"],
                             p[c + "mlp.1.weight"], p[c + "mlp.1.bias"], False, 0.1, 1e-5)
            u = F.linear(F.relu(a), p[c + "mlp.4.weight"], p[c + "mlp.4.bias"])
            u = F.layer_norm(u, (u.shape[1],), p[f"norms.{l}.weight"], p[f"norms.{l}.bias"], 1e-5)
            h = h + u
        t = F.relu(F.linear(h, p["head.0.weight"], p["head.0.bias"]))
        return F.linear(t, p["head.2.weight"], p["head.2.bias"])

    @torch.inference_mode()
    def run_microbatch(self, shard, a, b, embedding_dtype):
        n0, n1 = int(shard.node_ptr[a]), int(shard.node_ptr[b])
        e0, e1 = int(shard.edge_ptr[a]), int(shard.edge_ptr[b])
        # shard.slice: copies + rebasing (graph.py:414-444)
        feats = np.ascontiguousarray(shard.node_features[n0:n1])
        ei = np.ascontiguousarray(shard.edge_index[:, e0:e1] - np.int32(n0), dtype=np.int32)
        et = np.ascontiguousarray(shard.edge_types[e0:e1])
        roles = shard.node_roles[n0:n1]
        node = torch.from_numpy(feats).to(dtype=self.dtype)
        edge_index = torch.from_numpy(ei).to(dtype=torch.long)
        edge_types = torch.from_numpy(et).to(dtype=torch.long)
        attr = F.one_hot(edge_types, num_classes=self.edge_dim).to(dtype=self.dtype)
        emb = self.forward(node, edge_index, attr).to(torch.float32).numpy().astype(np.float64)
        emb = emb / np.maximum(np.linalg.norm(emb, axis=1, keepdims=True), 1e-12)
        out = []
        for i in range(a, b):
            s, t = int(shard.node_ptr[i]) - n0, int(shard.node_ptr[i + 1]) - n0
            core = roles[s:t] == 0
            out.append(np.ascontiguousarray(emb[s:t][core], dtype=embedding_dtype))
        return out

    def encode_graphs(self, shard, *, max_batch_nodes=60_000, max_batch_edges=300_000,
                      embedding_dtype=np.float16):
        lengths = np.diff(shard.node_ptr).tolist()
        ecounts = np.diff(shard.edge_ptr).tolist()
        bounds = pack_microbatches(lengths, ecounts, max_batch_nodes, max_batch_edges)
        out = []
        for a, b in zip(bounds[:-1].tolist(), bounds[1:].tolist()):
            out.extend(self.run_microbatch(shard, a, b, embedding_dtype))
        return out
