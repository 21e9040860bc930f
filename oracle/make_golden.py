#!/usr/bin/env python
"""Generate tests/golden/* by RUNNING THE REFERENCE in the build container.

TEST INFRASTRUCTURE.  Imports /root/reference/src (read-only) and records,
for a small fixed set of RNAs, what the unmodified reference produces:

  golden_records.json   ids / sequences / structures / windows, builder
                        options, packing known-answers, spec fingerprint
  golden_graphs.npz     GraphBuilder(...).build_shard arrays per case
  golden_embeddings.npz reference embeddings: fp32 model -> float32,
                        fp16 model (package default) -> float16

The reference cannot travel to the GPU box, these files can.
Usage:  python oracle/make_golden.py   (needs /root/reference)
"""
import json
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))
import ginfinity as ref  # noqa: E402
import ginfinity.api as ref_api  # noqa: E402

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)

KAT = [  # RNAs used by the reference's own tests
    ("rna-1", "ACGUACGU", "((....))"), ("rna-2", "GGAACCUU", "........"),
    ("first", "ACGU", "...."), ("second", "GGAA", "(())"),
    ("one", "A", "."), ("two", "AC", ".."), ("three", "GAC", "(.)"),
]
STEM = ("stem", "GGGAAACCCUUUUGGG", "......(((....)))")


def shard_arrays(shard, prefix):
    return {f"{prefix}/{name}": getattr(shard, name) for name in (
        "node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr",
        "residue_index", "node_roles")}


def observed_boundaries(encoder, shard, max_nodes, max_edges):
    """Record the microbatch boundaries the reference's greedy loop
    (api.py:211-229) actually produces, by watching _run_graph_shard."""
    sizes = []
    original = ref_api.Ginfinity._run_graph_shard

    def spy(self, sub, dtype):
        sizes.append(sub.record_count)
        return [np.zeros((0, 128), dtype)] * sub.record_count

    ref_api.Ginfinity._run_graph_shard = spy
    try:
        encoder.encode_graphs(shard, max_batch_nodes=max_nodes,
                              max_batch_edges=max_edges)
    finally:
        ref_api.Ginfinity._run_graph_shard = original
    return [0] + np.cumsum(sizes).tolist()


def main():
    rng = np.random.default_rng(20261018)
    table = ref.read_rna_table(REF / "tests" / "rouskin_sample_6k.tsv")
    lengths = np.array([r.length for r in table])
    # a spread of real records: short, median, long, and the longest <= 700
    picks = []
    for lo, hi, count in ((14, 60, 4), (60, 140, 6), (140, 260, 6),
                          (260, 420, 3), (420, 700, 1)):
        pool = np.flatnonzero((lengths >= lo) & (lengths < hi))
        picks += sorted(rng.choice(pool, size=count, replace=False).tolist())
    real = [table[i] for i in picks]
    full_records = [ref.RNA(*t) for t in KAT] + real

    windows = [
        dict(record=STEM + (9, 16), keep=False, hops=1),
        dict(record=STEM + (9, 16), keep=True, hops=1),
        dict(record=STEM + (9, 16), keep=True, hops=2),
        dict(record=STEM + (9, 16), keep=True, hops=3),
    ]
    long_one = max(real, key=lambda r: r.length)
    third = long_one.length // 3
    windows.append(dict(record=(long_one.identifier + ":w", long_one.sequence,
                                long_one.structure, third, 2 * third),
                        keep=True, hops=2))

    enc32 = ref.Ginfinity.load(full_precision=True)
    enc16 = ref.Ginfinity.load()
    graphs, embeds = {}, {}
    meta = {
        "reference_version": ref.__version__,
        "graph_spec_sha256": enc32.graph_spec.sha256,
        "checkpoint_sha256": enc32.info()["checkpoint_sha256"],
        "full": [[r.identifier, r.sequence, r.structure] for r in full_records],
        "windows": [dict(record=list(w["record"]), keep=w["keep"], hops=w["hops"])
                    for w in windows],
        "packing": [],
    }

    shard = ref.GraphBuilder().build_shard(full_records)
    graphs.update(shard_arrays(shard, "full"))
    embeds["full/fp32_model_f32"] = np.concatenate(
        enc32.encode_graphs(shard, embedding_dtype=np.float32))
    embeds["full/fp16_model_f16"] = np.concatenate(
        enc16.encode_graphs(shard))                       # package default
    # microbatch layout must not matter (tests/test_graph.py:85-98)
    small = enc32.encode_graphs(shard, max_batch_nodes=int(lengths[picks].max()),
                                max_batch_edges=5 * int(lengths[picks].max()),
                                embedding_dtype=np.float32)
    assert np.allclose(np.concatenate(small), embeds["full/fp32_model_f32"],
                       rtol=1e-5, atol=3e-7)

    for k, w in enumerate(windows):
        rec = ref.RNA(*w["record"][:3], start=w["record"][3], end=w["record"][4])
        builder = ref.GraphBuilder(keep_paired_neighbours=w["keep"],
                                   context_hops=w["hops"])
        ws = builder.build_shard([rec])
        graphs.update(shard_arrays(ws, f"window{k}"))
        embeds[f"window{k}/fp32_model_f32"] = enc32.encode_graphs(
            ws, embedding_dtype=np.float32)[0]
        embeds[f"window{k}/fp16_model_f16"] = enc16.encode_graphs(ws)[0]

    for max_nodes, max_edges in ((60_000, 300_000), (700, 3000), (300, 100_000),
                                 (100_000, 1400), (1000, 4600)):
        try:
            bounds = observed_boundaries(enc32, shard, max_nodes, max_edges)
        except ValueError as exc:
            bounds = {"error": str(exc)}
        meta["packing"].append(dict(max_batch_nodes=max_nodes,
                                    max_batch_edges=max_edges, bounds=bounds))
    # the whole rouskin file at default limits: 15 microbatches
    big = ref.GraphBuilder().build_shard(table)
    meta["rouskin"] = dict(
        records=big.record_count, nodes=big.node_count, edges=big.edge_count,
        bounds_default=observed_boundaries(enc32, big, 60_000, 300_000),
        bounds_8k_40k=observed_boundaries(enc32, big, 8_000, 40_000))
    import hashlib
    meta["rouskin"]["sha256"] = {
        name: hashlib.sha256(np.ascontiguousarray(getattr(big, name)).tobytes()
                             ).hexdigest()
        for name in ("node_features", "edge_index", "edge_types", "node_ptr",
                     "edge_ptr", "residue_index", "node_roles")}
    meta["rouskin"]["sha256_features_cols_0_4"] = hashlib.sha256(
        np.ascontiguousarray(big.node_features[:, :5]).tobytes()).hexdigest()

    (OUT / "golden_records.json").write_text(json.dumps(meta, indent=1) + "\n")
    np.savez_compressed(OUT / "golden_graphs.npz", **graphs)
    np.savez_compressed(OUT / "golden_embeddings.npz", **embeds)
    total = sum(r.length for r in full_records)
    print(f"golden: {len(full_records)} full records / {total} nt, "
          f"{len(windows)} windows -> {OUT}")
    for path in sorted(OUT.iterdir()):
        print(f"  {path.name}: {path.stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
