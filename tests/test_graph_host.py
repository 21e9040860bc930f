"""Host data plane: RNA contract, GraphSpec fingerprint, vectorised
GraphBuilder, GraphShard validation/slicing, shard files, partitioning.

These restate the reference's tests for the same surface
(tests/test_validation.py, tests/test_graph.py, tests/test_sliced_graphs.py)
and add bit-exact comparisons with the arrays the reference produced
(tests/golden/golden_graphs.npz) and with the loop-based oracle.
"""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest
from safetensors import safe_open
from safetensors.numpy import save_file

import ginfinity_b200 as g
from ginfinity_b200 import graph as graph_module
from ginfinity_b200.records import parse_position_list
from helpers import random_records
from oracle import gine_oracle as O

ARRAYS = ("node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr",
          "residue_index", "node_roles")
_SEQ, _DBN = "GGGAAACCCUUUUGGG", "......(((....)))"


def _records():
    return [g.RNA("rna-1", "ACGUACGU", "((....))"), g.RNA("rna-2", "GGAACCUU", "........")]


# ---- RNA contract (reference tests/test_validation.py) ---------------------
def test_rna_normalises_and_validates():
    r = g.RNA(" id ", "acgt", "(..)")
    assert (r.identifier, r.sequence, r.structure, r.length) == ("id", "ACGU", "(..)", 4)
    assert not r.sliced and r.core_length == 4
    with pytest.raises(AttributeError):
        r.sequence = "A"
    for seq, dbn, frag in (("", "", "empty sequence"), ("ACGN", "....", "unsupported sequence"),
                           ("ACGU", "..[.", "unsupported structure"), ("ACGU", "...", "structure is 3"),
                           ("ACGU", "())(", "unmatched ')' at 0-based position 2"),
                           ("ACGU", "((.)", "unmatched '(' at 0-based position 0"),
                           ("A" * 4097, "." * 4097, "exceeds maximum 4096")):
        with pytest.raises(g.InputValidationError, match=frag.replace("(", r"\(").replace(")", r"\)")):
            g.RNA("x", seq, dbn)
    with pytest.raises(g.InputValidationError, match="empty identifier"):
        g.RNA("  ", "A", ".")
    with pytest.raises(g.InputValidationError, match="tabs"):
        g.RNA("a\tb", "A", ".")


def test_rna_windows():
    r = g.RNA("x", _SEQ, _DBN, start=9, end=16)
    assert r.sliced and r.core_length == 7
    for kw in (dict(start=3), dict(start=5, end=5), dict(start=-1, end=2), dict(start=0, end=17),
               dict(start=True, end=3), dict(start=1.0, end=3)):
        with pytest.raises(g.InputValidationError):
            g.RNA("x", _SEQ, _DBN, **kw)
    many = g.RNA.many_from_mapping({"transcript_id": "stem", "sequence": _SEQ,
                                    "secondary_structure": _DBN, "start": "6,9", "end": "12, 16"})
    assert [m.identifier for m in many] == ["stem:6-12", "stem:9-16"]
    assert parse_position_list(" 1, 2 ", name="start") == [1, 2] and parse_position_list(None, name="s") == []
    with pytest.raises(g.InputValidationError, match="1 value"):
        g.RNA.many_from_mapping({"transcript_id": "s", "sequence": _SEQ, "secondary_structure": _DBN,
                                 "start": "1", "end": "3,4"})


def test_read_rna_table(tmp_path):
    p = tmp_path / "t.tsv"
    p.write_text("transcript_id\tsequence\tsecondary_structure\na\tACGU\t....\nb\tGGAA\t(())\n")
    recs = g.read_rna_table(p)
    assert [r.identifier for r in recs] == ["a", "b"]
    p.write_text("transcript_id\tsequence\tsecondary_structure\na\tACGU\t....\na\tGGAA\t(())\n")
    with pytest.raises(g.InputValidationError, match="duplicate"):
        g.read_rna_table(p)


# ---- GraphSpec ---------------------------------------------------------------
def test_graph_spec_fingerprint_is_the_reference_one(golden_meta):
    spec = g.GraphSpec.bundled()
    assert spec.sha256 == golden_meta["graph_spec_sha256"]
    assert spec.sha256 == "da2e670e377e47667fec8a8ebb1c90c6e506b9cdd8a5555a6bfab50b202fb9bd"
    assert spec.node_feature_dim == 7 and len(spec.edge_types) == 6
    assert g.GraphSpec.from_dict(spec.to_dict()) == spec
    other = g.GraphSpec(struct_feature="B")
    assert other.sha256 != spec.sha256 and other.node_feature_dim == 9
    with pytest.raises(g.GraphValidationError):
        g.GraphSpec(edge_dim=5)
    with pytest.raises(g.GraphValidationError):
        g.GraphSpec(extra_edges=("skip3",))
    with pytest.raises(g.GraphValidationError, match="inconsistent"):
        g.GraphSpec.from_dict({**spec.to_dict(), "node_feature_dimension": 8})


# ---- builder -------------------------------------------------------------------
def test_graph_builder_exposes_compact_model_versioned_arrays():
    graph = g.GraphBuilder().build(_records()[0])
    assert graph.node_features.shape == (8, 7) and graph.node_features.dtype == np.float32
    assert graph.edge_index.shape == (2, 30) and graph.edge_index.dtype == np.int32
    assert graph.edge_types.shape == (30,) and graph.edge_types.dtype == np.uint8
    np.testing.assert_array_equal(graph.residue_index, np.arange(8, dtype=np.int32))
    assert graph.node_count == graph.core_count == 8
    assert graph.spec.sha256 == g.GraphSpec.bundled().sha256


def test_builder_is_bit_exact_with_the_reference_arrays(golden_shard, golden_graphs):
    for name in ARRAYS:
        got, want = getattr(golden_shard, name), golden_graphs["full/" + name]
        assert got.dtype == want.dtype and got.shape == want.shape
        assert np.array_equal(got, want), name


def test_batched_shard_equals_per_record_graphs(golden_meta):
    records = [g.RNA(*t) for t in golden_meta["full"]]
    builder = g.GraphBuilder()
    a = builder.build_shard(records)
    b = g.GraphShard.from_graphs(builder.build_many(records))
    for name in ARRAYS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_builder_matches_loop_oracle_on_random_structures():
    for rec in random_records(5, 40, mean=90, sd=40, lo=1, hi=300):
        graph = g.GraphBuilder().build(rec)
        x, ei, et = O.build_full_graph(rec.sequence, rec.structure)
        assert np.array_equal(graph.node_features, x)
        assert np.array_equal(graph.edge_index, ei) and np.array_equal(graph.edge_types, et)
        assert np.array_equal(graph_module._pair_table(rec.structure), O.pair_table(rec.structure))


def test_struct_feature_b_layout():
    spec = g.GraphSpec(struct_feature="B")
    graph = g.GraphBuilder(spec).build(g.RNA("r", "ACGU", "(..)"))
    assert graph.node_features.shape == (4, 9)
    assert graph.node_features[:, 4:7].tolist() == [[1, 0, 0], [0, 1, 0], [0, 1, 0], [0, 0, 1]]


def test_windows_match_reference_known_answers_and_arrays(golden_meta, golden_graphs):
    graph = g.GraphBuilder().build(g.RNA("stem", _SEQ, _DBN, start=9, end=16))
    assert graph.residue_index.tolist() == list(range(9, 16)) and graph.core_span == (9, 16)
    assert graph.node_count == graph.core_count == 7
    graph = g.GraphBuilder(keep_paired_neighbours=True).build(g.RNA("stem", _SEQ, _DBN, start=9, end=16))
    assert graph.residue_index.tolist() == list(range(6, 16))
    assert graph.node_roles.tolist() == [1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
    graph = g.GraphBuilder(keep_paired_neighbours=True, context_hops=3).build(
        g.RNA("stem", _SEQ, _DBN, start=9, end=16))
    assert graph.residue_index.tolist() == list(range(2, 16))
    with pytest.raises(ValueError, match="context_hops"):
        g.GraphBuilder(context_hops=0)
    for k, w in enumerate(golden_meta["windows"]):
        r = w["record"]
        shard = g.GraphBuilder(keep_paired_neighbours=w["keep"], context_hops=w["hops"]).build_shard(
            [g.RNA(r[0], r[1], r[2], start=r[3], end=r[4])])
        for name in ARRAYS:
            assert np.array_equal(getattr(shard, name), golden_graphs[f"window{k}/{name}"]), (k, name)


def test_sliced_node_features_match_the_full_molecule():
    full = g.GraphBuilder().build(g.RNA("stem", _SEQ, _DBN))
    part = g.GraphBuilder(keep_paired_neighbours=True, context_hops=2).build(
        g.RNA("stem", _SEQ, _DBN, start=9, end=16))
    np.testing.assert_array_equal(part.node_features, full.node_features[part.residue_index])


@pytest.mark.skipif(not Path("/root/reference/tests/rouskin_sample_6k.tsv").is_file(),
                    reason="reference checkout not present")
def test_full_rouskin_file_hashes_match_the_reference(golden_meta):
    """BASELINE config 1 input: every array of the 5,840-record shard is
    byte-identical to what the reference builds (hashes recorded by
    oracle/make_golden.py)."""
    table = g.read_rna_table("/root/reference/tests/rouskin_sample_6k.tsv")
    shard = g.GraphBuilder().build_shard(table)
    info = golden_meta["rouskin"]
    assert (shard.record_count, shard.node_count, shard.edge_count) == (
        info["records"], info["nodes"], info["edges"])
    for name, digest in info["sha256"].items():
        got = hashlib.sha256(np.ascontiguousarray(getattr(shard, name)).tobytes()).hexdigest()
        assert got == digest, name
    lengths, ecounts = np.diff(shard.node_ptr).tolist(), np.diff(shard.edge_ptr).tolist()
    assert O.pack_microbatches(lengths, ecounts, 60_000, 300_000).tolist() == info["bounds_default"]
    assert len(info["bounds_default"]) - 1 == 15


# ---- shard validation / slicing ---------------------------------------------------
def test_shard_rejects_bad_inputs():
    shard = g.GraphBuilder().build_shard(_records())
    fields = {name: getattr(shard, name) for name in (
        "identifiers", "sequences", "structures", "node_features", "edge_index", "edge_types",
        "node_ptr", "edge_ptr", "spec", "residue_index", "node_roles")}

    def broken(**kw):
        return g.GraphShard(**{**fields, **kw})

    with pytest.raises(g.GraphValidationError, match="duplicate"):
        broken(identifiers=("a", "a"))
    with pytest.raises(g.GraphValidationError, match="node_ptr"):
        broken(node_ptr=shard.node_ptr.astype(np.int32))
    with pytest.raises(g.GraphValidationError, match="offsets"):
        broken(node_ptr=np.array([0, 8, 8], np.int64))
    bad = shard.edge_index.copy(); bad[0, 0] = 16
    with pytest.raises(g.GraphValidationError, match="outside shard node range"):
        broken(edge_index=bad)
    with pytest.raises(g.GraphValidationError, match="edge type"):
        broken(edge_types=np.full_like(shard.edge_types, 10))
    ri = shard.residue_index.copy(); ri[3] = 1
    with pytest.raises(g.GraphValidationError, match="strictly increasing"):
        broken(residue_index=ri)
    ri = shard.residue_index.copy(); ri[7] = 8
    with pytest.raises(g.GraphValidationError, match="outside source sequence"):
        broken(residue_index=ri)
    roles = shard.node_roles.copy(); roles[8:] = 1
    with pytest.raises(g.GraphValidationError, match="no core"):
        broken(node_roles=roles)
    roles = shard.node_roles.copy(); roles[0] = 2
    with pytest.raises(g.GraphValidationError, match="unknown node role"):
        broken(node_roles=roles)
    with pytest.raises(g.GraphValidationError, match="empty"):
        g.GraphShard.from_graphs([]) if False else broken(identifiers=(), sequences=(), structures=())
    with pytest.raises(g.GraphValidationError, match="without graphs"):
        g.GraphShard.from_graphs([])
    # an edge that leaves its graph passes metadata validation but not "full"
    cross = shard.edge_index.copy(); cross[1, 0] = 9
    with pytest.raises(g.GraphValidationError, match="crosses"):
        broken(edge_index=cross).validate_values()


def test_shard_slice_rebases_like_the_reference(golden_shard):
    sub = golden_shard.slice(7, 12)
    n0, e0 = int(golden_shard.node_ptr[7]), int(golden_shard.edge_ptr[7])
    assert sub.record_count == 5 and sub.node_ptr[0] == 0 and sub.edge_ptr[0] == 0
    assert np.array_equal(sub.edge_index, golden_shard.edge_index[:, e0:e0 + sub.edge_count] - n0)
    assert sub.lengths == golden_shard.lengths[7:12] and sub.edge_counts == golden_shard.edge_counts[7:12]
    assert sub.core_counts == sub.lengths
    with pytest.raises(IndexError):
        golden_shard.slice(3, 3)


# ---- shard files -------------------------------------------------------------------
def test_graph_shard_round_trip(tmp_path):
    original = g.GraphBuilder().build_shard(_records())
    tensor_path = tmp_path / "graphs.safetensors"
    _, metadata_path = g.save_graph_shard(original, tensor_path)
    assert metadata_path == g.graph_metadata_path(tensor_path)
    assert "tensor_sha256" not in json.loads(metadata_path.read_text())
    restored = g.load_graph_shard(tensor_path, validation="full")
    assert restored.identifiers == original.identifiers and restored.spec == original.spec
    for name in ARRAYS:
        np.testing.assert_array_equal(getattr(restored, name), getattr(original, name))
    with safe_open(str(tensor_path), framework="np") as handle:
        assert set(handle.keys()) == {"node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr"}
        assert handle.metadata()["format"] == g.GRAPH_SHARD_FORMAT


def test_sliced_shards_store_and_restore_node_metadata(tmp_path):
    recs = [g.RNA("stem", _SEQ, _DBN, start=9, end=16)]
    shard = g.GraphBuilder(keep_paired_neighbours=True, context_hops=2).build_shard(recs)
    path = tmp_path / "w.safetensors"
    g.save_graph_shard(shard, path)
    with safe_open(str(path), framework="np") as handle:
        assert {"residue_index", "node_roles"} <= set(handle.keys())
    back = g.load_graph_shard(path)
    assert np.array_equal(back.node_roles, shard.node_roles)
    assert np.array_equal(back.residue_index, shard.residue_index)


def test_checksum_and_header_checks(tmp_path, monkeypatch):
    path = tmp_path / "graphs.safetensors"
    g.save_graph_shard(g.GraphBuilder().build_shard(_records()), path, checksum=True)
    assert g.load_graph_shard(path, verify_checksum=True).record_count == 2
    payload = bytearray(path.read_bytes()); payload[-1] ^= 1; path.write_bytes(payload)
    with pytest.raises(g.GraphValidationError, match="checksum"):
        g.load_graph_shard(path, verify_checksum=True)
    # hashing stays opt-in
    monkeypatch.setattr(graph_module, "_file_sha256",
                        lambda _p: (_ for _ in ()).throw(AssertionError("hashing must be opt-in")))
    path2 = tmp_path / "plain.safetensors"
    g.save_graph_shard(g.GraphBuilder().build_shard(_records()), path2)
    assert g.load_graph_shard(path2).record_count == 2
    with pytest.raises(g.GraphValidationError, match="no stored checksum"):
        monkeypatch.undo()
        g.load_graph_shard(path2, verify_checksum=True)


def test_loader_rejects_foreign_or_inconsistent_files(tmp_path):
    shard = g.GraphBuilder().build_shard(_records())
    path = tmp_path / "graphs.safetensors"
    _, meta_path = g.save_graph_shard(shard, path)
    other = g.GraphSpec(struct_feature="B")
    with pytest.raises(g.GraphCompatibilityError, match="incompatible"):
        g.load_graph_shard(path, expected_spec=other)
    meta = json.loads(meta_path.read_text())
    meta_path.write_text(json.dumps({**meta, "node_count": 3}))
    with pytest.raises(g.GraphValidationError, match="count metadata"):
        g.load_graph_shard(path)
    meta_path.write_text(json.dumps({**meta, "format": "something-else"}))
    with pytest.raises(g.GraphValidationError, match="unsupported graph shard format"):
        g.load_graph_shard(path)
    meta_path.write_text(json.dumps(meta))
    tensors = {name: getattr(shard, name) for name in ("node_features", "edge_index", "edge_types",
                                                       "node_ptr", "edge_ptr")}
    save_file({**tensors, "extra": np.zeros(1, np.float32)}, str(path),
              metadata={"format": g.GRAPH_SHARD_FORMAT, "format_version": "1",
                        "graph_spec_sha256": shard.spec.sha256})
    with pytest.raises(g.GraphValidationError, match="unexpected"):
        g.load_graph_shard(path)
    save_file(tensors, str(path), metadata={"format": "nope"})
    with pytest.raises(g.GraphValidationError, match="header"):
        g.load_graph_shard(path)
    with pytest.raises(ValueError, match="validation"):
        g.load_graph_shard(path, validation="paranoid")


def test_legacy_shard_without_roles_loads_as_all_core(tmp_path):
    shard = g.GraphBuilder().build_shard(_records())
    path = tmp_path / "legacy.safetensors"
    g.save_graph_shard(shard, path)            # full molecules: optional tensors omitted
    back = g.load_graph_shard(path)
    assert back.all_core and np.array_equal(back.residue_index, shard.residue_index)


# ---- partitioning ---------------------------------------------------------------------
def test_record_partitioning_respects_record_and_node_limits():
    records = [g.RNA(str(i), "ACGU", "....") for i in range(5)]
    parts = list(g.partition_records(records, max_records=3, max_nodes=8))
    assert [[r.identifier for r in p] for p in parts] == [["0", "1"], ["2", "3"], ["4"]]
    assert [len(p) for p in g.partition_records(records, max_records=2)] == [2, 2, 1]
    with pytest.raises(ValueError, match="exceeds max_nodes"):
        list(g.partition_records(records, max_records=2, max_nodes=3))
    with pytest.raises(ValueError, match="positive"):
        list(g.partition_records(records, max_records=0))


# ---- host side of the device graph builder (K6) ---------------------------------------
def test_position_table_reproduces_the_builder_columns_bit_for_bit():
    """The device builder gathers sin/cos from a per-length table made on the
    host; it must hold exactly what the host builder (and the reference,
    graph.py:510-514) computes per record."""
    pytest.importorskip("torch")
    import ginfinity_b200 as g
    try:
        from ginfinity_b200 import device_builder as db
    except ImportError as exc:                      # libgfx.so not built
        pytest.skip(str(exc))
    records = random_records(11, 300) + [g.RNA("one", "A", "."), g.RNA("two", "AC", "..")]
    shard = g.GraphBuilder().build_shard(records)
    lengths = np.diff(shard.node_ptr)
    table, offset = db.position_table(lengths)
    assert table.dtype == np.float32 and offset.dtype == np.int64
    assert table.shape[0] == int(np.unique(lengths).sum())
    rows = np.concatenate([np.arange(o, o + n) for o, n in zip(offset.tolist(), lengths.tolist())])
    assert np.array_equal(table[rows].view(np.uint32),
                          np.ascontiguousarray(shard.node_features[:, 5:7]).view(np.uint32))
    assert db.supports(g.GraphSpec(), records)
    windowed = records + [g.RNA("w", "ACGUACGU", "((....))", start=2, end=5)]
    assert db.supports(g.GraphSpec(), windowed) and db.any_sliced(windowed)
    assert not db.any_sliced(records)
    assert not db.supports(g.GraphSpec(positional=False), records)


def test_shard_files_load_through_a_caller_supplied_allocator(tmp_path):
    """`load_graph_shard(allocate=)`: the safetensors file is memory-mapped and every tensor is copied
    once into the buffer the caller hands out (page-locked pools on the GPU box); the result equals
    the ordinary loader's array for array, header problems are still GraphValidationErrors, and
    the node total can be read from the header alone."""
    import ginfinity_b200 as g
    from ginfinity_b200.multi_gpu import shard_file_node_count
    records = [g.RNA("rna-1", "ACGUACGU", "((....))"), g.RNA("rna-2", "GGAACCUU", "........"),
               g.RNA("stem", "GGGAAACCCUUUUGGG", "......(((....)))", start=9, end=16)]
    shard = g.GraphBuilder(keep_paired_neighbours=True, context_hops=2).build_shard(records)
    path = tmp_path / "s.safetensors"
    g.save_graph_shard(shard, path)
    handed = {}

    def allocate(name, shape, dtype):
        handed[name] = np.full(shape, 7, dtype)
        return handed[name]

    plain = g.load_graph_shard(path)
    pooled = g.load_graph_shard(path, allocate=allocate, validation="full")
    names = ("node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr", "residue_index",
             "node_roles")
    for name in names:
        assert np.array_equal(getattr(plain, name), getattr(pooled, name)), name
        assert getattr(pooled, name) is handed[name]          # no second copy
    assert pooled.identifiers == plain.identifiers and pooled.spec.sha256 == plain.spec.sha256
    assert shard_file_node_count(path) == shard.node_count
    # a truncated file is rejected, not read past its end
    broken = tmp_path / "broken.safetensors"
    broken.write_bytes(path.read_bytes()[:-16])
    (tmp_path / "broken.json").write_bytes(g.graph_metadata_path(path).read_bytes())
    with pytest.raises(g.GraphValidationError, match="cannot load graph shard tensors"):
        g.load_graph_shard(broken, metadata_path=tmp_path / "broken.json", allocate=allocate)
