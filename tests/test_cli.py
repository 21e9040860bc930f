"""`python -m ginfinity_b200 ...`: the reference's command-line tests
(tests/test_encoder_cli.py in the reference) restated.  Table -> shard needs no GPU; the
embedding commands run on cuda:0 and are checked against the reference-recorded goldens."""
import json

import numpy as np
import pytest

from conftest import staged_model_dir
from ginfinity_b200.cli import build_parser, main

TABLE = ("transcript_id\tsequence\tsecondary_structure\n"
         "rna-1\tACGUACGU\t((....))\n"
         "rna-2\tGGAACCUU\t........\n")


def _needs_gpu_and_model():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if staged_model_dir() is None:
        pytest.skip("reference checkpoint not staged")


# ---- host only ----------------------------------------------------------------------
def test_parser_keeps_the_reference_flags_and_defaults():
    args = build_parser().parse_args(["embed", "--input", "a.tsv", "--output", "b.npz"])
    assert (args.max_batch_nodes, args.max_batch_edges) == (60_000, 300_000)
    assert args.embedding_dtype == "float16" and args.delimiter == "\t"
    assert (args.id_column, args.sequence_column, args.structure_column) == (
        "transcript_id", "sequence", "secondary_structure")
    assert (args.start_column, args.end_column, args.no_slices) == ("start", "end", False)
    assert args.context_hops is None and args.keep_paired_neighbours is False
    assert args.device == "cuda"            # the one deliberate difference: no CPU path
    args = build_parser().parse_args(["embed-graphs", "--input", "g.safetensors", "--output", "e.npz",
                                      "--verify-checksum", "--full-validation", "--checksum"])
    assert args.verify_checksum and args.full_validation and args.checksum


def test_build_graphs_writes_a_shard_the_loader_accepts(tmp_path, capsys):
    import ginfinity_b200 as g
    source = tmp_path / "molecules.tsv"
    source.write_text(TABLE)
    graphs, sidecar = tmp_path / "graphs.safetensors", tmp_path / "graphs.json"
    assert main(["build-graphs", "--input", str(source), "--output", str(graphs),
                 "--metadata", str(sidecar), "--checksum"]) == 0
    said = json.loads(capsys.readouterr().out)
    assert (said["records"], said["nodes"], said["edges"]) == (2, 16, 30 + 26)
    assert said["graph_spec_sha256"] == g.GraphSpec().sha256 and said["checksum"] is True
    shard = g.load_graph_shard(graphs, metadata_path=sidecar, verify_checksum=True, validation="full")
    want = g.GraphBuilder().build_shard([g.RNA("rna-1", "ACGUACGU", "((....))"),
                                         g.RNA("rna-2", "GGAACCUU", "........")])
    assert shard.identifiers == want.identifiers
    assert np.array_equal(shard.edge_index, want.edge_index)
    assert np.array_equal(shard.node_features, want.node_features)


def test_build_graphs_windows_and_context_flags(tmp_path, capsys):
    import ginfinity_b200 as g
    source = tmp_path / "w.tsv"
    source.write_text("transcript_id\tsequence\tsecondary_structure\tstart\tend\n"
                      "stem\tGGGAAACCCUUUUGGG\t......(((....)))\t9\t16\n")
    out = tmp_path / "w.safetensors"
    assert main(["build-graphs", "--input", str(source), "--output", str(out), "--context-hops", "2"]) == 0
    capsys.readouterr()
    shard = g.load_graph_shard(out)
    want = g.GraphBuilder(keep_paired_neighbours=True, context_hops=2).build_shard(
        [g.RNA("stem:9-16", "GGGAAACCCUUUUGGG", "......(((....)))", start=9, end=16)])
    assert list(shard.identifiers) == list(want.identifiers) == ["stem:9-16"]   # the table reader suffixes windows
    assert np.array_equal(shard.node_roles, want.node_roles)
    assert np.array_equal(shard.edge_index, want.edge_index)
    # --no-slices ignores the window columns
    full = tmp_path / "full.safetensors"
    assert main(["build-graphs", "--input", str(source), "--output", str(full), "--no-slices"]) == 0
    assert g.load_graph_shard(full).node_count == 16


def test_errors_become_exit_status_2(tmp_path, capsys):
    missing = tmp_path / "nope.tsv"
    assert main(["build-graphs", "--input", str(missing), "--output", str(tmp_path / "o.safetensors")]) == 2
    assert "ginfinity_b200:" in capsys.readouterr().err
    source = tmp_path / "molecules.tsv"
    source.write_text(TABLE)
    # no CPU path: the device policy error surfaces as a message, not a traceback
    assert main(["embed", "--input", str(source), "--output", str(tmp_path / "e.npz"),
                 "--device", "cpu"]) == 2
    assert "CUDA" in capsys.readouterr().err


def test_alignment_config(tmp_path, capsys):
    if staged_model_dir() is None:
        pytest.skip("reference checkpoint not staged")
    output = tmp_path / "sub" / "alignment.json"
    assert main(["alignment-config", "--output", str(output)]) == 0
    assert json.loads(output.read_text())["scoring_parameters"]["sigma"] == 1.0
    assert main(["alignment-config"]) == 0
    assert "scoring_parameters" in json.loads(capsys.readouterr().out)


# ---- on the GPU -----------------------------------------------------------------------
@pytest.mark.gpu
def test_embed_writes_embeddings_and_integrity_manifest(tmp_path, golden_meta, golden_embeddings):
    _needs_gpu_and_model()
    source = tmp_path / "molecules.tsv"
    source.write_text("transcript_id\tsequence\tsecondary_structure\n"
                      + "".join("\t".join(t) + "\n" for t in golden_meta["full"]))
    output, manifest = tmp_path / "embeddings.npz", tmp_path / "manifest.json"
    assert main(["embed", "--input", str(source), "--output", str(output),
                 "--manifest", str(manifest)]) == 0
    with np.load(output) as archive:
        assert archive["rna-1"].shape == (8, 128) and archive["rna-1"].dtype == np.float16
        got = np.concatenate([archive[t[0]] for t in golden_meta["full"]]).astype(np.float32)
    ref16 = golden_embeddings["full/fp16_model_f16"].astype(np.float32)
    assert np.abs(got - ref16).max() <= 4e-3          # the reference's own fp16 path
    said = json.loads(manifest.read_text())
    assert said["status"] == "complete" and said["records"][0]["identifier"] == "rna-1"
    assert said["checkpoint_sha256"] == golden_meta["checkpoint_sha256"]
    assert said["records"][0] == {"identifier": "rna-1", "length": 8, "core_length": 8, "shape": [8, 128]}
    assert len(said["input_sha256"]) == 64 and len(said["output_sha256"]) == 64


@pytest.mark.gpu
def test_embed_columns_dtype_and_full_precision(tmp_path, golden_embeddings, golden_meta):
    _needs_gpu_and_model()
    source = tmp_path / "structures.csv"
    source.write_text("name,bases,fold,family\n" +
                      "".join(",".join(t) + ",x\n" for t in golden_meta["full"][:4]))
    output = tmp_path / "embeddings.npz"
    assert main(["embed", "--input", str(source), "--output", str(output), "--id-column", "name",
                 "--sequence-column", "bases", "--structure-column", "fold", "--delimiter", ",",
                 "--embedding-dtype", "float32", "--full-precision"]) == 0
    with np.load(output) as archive:
        got = np.concatenate([archive[t[0]] for t in golden_meta["full"][:4]])
    assert got.dtype == np.float32
    ref = golden_embeddings["full/fp32_model_f32"][: got.shape[0]]
    assert np.abs(got - ref).max() <= 2e-5


@pytest.mark.gpu
def test_build_then_embed_a_graph_shard(tmp_path):
    _needs_gpu_and_model()
    source = tmp_path / "molecules.tsv"
    source.write_text(TABLE)
    graphs, sidecar = tmp_path / "graphs.safetensors", tmp_path / "graphs.json"
    assert main(["build-graphs", "--input", str(source), "--output", str(graphs),
                 "--metadata", str(sidecar), "--checksum"]) == 0
    embeddings, manifest = tmp_path / "embeddings.npz", tmp_path / "embeddings.json"
    assert main(["embed-graphs", "--input", str(graphs), "--metadata", str(sidecar),
                 "--output", str(embeddings), "--manifest", str(manifest), "--verify-checksum",
                 "--full-validation", "--max-batch-nodes", "8", "--checksum"]) == 0
    direct = tmp_path / "direct.npz"
    assert main(["embed", "--input", str(source), "--output", str(direct)]) == 0
    with np.load(embeddings) as a, np.load(direct) as b:
        assert a["rna-1"].shape == (8, 128) and a["rna-2"].shape == (8, 128)
        assert np.array_equal(a["rna-1"], b["rna-1"]) and np.array_equal(a["rna-2"], b["rna-2"])
    said = json.loads(manifest.read_text())
    assert said["status"] == "complete" and len(said["graph_spec_sha256"]) == 64
    assert said["records"][1] == {"identifier": "rna-2", "length": 8, "node_count": 8,
                                  "core_length": 8, "shape": [8, 128]}
    assert len(said["output_sha256"]) == 64


@pytest.mark.gpu
def test_info(capsys):
    _needs_gpu_and_model()
    assert main(["info"]) == 0
    said = json.loads(capsys.readouterr().out)
    assert said["embedding_dimension"] == 128 and len(said["checkpoint_sha256"]) == 64


@pytest.mark.gpu
def test_embed_writes_one_array_per_window_and_no_slices_ignores_them(tmp_path):
    """tests/test_sliced_graphs.py:205-243 of the reference: comma-separated windows become one
    archive entry each (context nodes are computed but not returned); --no-slices encodes the
    whole molecule under the plain identifier."""
    _needs_gpu_and_model()
    sequence, structure = "GGGAAACCCUUUUGGG", "......(((....)))"
    source = tmp_path / "molecules.tsv"
    source.write_text("transcript_id\tsequence\tsecondary_structure\tstart\tend\n"
                      f"stem\t{sequence}\t{structure}\t9,6\t16,12\n")
    output, manifest = tmp_path / "embeddings.npz", tmp_path / "manifest.json"
    assert main(["embed", "--input", str(source), "--output", str(output), "--manifest", str(manifest),
                 "--keep-paired-neighbours", "--context-hops", "2"]) == 0
    with np.load(output) as archive:
        assert set(archive.files) == {"stem:9-16", "stem:6-12"}
        assert archive["stem:9-16"].shape == (7, 128) and archive["stem:6-12"].shape == (6, 128)
    rows = {row["identifier"]: row for row in json.loads(manifest.read_text())["records"]}
    assert (rows["stem:9-16"]["start"], rows["stem:9-16"]["end"], rows["stem:9-16"]["core_length"]) == (9, 16, 7)
    single = tmp_path / "single.tsv"
    single.write_text("transcript_id\tsequence\tsecondary_structure\tstart\tend\n"
                      f"stem\t{sequence}\t{structure}\t9\t16\n")
    whole = tmp_path / "whole.npz"
    assert main(["embed", "--input", str(single), "--output", str(whole), "--no-slices"]) == 0
    with np.load(whole) as archive:
        assert archive.files == ["stem"] and archive["stem"].shape == (16, 128)


def test_parallel_npz_writer_is_numpy_compatible(tmp_path):
    """ginfinity_b200/npz.py writes what numpy.savez_compressed writes (the reference's output
    format, cli.py:86,159), members deflated concurrently: same names in the same order, same
    arrays and dtypes, the same compressed member streams; ".npz" appended when missing; empty
    and 0-row arrays, a non-ASCII name, float16 / float32 / float64; duplicates refused."""
    import struct
    import zipfile

    from ginfinity_b200.npz import write_npz_compressed
    rng = np.random.default_rng(3)
    arrays = [rng.normal(size=(int(n), 128)).astype(dt)
              for n, dt in zip((0, 1, 57, 300, 4096), (np.float16, np.float32, np.float16, np.float64, np.float16))]
    arrays.append(np.zeros((0,), np.float16))
    names = ["r0", "r1", "RNA two", "r3", "largest", "señal"]
    ours = write_npz_compressed(tmp_path / "ours", names, arrays, workers=3)
    assert ours == tmp_path / "ours.npz"
    np.savez_compressed(tmp_path / "numpy.npz", **dict(zip(names, arrays)))
    with np.load(ours) as a, np.load(tmp_path / "numpy.npz") as b:
        assert a.files == b.files == names
        for key in names:
            assert a[key].dtype == b[key].dtype and a[key].shape == b[key].shape
            assert np.array_equal(a[key], b[key])

    def stream(z, info):
        z.fp.seek(info.header_offset + 26)
        n, e = struct.unpack("<HH", z.fp.read(4))
        z.fp.seek(info.header_offset + 30 + n + e)
        return z.fp.read(info.compress_size)

    with zipfile.ZipFile(ours) as za, zipfile.ZipFile(tmp_path / "numpy.npz") as zb:
        assert za.testzip() is None
        for x, y in zip(za.infolist(), zb.infolist()):
            assert (x.filename, x.CRC, x.file_size, x.compress_size) == (y.filename, y.CRC, y.file_size, y.compress_size)
            assert stream(za, x) == stream(zb, y)
    with pytest.raises(ValueError, match="duplicate"):
        write_npz_compressed(tmp_path / "dup.npz", ["a", "a"], arrays[:2])


def test_parallel_npz_writer_beyond_65535_members(tmp_path):
    """More members than a classic ZIP end record can count (a 100k-record shard through
    `embed-graphs`): the ZIP64 end-of-central-directory record is written and numpy reads every
    member back."""
    from ginfinity_b200.npz import write_npz_compressed
    count = 66_000
    arrays = [np.full((1, 4), i, np.float16) for i in range(count)]
    path = write_npz_compressed(tmp_path / "many.npz", (f"m{i}" for i in range(count)), arrays)
    with np.load(path) as z:
        assert len(z.files) == count and z.files[0] == "m0" and z.files[-1] == f"m{count - 1}"
        for i in (0, 1, 65_534, 65_535, 65_536, count - 1):
            assert np.array_equal(z[f"m{i}"], arrays[i])
