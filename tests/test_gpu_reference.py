"""Parity against the REAL reference, run on the GPU box from the staged copy
(oracle/_ref, `python -m oracle.stage_ref`; skipped when nothing is staged).

  * the reference's own test-suite passes with libgfx.so bound under its API
    (INTEGRATION.md's binding, ginfinity_b200/reference_binding.py);
  * all of C1 (BASELINE configs[0]: tests/rouskin_sample_6k.tsv, 5,840 RNAs,
    897,588 nt) and a >= 1 M-nt slice of C2 (configs[1]) encode to the
    reference's embeddings within the north-star tolerances:
      cosine >= 0.999 per nucleotide vs the reference fp32 path,
      max-abs <= 4e-3 vs the reference fp16 path (its own fp16<->fp32 gap is 2e-3),
      max-abs <= 2e-5 for full_precision=True vs the reference fp32 path;
  * shard files written by either package load in the other and encode alike.
"""
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(ref_loader.reference_root() is None,
                                reason="reference not staged (python -m oracle.stage_ref)")

COS_MIN_VS_FP32 = 0.999          # north star: per-nucleotide cosine vs the reference fp32 path
MAX_ABS_VS_FP16 = 4e-3           # north star: max-abs vs the reference fp16 path
MAX_ABS_FP32_PATH = 2e-5         # full_precision=True vs the reference fp32 path


def _cos(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))


@pytest.fixture(scope="module")
def ref():
    return ref_loader.import_reference()[0]


@pytest.fixture(scope="module")
def gb():
    import ginfinity_b200
    return ginfinity_b200


def test_staged_reference_is_byte_identical_to_its_manifest():
    root = ref_loader.reference_root()
    manifest = root / "STAGED_MANIFEST.json"
    if not manifest.is_file():
        pytest.skip("running from the read-only checkout, nothing staged")
    import hashlib
    files = json.loads(manifest.read_text())["files"]
    assert "src/ginfinity/api.py" in files and "src/ginfinity/data/encoder.pt" in files
    for name, digest in files.items():
        assert hashlib.sha256((root / name).read_bytes()).hexdigest() == digest, name


@pytest.mark.gpu
def test_reference_test_suite_passes_through_the_c_abi_binding():
    """INTEGRATION.md's stub, executed: the reference's own tests for this path run against
    the unmodified api.py with `_run_graph_shard` bound to libgfx.so."""
    root = ref_loader.reference_root()
    tests = [str(root / "tests" / name) for name in
             ("test_api.py", "test_graph.py", "test_sliced_graphs.py", "test_encoder_cli.py")]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join((str(ROOT / "tests"), str(ROOT))),
               PYTHONDONTWRITEBYTECODE="1")
    proc = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-p",
         "ref_binding_plugin", "--rootdir", str(root), *tests],
        cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=900)
    tail = proc.stdout[-3000:] + proc.stderr[-2000:]
    assert proc.returncode == 0, tail
    line = re.search(r"gfx-binding: calls=(\d+) nodes=(\d+) libgfx_launches=(\d+)", proc.stdout)
    assert line, tail
    calls, nodes, launches = map(int, line.groups())
    assert calls >= 20 and nodes > 0 and launches > 0, line.group(0)
    passed = re.search(r"(\d+) passed", proc.stdout)
    assert passed and int(passed.group(1)) >= 30, tail


def _reference_embeddings(ref, shard, *, device, full_precision, dtype):
    kw = dict(device=device, full_precision=full_precision)
    if device != "cpu":
        kw["allow_nondeterministic_cuda"] = True
    encoder = ref.Ginfinity.load(**kw)
    return np.concatenate(encoder.encode_graphs(ref_loader.to_reference_shard(ref, shard),
                                                embedding_dtype=dtype))


@pytest.mark.gpu
def test_all_of_c1_matches_the_reference(ref, gb):
    """BASELINE configs[0]: every RNA of rouskin_sample_6k.tsv, reference on its own CPU path
    (fp32 model, the configuration the config names) and its fp16 default."""
    import torch
    records = gb.read_rna_table(ref_loader.rouskin_table())
    assert len(records) == 5840
    shard = gb.GraphBuilder().build_shard(records)
    assert shard.node_count == 897_588 and shard.edge_count == 4_064_014
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    want32 = _reference_embeddings(ref, shard, device="cpu", full_precision=True, dtype=np.float32)
    enc16 = gb.Ginfinity.load("cuda:0")
    got16 = np.concatenate(enc16.encode_graphs(shard, embedding_dtype=np.float32))
    assert got16.shape == want32.shape == (897_588, 128)
    assert _cos(got16, want32).min() >= COS_MIN_VS_FP32
    enc32 = gb.Ginfinity.load("cuda:0", full_precision=True)
    got32 = np.concatenate(enc32.encode_graphs(shard, embedding_dtype=np.float32))
    assert np.abs(got32 - want32).max() <= MAX_ABS_FP32_PATH
    # the records path (graphs built on the GPU) gives the same bits as the shard path
    many = np.concatenate(enc16.encode_many(records, embedding_dtype=np.float32))
    assert np.array_equal(many, got16)
    # reference fp16 model (package default), run through its own eager CUDA path for time
    want16 = _reference_embeddings(ref, shard, device="cuda:0", full_precision=False,
                                   dtype=np.float32)
    assert np.abs(got16 - want16).max() <= MAX_ABS_VS_FP16


@pytest.mark.gpu
def test_a_million_nucleotides_of_c2_match_the_reference(ref, gb):
    """BASELINE configs[1] at >= 1 M nt (first 5,500 records of the bench's own shard, default
    60k/300k limits: 19 microbatches).  The reference runs its eager device="cuda" path here
    (the same torch modules as its CPU path; its CPU path at this size takes minutes)."""
    from ginfinity_b200.synthetic import synthetic_shard
    shard = synthetic_shard(0, 5500)
    assert shard.node_count >= 1_000_000
    want32 = _reference_embeddings(ref, shard, device="cuda:0", full_precision=True,
                                   dtype=np.float32)
    want16 = _reference_embeddings(ref, shard, device="cuda:0", full_precision=False,
                                   dtype=np.float32)
    enc16 = gb.Ginfinity.load("cuda:0")
    got16 = np.concatenate(enc16.encode_graphs(shard, embedding_dtype=np.float32))
    assert _cos(got16, want32).min() >= COS_MIN_VS_FP32
    assert np.abs(got16 - want16).max() <= MAX_ABS_VS_FP16
    enc32 = gb.Ginfinity.load("cuda:0", full_precision=True)
    got32 = np.concatenate(enc32.encode_graphs(shard, embedding_dtype=np.float32))
    assert np.abs(got32 - want32).max() <= MAX_ABS_FP32_PATH
    # default output dtype: float16 arrays, unit norm to fp16 rounding
    half = np.concatenate(enc16.encode_graphs(shard))
    assert half.dtype == np.float16
    assert np.abs(np.linalg.norm(half.astype(np.float32), axis=1) - 1).max() <= 2e-3


def test_shard_files_interchange_with_the_reference(ref, gb, tmp_path):
    """Files written by either package load in the other, array for array
    (reference graph.py:756-923; both validation modes; checksums on and off)."""
    records = [gb.RNA("rna-1", "ACGUACGU", "((....))"), gb.RNA("rna-2", "GGAACCUU", "........"),
               gb.RNA("stem", "GGGAAACCCUUUUGGG", "......(((....)))", start=9, end=16)]
    names = ("node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr",
             "residue_index", "node_roles")
    ours = gb.GraphBuilder(keep_paired_neighbours=True, context_hops=2).build_shard(records)
    theirs = ref.GraphBuilder(keep_paired_neighbours=True, context_hops=2).build_shard(
        [ref.RNA(r.identifier, r.sequence, r.structure, start=r.start, end=r.end)
         for r in records])
    for name in names:
        assert np.array_equal(getattr(ours, name), getattr(theirs, name)), name
    for checksums in (False, True):
        a, b = tmp_path / f"ours{checksums}.safetensors", tmp_path / f"theirs{checksums}.safetensors"
        gb.save_graph_shard(ours, a, checksum=checksums)
        ref.save_graph_shard(theirs, b, checksum=checksums)
        for validation in ("metadata", "full"):
            into_ref = ref.load_graph_shard(a, expected_spec=theirs.spec, validation=validation,
                                            verify_checksum=checksums)
            into_ours = gb.load_graph_shard(b, expected_spec=ours.spec, validation=validation,
                                            verify_checksum=checksums)
            for name in names:
                assert np.array_equal(getattr(into_ref, name), getattr(ours, name)), name
                assert np.array_equal(getattr(into_ours, name), getattr(theirs, name)), name
            assert into_ref.identifiers == into_ours.identifiers == ours.identifiers
            assert into_ours.spec.sha256 == into_ref.spec.sha256
        # the sidecars describe the same shard
        ja = json.loads(gb.graph_metadata_path(a).read_text())
        jb = json.loads(ref.graph_metadata_path(b).read_text())
        for key in ("identifiers", "sequences", "structures", "node_count", "edge_count",
                    "graph_spec_sha256"):
            if key in jb:
                assert ja.get(key) == jb[key], key
