"""K6: graphs built on the GPU are bit-identical to the host builder's (which
tests/test_graph_host.py and tests/test_oracle_golden.py pin to the
reference), including the reference-recorded golden arrays."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import random_records  # noqa: E402


@pytest.fixture(scope="module")
def gb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ginfinity_b200
    return ginfinity_b200


def compare(gb, records):
    from ginfinity_b200.device_builder import build_device_shard
    want = gb.GraphBuilder().build_shard(records)
    ds = build_device_shard(records, "cuda:0", with_node_metadata=True)
    torch.cuda.synchronize()
    for name in ("node_features", "edge_index", "edge_types", "node_ptr", "edge_ptr"):
        got = getattr(ds, name).cpu().numpy()
        ref = getattr(want, name)
        assert got.dtype == ref.dtype and got.shape == ref.shape, name
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), name   # bit for bit
    assert np.array_equal(ds.residue_index.cpu().numpy(), want.residue_index)
    assert np.array_equal(ds.node_roles_full.cpu().numpy(), want.node_roles)
    assert ds.max_nodes_per_record == int(np.diff(want.node_ptr).max())
    assert ds.max_edges_per_record == int(np.diff(want.edge_ptr).max())
    return ds, want


def test_golden_records_from_the_reference(gb, golden_meta, golden_graphs):
    records = [gb.RNA(*t) for t in golden_meta["full"]]
    ds, want = compare(gb, records)
    assert np.array_equal(ds.edge_index.cpu().numpy(), golden_graphs["full/edge_index"])
    assert np.array_equal(ds.node_features.cpu().numpy().view(np.uint32),
                          golden_graphs["full/node_features"].view(np.uint32))


@pytest.mark.parametrize("seed,count", [(1, 1), (2, 37), (3, 2500)])
def test_random_records_match_the_host_builder(gb, seed, count):
    compare(gb, random_records(seed, count))


def test_tiny_unpaired_and_long_records(gb):
    from ginfinity_b200.synthetic import synthetic_records
    recs = [gb.RNA("one", "A", "."), gb.RNA("two", "AC", ".."), gb.RNA("three", "GAC", "(.)"),
            gb.RNA("flat", "ACGUACGUAC", ".........."), gb.RNA("nest", "GGGGAAAACCCC", "((((....))))")]
    recs += synthetic_records(4, 3, lo=3000, hi=9000, log_uniform=True, workers=1)
    compare(gb, recs)


def test_invalid_input_is_rejected(gb):
    from collections import namedtuple
    from ginfinity_b200.device_builder import build_device_shard
    R = namedtuple("R", "identifier sequence structure")
    with pytest.raises(gb.GraphValidationError, match="unbalanced"):
        build_device_shard([R("a", "ACGU", "((..")], "cuda:0")
    with pytest.raises(gb.GraphValidationError, match="unbalanced"):
        build_device_shard([R("a", "ACGU", "..))")], "cuda:0")
    with pytest.raises(gb.GraphValidationError, match="ACGU"):
        build_device_shard([R("a", "ACGT", "....")], "cuda:0")
    with pytest.raises(gb.GraphValidationError, match="lengths"):
        build_device_shard([R("a", "ACG", "....")], "cuda:0")
    with pytest.raises(gb.GraphValidationError):
        build_device_shard([], "cuda:0")


def test_encode_many_is_the_same_through_either_builder(gb, synthetic_state):
    enc = gb.Ginfinity.from_state(synthetic_state, device="cuda:0")
    recs = random_records(9, 300)
    assert enc.device_builder
    a = enc.encode_many(recs, max_batch_nodes=7000, max_batch_edges=40000)
    enc.device_builder = False
    b = enc.encode_many(recs, max_batch_nodes=7000, max_batch_edges=40000)
    assert len(a) == len(b) == 300
    for x, y in zip(a, b):
        assert x.dtype == np.float16 and np.array_equal(x, y)
    enc.device_builder = True
    with pytest.raises(gb.GraphValidationError, match="duplicate"):
        enc.encode_many([recs[0], recs[0]])
    with pytest.raises(ValueError, match="max_batch_nodes"):
        enc.encode_many(recs, max_batch_nodes=10)
    # windowed records still go through the host builder
    win = gb.RNA("w", "GGGAAACCCUUUUGGG", "......(((....)))", start=9, end=16)
    out = enc.encode_many([win, recs[1]], keep_paired_neighbours=True, context_hops=2)
    assert out[0].shape == (7, 128)
