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
    # windowed records: selection and induced subgraphs on the device (K7) vs the host builder
    win = gb.RNA("w", "GGGAAACCCUUUUGGG", "......(((....)))", start=9, end=16)
    for keep, hops in ((False, 1), (True, 1), (True, 2), (True, 3)):
        enc.device_builder = True
        out = enc.encode_many([win, recs[1]], keep_paired_neighbours=keep, context_hops=hops)
        enc.device_builder = False
        ref = enc.encode_many([win, recs[1]], keep_paired_neighbours=keep, context_hops=hops)
        assert out[0].shape == (7, 128) and out[1].shape == ref[1].shape
        for x, y in zip(out, ref):
            assert np.array_equal(x, y)
    enc.device_builder = True


# ---- K7: windowed records ---------------------------------------------------------------
def compare_sliced(gb, records, keep, hops):
    from ginfinity_b200.device_builder import build_device_shard
    want = gb.GraphBuilder(keep_paired_neighbours=keep, context_hops=hops).build_shard(records)
    ds = build_device_shard(records, "cuda:0", keep_paired_neighbours=keep, context_hops=hops)
    torch.cuda.synchronize()
    for name in ("node_ptr", "edge_ptr", "node_features", "edge_index", "edge_types"):
        got = getattr(ds, name).cpu().numpy()
        ref = getattr(want, name)
        assert got.dtype == ref.dtype and got.shape == ref.shape, name
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), name   # bit for bit
    assert np.array_equal(ds.residue_index.cpu().numpy(), want.residue_index)
    assert np.array_equal(ds.node_roles_full.cpu().numpy(), want.node_roles)
    assert ds.core_count == int((want.node_roles == 0).sum())
    assert np.array_equal(np.diff(ds.core_ptr_host), want.core_count_array())
    assert (ds.node_roles is None) == bool(want.all_core)
    return ds, want


def test_golden_windows_from_the_reference(gb, golden_meta, golden_graphs):
    """The reference's own sliced graphs (recorded by oracle/make_golden.py,
    including the known-answer windows of tests/test_sliced_graphs.py:23-69)."""
    from ginfinity_b200.device_builder import build_device_shard
    for k, w in enumerate(golden_meta["windows"]):
        ident, seq, dbn, start, end = w["record"]
        rec = gb.RNA(ident, seq, dbn, start=start, end=end)
        ds = build_device_shard([rec], "cuda:0", keep_paired_neighbours=w["keep"],
                                context_hops=w["hops"])
        torch.cuda.synchronize()
        for name, got in (("node_features", ds.node_features), ("edge_index", ds.edge_index),
                          ("edge_types", ds.edge_types), ("node_ptr", ds.node_ptr),
                          ("edge_ptr", ds.edge_ptr), ("residue_index", ds.residue_index),
                          ("node_roles", ds.node_roles_full)):
            ref = golden_graphs[f"window{k}/{name}"]
            got = got.cpu().numpy()
            assert got.dtype == ref.dtype and got.shape == ref.shape, (k, name)
            assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), (k, name)


def _random_windows(gb, seed, count, lo=20, hi=400):
    from ginfinity_b200.synthetic import synthetic_records
    rng = np.random.default_rng(seed)
    out = []
    for r in synthetic_records(seed, count, lo=lo, hi=hi, workers=1):
        L = len(r.sequence)
        if rng.random() < 0.2:
            out.append(r)                                # full molecules mixed in
            continue
        a = int(rng.integers(0, L))
        b = int(rng.integers(a + 1, min(L, a + 1 + int(rng.integers(1, 120))) + 1))
        out.append(gb.RNA(r.identifier, r.sequence, r.structure, start=a, end=b))
    return out


@pytest.mark.parametrize("keep,hops", [(False, 1), (False, 4), (True, 1), (True, 2), (True, 3),
                                       (True, 7), (True, 300)])
def test_random_windows_match_the_host_builder(gb, keep, hops):
    compare_sliced(gb, _random_windows(gb, 50 + hops, 300), keep, hops)


def test_windows_edge_cases(gb):
    recs = [gb.RNA("one", "A", ".", start=0, end=1),
            gb.RNA("two", "AC", "..", start=1, end=2),
            gb.RNA("pair", "GAC", "(.)", start=0, end=1),
            gb.RNA("whole", "GGGGAAAACCCC", "((((....))))", start=0, end=12),
            gb.RNA("tail", "GGGGAAAACCCC", "((((....))))", start=11, end=12),
            gb.RNA("full", "GGGGAAAACCCC", "((((....))))")]
    for keep, hops in ((False, 1), (True, 1), (True, 2), (True, 5)):
        compare_sliced(gb, recs, keep, hops)
    # long molecules: several block passes per record in the scans
    from ginfinity_b200.synthetic import synthetic_records
    long = synthetic_records(8, 3, lo=3000, hi=9000, log_uniform=True, workers=1)
    from collections import namedtuple
    W = namedtuple("W", "identifier sequence structure start end sliced")   # past RNA's 4096-nt cap
    wins = [W(r.identifier, r.sequence, r.structure, len(r.sequence) // 3,
              len(r.sequence) // 3 + 700, True) for r in long]
    for keep, hops in ((True, 1), (True, 4), (True, 40)):
        compare_sliced(gb, wins, keep, hops)


def test_slice_argument_errors(gb):
    from collections import namedtuple
    from ginfinity_b200.device_builder import build_device_shard
    R = namedtuple("R", "identifier sequence structure start end sliced")
    with pytest.raises(gb.GraphValidationError, match="window"):
        build_device_shard([R("a", "ACGU", "....", 2, 9, True)], "cuda:0")
    with pytest.raises(gb.GraphValidationError, match="unbalanced"):
        build_device_shard([R("a", "ACGU", "((..", 0, 2, True)], "cuda:0")
    with pytest.raises(ValueError, match="context_hops"):
        build_device_shard([R("a", "ACGU", "....", 0, 2, True)], "cuda:0", context_hops=0)
