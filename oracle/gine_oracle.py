"""CPU oracle for the GINFINITY encoder hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain NumPy / Python loops, the algorithm of the
reference (nicoaira/GINFINITY 1.2.1) for the path the CUDA library
accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it;
the product package ``ginfinity_b200`` never does.

Parity pinning: the reference's tests hold no stored embedding vectors
(SURVEY.md section 8c), so this oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF, generated in the build container by
``oracle/make_golden.py`` (which imports /root/reference/src) and committed
under ``tests/golden/``; plus the reference's exact integer known-answer
tests (tests/test_sliced_graphs.py:23-69, tests/test_graph.py:21-33),
re-stated in ``tests/test_oracle_golden.py``.  The similarity-search
functions at the bottom have NO reference counterpart: parity unpinned.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5
LN_EPS = 1e-5


# --------------------------------------------------------------------------
# graph construction  (src/ginfinity/graph.py)
# --------------------------------------------------------------------------
def pair_table(structure: str) -> np.ndarray:
    """graph.py:737-747 -- stack matcher."""
    partners = np.full(len(structure), -1, dtype=np.int32)
    stack = []
    for i, ch in enumerate(structure):
        if ch == "(":
            stack.append(i)
        elif ch == ")":
            j = stack.pop()
            partners[j] = i
            partners[i] = j
    return partners


def build_full_graph(sequence: str, structure: str):
    """graph.py:494-561 for the bundled spec (struct 'A', positional, skip2).

    Returns node_features f32 [L,7], edge_index i32 [2,E], edge_types u8 [E].
    """
    L = len(sequence)
    x = np.zeros((L, 7), dtype=np.float32)
    for i, base in enumerate(sequence):
        x[i, "ACGU".index(base)] = 1.0
        x[i, 4] = 0.0 if structure[i] == "." else 1.0
    rel = (np.arange(L, dtype=np.float32) / max(L - 1, 1))[:, None]
    x[:, 5:7] = np.concatenate([np.sin(np.pi * rel), np.cos(np.pi * rel)], axis=1)
    partners = pair_table(structure)
    src, dst, typ = [], [], []
    for i in range(L - 1):
        src.append(i); dst.append(i + 1); typ.append(0)
    for i in range(L - 1):
        src.append(i + 1); dst.append(i); typ.append(1)
    opens = [i for i in range(L) if partners[i] > i]
    for i in opens:
        src.append(i); dst.append(int(partners[i])); typ.append(2)
    for i in opens:
        src.append(int(partners[i])); dst.append(i); typ.append(3)
    for i in range(L - 2):
        src.extend((i, i + 2)); dst.extend((i + 2, i)); typ.extend((4, 5))
    edge_index = (np.asarray([src, dst], dtype=np.int32) if src
                  else np.zeros((2, 0), dtype=np.int32))
    return x, edge_index, np.asarray(typ, dtype=np.uint8)


def select_slice_nodes(L, edge_index, structure, start, end,
                       keep_paired_neighbours, context_hops):
    """graph.py:608-646 -- window + pair partners + BFS hops."""
    core = set(range(start, end))
    selected = set(core)
    if keep_paired_neighbours:
        partners = pair_table(structure)
        frontier = []
        for i in sorted(core):
            p = int(partners[i])
            if p >= 0 and p not in selected:
                selected.add(p)
                frontier.append(p)
        if context_hops > 1 and frontier:
            adj = [[] for _ in range(L)]
            for s, d in zip(edge_index[0], edge_index[1]):
                adj[int(s)].append(int(d))
            for _ in range(context_hops - 1):
                nxt = []
                for node in frontier:
                    for nb in adj[node]:
                        if nb not in selected:
                            selected.add(nb)
                            nxt.append(nb)
                frontier = nxt
                if not frontier:
                    break
    residue = np.asarray(sorted(selected), dtype=np.int32)
    roles = np.where((residue >= start) & (residue < end), 0, 1).astype(np.uint8)
    return residue, roles


def extract_slice(x, edge_index, edge_types, residue):
    """graph.py:649-695 -- induced subgraph, original edge order."""
    remap = {int(old): new for new, old in enumerate(residue)}
    keep = [k for k in range(edge_index.shape[1])
            if int(edge_index[0, k]) in remap and int(edge_index[1, k]) in remap]
    ei = np.asarray([[remap[int(edge_index[0, k])] for k in keep],
                     [remap[int(edge_index[1, k])] for k in keep]],
                    dtype=np.int32).reshape(2, len(keep))
    return x[residue], ei, edge_types[keep]


# --------------------------------------------------------------------------
# microbatch packing  (src/ginfinity/api.py:211-229)
# --------------------------------------------------------------------------
def pack_microbatches(lengths, edge_counts, max_batch_nodes, max_batch_edges):
    """Greedy contiguous first-fit; returns record boundaries [0, ..., B]."""
    B = len(lengths)
    bounds = [0]
    start = 0
    while start < B:
        stop, nodes, edges = start, 0, 0
        while stop < B:
            if stop > start and (nodes + lengths[stop] > max_batch_nodes
                                 or edges + edge_counts[stop] > max_batch_edges):
                break
            nodes += lengths[stop]
            edges += edge_counts[stop]
            stop += 1
        bounds.append(stop)
        start = stop
    return np.asarray(bounds, dtype=np.int64)


# --------------------------------------------------------------------------
# CSR by destination (the order index_add_ sums in: _model.py:44-45)
# --------------------------------------------------------------------------
def csr_by_destination(edge_index, edge_types, num_nodes):
    """Stable sort of edges by destination: within one destination the
    reference's edge order (= its summation order) is kept."""
    dst = edge_index[1].astype(np.int64)
    order = np.argsort(dst, kind="stable")
    row_ptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=num_nodes), out=row_ptr[1:])
    return (row_ptr, edge_index[0][order].astype(np.int32),
            edge_types[order].astype(np.uint8))


# --------------------------------------------------------------------------
# Row descriptors of the banded fused layer (include/gfx.h: gfx_row_describe).
# The PREMISE restated here is the reference's: GraphBuilder._build_full
# (graph.py:494-561) lists backbone fwd, backbone rev, pairs fwd, pairs rev, then
# skip-2 fwd/rev interleaved, so after the stable sort by destination the row of
# nucleotide i reads (i-1, 0) (i+1, 1) [(partner, 2|3)] (i-2, 4) (i+2, 5).
# --------------------------------------------------------------------------
DESC_PREV, DESC_NEXT, DESC_PAIR, DESC_PAIR_REV, DESC_PREV2, DESC_NEXT2 = 1, 2, 4, 8, 16, 32
DESC_GENERIC, DESC_PARTNER_SHIFT, DESC_PARTNER_BITS = 0x80000000, 6, 25


def describe_rows(row_ptr, col_src, col_type):
    """uint32 descriptor per CSR row: which of the five pattern edges the row holds (each
    optional, in that order), pair type and partner index; GENERIC for any other row."""
    n = len(row_ptr) - 1
    out = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        k, end, d = int(row_ptr[i]), int(row_ptr[i + 1]), 0

        def at(src, typ):
            return k < end and col_src[k] == src and col_type[k] == typ

        if at(i - 1, 0):
            d |= DESC_PREV; k += 1
        if at(i + 1, 1):
            d |= DESC_NEXT; k += 1
        if k < end and col_type[k] in (2, 3) and 0 <= col_src[k] < (1 << DESC_PARTNER_BITS):
            d |= DESC_PAIR | (DESC_PAIR_REV if col_type[k] == 3 else 0)
            d |= int(col_src[k]) << DESC_PARTNER_SHIFT
            k += 1
        if at(i - 2, 4):
            d |= DESC_PREV2; k += 1
        if at(i + 2, 5):
            d |= DESC_NEXT2; k += 1
        out[i] = d if k == end else DESC_GENERIC
    return out


# --------------------------------------------------------------------------
# encoder forward, literal form  (src/ginfinity/_model.py:39-46, 65-72)
# --------------------------------------------------------------------------
def forward_literal(state, x, edge_index, edge_types, *, layers=4, edge_dim=10,
                    dtype=np.float32):
    """The reference forward with its own op order: one-hot edge attributes
    through edge_lin, gather, add, ReLU, scatter-add, (1+eps)x + agg,
    Linear, eval BatchNorm, ReLU, Linear, LayerNorm, residual, head."""
    f = lambda name: state[name].astype(dtype)  # noqa: E731
    src, dst = edge_index[0].astype(np.int64), edge_index[1].astype(np.int64)
    attr = np.zeros((edge_types.shape[0], edge_dim), dtype=dtype)
    attr[np.arange(edge_types.shape[0]), edge_types.astype(np.int64)] = 1
    h = x.astype(dtype) @ f("input.weight").T + f("input.bias")
    for l in range(layers):
        p = f"convs.{l}."
        e = attr @ f(p + "edge_lin.weight").T + f(p + "edge_lin.bias")
        m = np.maximum(h[src] + e, 0)
        agg = np.zeros_like(h)
        np.add.at(agg, dst, m)                       # edge order, like index_add_
        z = (dtype(1.0) + f(p + "eps")) * h + agg
        a = z @ f(p + "mlp.0.weight").T + f(p + "mlp.0.bias")
        a = ((a - f(p + "mlp.1.running_mean"))
             / np.sqrt(f(p + "mlp.1.running_var") + dtype(BN_EPS))
             * f(p + "mlp.1.weight") + f(p + "mlp.1.bias"))
        a = np.maximum(a, 0)
        u = a @ f(p + "mlp.4.weight").T + f(p + "mlp.4.bias")
        mu = u.mean(axis=1, keepdims=True)
        var = ((u - mu) ** 2).mean(axis=1, keepdims=True)
        u = (u - mu) / np.sqrt(var + dtype(LN_EPS)) * f(f"norms.{l}.weight") \
            + f(f"norms.{l}.bias")
        h = h + u
    y = np.maximum(h @ f("head.0.weight").T + f("head.0.bias"), 0)
    return y @ f("head.2.weight").T + f("head.2.bias")


# --------------------------------------------------------------------------
# encoder forward, folded form (what the CUDA kernels compute)
# --------------------------------------------------------------------------
def fold_state(state, *, layers=4):
    """Table + BatchNorm fold, float64 then one rounding to float32.
    Follows _model.py:33,35,43 and api.py:243-245 algebraically."""
    d = lambda n: state[n].astype(np.float64)  # noqa: E731
    out = {"w_in": state["input.weight"].astype(np.float32),
           "b_in": state["input.bias"].astype(np.float32)}
    tab, eps1, w1, b1, w2, b2, g, b = ([] for _ in range(8))
    for l in range(layers):
        p = f"convs.{l}."
        tab.append(d(p + "edge_lin.weight").T + d(p + "edge_lin.bias"))
        eps1.append(1.0 + d(p + "eps")[0])
        s = d(p + "mlp.1.weight") / np.sqrt(d(p + "mlp.1.running_var") + BN_EPS)
        w1.append(d(p + "mlp.0.weight") * s[:, None])
        b1.append((d(p + "mlp.0.bias") - d(p + "mlp.1.running_mean")) * s
                  + d(p + "mlp.1.bias"))
        w2.append(d(p + "mlp.4.weight")); b2.append(d(p + "mlp.4.bias"))
        g.append(d(f"norms.{l}.weight")); b.append(d(f"norms.{l}.bias"))
    for key, val in (("table", tab), ("eps1", eps1), ("w1", w1), ("b1", b1),
                     ("w2", w2), ("b2", b2), ("ln_g", g), ("ln_b", b)):
        out[key] = np.asarray(val).astype(np.float32)
    for key, name in (("wa", "head.0.weight"), ("ba", "head.0.bias"),
                      ("wb", "head.2.weight"), ("bb", "head.2.bias")):
        out[key] = state[name].astype(np.float32)
    return out


def _h(a):
    """Round to fp16 storage and come back to fp32 (what an fp16 buffer does)."""
    return a.astype(np.float16).astype(np.float32)


def forward_folded(fw, x, row_ptr, col_src, col_type, *, half_storage=False,
                   layers=4, return_intermediates=False, half_sums=False):
    """Folded forward over a destination-CSR graph, float32 arithmetic.

    ``half_sums=True`` (with ``half_storage``) additionally models what the
    reference's fp16 path does on a ``.half()`` module (_model.py:42-46, 68-71)
    and what the fused sm_100a layer kernels compute: the neighbour sum is a
    chain of fp16 additions in CSR order, the self term one fp16 fused
    multiply-add with (1 + eps) rounded to fp16, and the LayerNorm output is
    rounded to fp16 before the fp16 residual addition.

    ``half_storage=True`` models the device fp16 path: weights, tables and
    every tensor that the kernels store between stages (h, z, the hidden
    activation fed to the second GEMM, the head's hidden activation) and the
    per-edge message relu(h_src + table) (as in the reference's own fp16
    path, _model.py:43) are rounded to fp16; all sums are float32.
    """
    q = _h if half_storage else (lambda a: a)
    n = x.shape[0]
    deg = np.diff(row_ptr).astype(np.int64)
    dst = np.repeat(np.arange(n, dtype=np.int64), deg)
    src = col_src.astype(np.int64)
    typ = col_type.astype(np.int64)
    keep = {}
    h = q(x.astype(np.float32) @ q(fw["w_in"]).T + fw["b_in"])
    keep["h0"] = h
    for l in range(layers):
        m = q(np.maximum(h[src] + q(fw["table"][l])[typ], 0))
        agg = np.zeros_like(h)
        if half_sums:
            for k in range(int(deg.max()) if n else 0):   # k-th edge of every row that has one
                rows = np.flatnonzero(deg > k)
                agg[rows] = q(agg[rows] + m[row_ptr[rows] + k])
            e16 = np.float64(np.float16(fw["eps1"][l]))
            z = q((e16 * h.astype(np.float64) + agg.astype(np.float64)).astype(np.float32))
        else:
            np.add.at(agg, dst, m)                   # CSR order per row
            z = q(fw["eps1"][l] * h + agg)
        a = q(np.maximum(z @ q(fw["w1"][l]).T + fw["b1"][l], 0))
        u = a @ q(fw["w2"][l]).T + fw["b2"][l]
        mu = u.mean(axis=1, keepdims=True)
        var = ((u - mu) ** 2).mean(axis=1, keepdims=True)
        u = (u - mu) / np.sqrt(var + np.float32(LN_EPS)) * fw["ln_g"][l] + fw["ln_b"][l]
        keep[f"z{l}"] = z
        h = q(h + q(u)) if half_sums else q(h + u)
        keep[f"h{l + 1}"] = h
    t = q(np.maximum(h @ q(fw["wa"]).T + fw["ba"], 0))
    y = t @ q(fw["wb"]).T + fw["bb"]
    return (y, keep) if return_intermediates else y


# --------------------------------------------------------------------------
# output stage  (src/ginfinity/api.py:250-259)
# --------------------------------------------------------------------------
def normalise_and_split(y, node_ptr, node_roles, embedding_dtype):
    """float64 L2-normalise with the 1e-12 clamp, keep core rows per record,
    cast to ``embedding_dtype``."""
    e = y.astype(np.float32).astype(np.float64)
    e = e / np.maximum(np.linalg.norm(e, axis=1, keepdims=True), 1e-12)
    out = []
    for i in range(len(node_ptr) - 1):
        a, b = int(node_ptr[i]), int(node_ptr[i + 1])
        core = node_roles[a:b] == 0
        out.append(np.ascontiguousarray(e[a:b][core], dtype=embedding_dtype))
    return out


def encode_shard(fw, shard, *, max_batch_nodes=60_000, max_batch_edges=300_000,
                 embedding_dtype=np.float16, half_storage=False):
    """api.py:180-260 end to end over any object with the GraphShard fields,
    microbatch by microbatch."""
    lengths = np.diff(shard.node_ptr).tolist()
    ecounts = np.diff(shard.edge_ptr).tolist()
    bounds = pack_microbatches(lengths, ecounts, max_batch_nodes, max_batch_edges)
    out = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        n0, n1 = int(shard.node_ptr[a]), int(shard.node_ptr[b])
        e0, e1 = int(shard.edge_ptr[a]), int(shard.edge_ptr[b])
        ei = shard.edge_index[:, e0:e1] - np.int32(n0)
        rp, cs, ct = csr_by_destination(ei, shard.edge_types[e0:e1], n1 - n0)
        y = forward_folded(fw, shard.node_features[n0:n1], rp, cs, ct,
                           half_storage=half_storage)
        out.extend(normalise_and_split(
            y, shard.node_ptr[a:b + 1] - n0, shard.node_roles[n0:n1],
            embedding_dtype))
    return out


# --------------------------------------------------------------------------
# similarity search -- NO REFERENCE COUNTERPART (parity unpinned)
# --------------------------------------------------------------------------
def topk_bruteforce(queries, database, k, metric="cosine"):
    """float64 brute force.  Order: score descending, database index
    ascending on ties.  ``metric`` 'cosine' scores by dot product of the
    given (unit) vectors; 'l2' scores by negative squared distance."""
    q = queries.astype(np.float64)
    d = database.astype(np.float64)
    s = q @ d.T
    if metric == "l2":
        s = -((q * q).sum(1)[:, None] - 2.0 * s + (d * d).sum(1)[None, :])
    elif metric != "cosine":
        raise ValueError(metric)
    k = min(k, d.shape[0])
    idx = np.empty((q.shape[0], k), dtype=np.int64)
    val = np.empty((q.shape[0], k), dtype=np.float64)
    ids = np.arange(d.shape[0])
    for r in range(q.shape[0]):
        order = np.lexsort((ids, -s[r]))[:k]
        idx[r], val[r] = order, s[r][order]
    return val, idx


def exact_scores_fp32(queries, rows, metric="cosine"):
    """The score definition of ginfinity_b200's search (no reference
    counterpart): a sequential fp32 FMA chain over dimensions 0..d-1 of the
    fp16 inputs.  queries [Q,d], rows [R,d] float16 -> float32 [Q,R].

    fma(a, b, acc) in fp32 is emulated as float32(float64(acc) + float64(a) *
    float64(b)): the product of two fp16 (or of an fp32 difference with
    itself) is exact in float64, so apart from astronomically rare double
    roundings this is the correctly rounded single-precision FMA."""
    q = np.asarray(queries, np.float16).astype(np.float32)
    x = np.asarray(rows, np.float16).astype(np.float32)
    acc = np.zeros((q.shape[0], x.shape[0]), np.float32)
    for d in range(q.shape[1]):
        if metric == "cosine":
            a = q[:, d, None].astype(np.float64)
            b = x[None, :, d].astype(np.float64)
        elif metric == "l2":
            diff = (q[:, d, None] - x[None, :, d]).astype(np.float32)
            a = b = diff.astype(np.float64)
        else:
            raise ValueError(metric)
        acc = (acc.astype(np.float64) + a * b).astype(np.float32)
    return acc if metric == "cosine" else -acc


def topk_exact(queries, database, k, metric="cosine", index_base=0):
    """Brute force under the package's own score definition
    (`exact_scores_fp32`), ordered by (score desc, index asc); slots beyond
    the number of rows hold (-inf, -1).  Small inputs only."""
    Q, D = queries.shape[0], database.shape[0]
    val = np.full((Q, k), -np.inf, np.float32)
    idx = np.full((Q, k), -1, np.int64)
    if D == 0 or Q == 0:
        return val, idx
    s = exact_scores_fp32(queries, database, metric)
    ids = np.arange(D)
    kk = min(k, D)
    for r in range(Q):
        order = np.lexsort((ids, -s[r].astype(np.float64)))[:kk]
        idx[r, :kk], val[r, :kk] = order + index_base, s[r][order]
    return val, idx


def merge_topk(all_scores, all_index, k):
    """k-way merge of [parts, Q, k] lists by (score desc, index asc);
    index < 0 marks padding."""
    parts, Q, kk = all_scores.shape
    s = np.transpose(all_scores, (1, 0, 2)).reshape(Q, parts * kk)
    i = np.transpose(all_index, (1, 0, 2)).reshape(Q, parts * kk)
    val = np.full((Q, k), -np.inf, np.float32)
    idx = np.full((Q, k), -1, np.int64)
    for r in range(Q):
        ok = i[r] >= 0
        order = np.lexsort((i[r][ok], -s[r][ok].astype(np.float64)))[:k]
        idx[r, :len(order)], val[r, :len(order)] = i[r][ok][order], s[r][ok][order]
    return val, idx
