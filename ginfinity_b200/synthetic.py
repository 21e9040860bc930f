"""Synthetic RNA records and shards of the shapes BASELINE.json names.

There is no network for datasets, so benchmarks and large-scale tests run on
seeded synthetic RNAs: uniform ACGU sequences and random properly nested
secondary structures (stems of 2-10 pairs, hairpin loops, bulges and
multiloops) tuned to the statistics of the reference's sample file
(tests/rouskin_sample_6k.tsv: 56.7 % paired, 4.53 edges per node;
SURVEY.md section 8d).  Output is deterministic in `seed` and independent
of the number of worker processes.
"""
from __future__ import annotations

import os
from collections import namedtuple
from concurrent.futures import ProcessPoolExecutor

import numpy as np

# Duck-typed record: GraphBuilder only needs these attributes.  Skips RNA's
# per-record validation (structures are balanced by construction), which also
# lets the long-RNA workload exceed RNA's 4096-nt cap the way SURVEY 8d (C3)
# prescribes.
SyntheticRecord = namedtuple("SyntheticRecord", "identifier sequence structure")
SyntheticRecord.length = property(lambda self: len(self.sequence))
SyntheticRecord.sliced = False
SyntheticRecord.start = None
SyntheticRecord.end = None

_CHUNK = 2000


def random_structure(rng, length: int) -> str:
    def gen(n, depth=0):
        if n < 7:
            return "." * n
        lead = int(rng.integers(0, 5)) if depth else int(rng.integers(0, 8))
        trail = int(rng.integers(0, 4))
        room = n - lead - trail
        max_stem = min(10, (room - 3) // 2)
        if max_stem < 2:
            return "." * n
        stem = int(rng.integers(2, max_stem + 1))
        inner = room - 2 * stem
        if inner <= 9 or rng.random() < 0.25:
            if inner > 12:          # long hairpin loops are rare: branch instead
                cut = int(rng.integers(4, inner - 3))
                body = gen(cut, depth + 1) + gen(inner - cut, depth + 1)
            else:
                body = "." * inner
        elif rng.random() < 0.5:    # bulge / interior loop, then continue
            a, b = int(rng.integers(0, 3)), int(rng.integers(0, 3))
            if inner - a - b < 7:
                a = b = 0
            body = "." * a + gen(inner - a - b, depth + 1) + "." * b
        else:                       # multiloop
            cut = int(rng.integers(4, inner - 3))
            body = gen(cut, depth + 1) + gen(inner - cut, depth + 1)
        return "." * lead + "(" * stem + body + ")" * stem + "." * trail

    def gen_long(n):
        """Domains of <= 900 nt enclosed by long-range stems (pairs that span
        thousands of nucleotides, like the RF00548-style rRNA folds)."""
        if n <= 900:
            return gen(n)
        stem = int(rng.integers(4, 11))
        inner = n - 2 * stem
        pieces = int(rng.integers(2, 5))
        cuts = np.sort(rng.integers(1, inner, pieces - 1))
        sizes = np.diff(np.concatenate(([0], cuts, [inner]))).tolist()
        return "(" * stem + "".join(gen_long(int(s)) for s in sizes) + ")" * stem

    return gen_long(length)


def _chunk(args):
    seed, chunk_index, lengths, prefix, first = args
    rng = np.random.default_rng([seed, chunk_index])
    letters = np.frombuffer(b"ACGU", dtype=np.uint8)
    out = []
    for k, n in enumerate(lengths):
        seq = letters[rng.integers(0, 4, n)].tobytes().decode("ascii")
        out.append(SyntheticRecord(f"{prefix}{first + k}", seq,
                                   random_structure(rng, int(n))))
    return out


def _lengths(seed, count, mean, sd, lo, hi, log_uniform):
    rng = np.random.default_rng([seed, 0xC0FFEE])
    if log_uniform:
        return np.exp(rng.uniform(np.log(lo), np.log(hi), count)).astype(np.int64)
    return np.clip(np.rint(rng.normal(mean, sd, count)), lo, hi).astype(np.int64)


def synthetic_records(seed: int, count: int, *, mean: float = 200, sd: float = 30,
                      lo: int = 50, hi: int = 400, log_uniform: bool = False,
                      prefix: str = "syn", workers: int | None = None) -> list:
    """`count` records; defaults are BASELINE config 2 (100k x ~200 nt is
    `synthetic_records(0, 100_000)`).  `log_uniform=True, lo=1000, hi=10000`
    gives the long-RNA config 3."""
    lengths = _lengths(seed, count, mean, sd, lo, hi, log_uniform)
    jobs = [(seed, c, lengths[i:i + _CHUNK].tolist(), prefix, i)
            for c, i in enumerate(range(0, count, _CHUNK))]
    if workers is None:
        workers = min(len(jobs), os.cpu_count() or 1, 32)
    if workers <= 1 or len(jobs) == 1:
        chunks = [_chunk(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=workers) as pool:
            chunks = list(pool.map(_chunk, jobs))
    return [r for chunk in chunks for r in chunk]


def synthetic_shard(seed: int, count: int, **kwargs):
    from .graph import GraphBuilder
    return GraphBuilder().build_shard(synthetic_records(seed, count, **kwargs))
