"""Shared helpers for the test-suite (synthetic RNA shards)."""
import numpy as np


def random_structure(rng, length, pair_fraction=0.57):
    """Random properly nested dot-bracket string, ~`pair_fraction` paired
    (stems of 2-10 pairs, hairpin loops of 3-8, bulges, multiloops)."""
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
        elif rng.random() < 0.5:    # bulge / interior loop then continue
            a, b = int(rng.integers(0, 3)), int(rng.integers(0, 3))
            if inner - a - b < 7:
                a = b = 0
            body = "." * a + gen(inner - a - b, depth + 1) + "." * b
        else:                       # multiloop
            cut = int(rng.integers(4, inner - 3))
            body = gen(cut, depth + 1) + gen(inner - cut, depth + 1)
        if body.count("(") == 0 and len(body) < 3:
            return "." * n
        return "." * lead + "(" * stem + body + ")" * stem + "." * trail
    out = gen(length)
    assert len(out) == length
    return out


def random_records(seed, count, mean=200, sd=30, lo=50, hi=400, prefix="syn"):
    import ginfinity_b200 as g
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        L = int(np.clip(round(rng.normal(mean, sd)), lo, hi))
        seq = "".join(rng.choice(list("ACGU"), size=L))
        out.append(g.RNA(f"{prefix}{i}", seq, random_structure(rng, L)))
    return out
