"""Shared helpers for the test-suite (small synthetic RNA sets)."""
import ginfinity_b200 as g
from ginfinity_b200.synthetic import synthetic_records


def random_records(seed, count, mean=200, sd=30, lo=50, hi=400, prefix="syn"):
    """Validated RNA records (goes through the RNA input contract)."""
    return [g.RNA(r.identifier, r.sequence, r.structure)
            for r in synthetic_records(seed, count, mean=mean, sd=sd, lo=lo, hi=hi,
                                       prefix=prefix, workers=1)]
