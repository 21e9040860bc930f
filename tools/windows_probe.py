#!/usr/bin/env python
"""Developer tool: throughput of windowed records through encode_many (bench.windows_bench);
GFX_LIBRARY=<other build> times another build on the same board."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ginfinity_b200.encoder import Ginfinity  # noqa: E402

state, _ = bench.load_weights()
enc = Ginfinity.from_state(state, device="cuda:0")
out = bench.windows_bench(enc, int(sys.argv[1]) if len(sys.argv) > 1 else 20_000)
print(json.dumps({k: v for k, v in out.items() if k != "workload"}))
