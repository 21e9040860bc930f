#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python tools/fused_trace.py banded > gpurun_out/trace_v8b.log 2>&1; echo "trace rc=$?"
python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-records-e2e > gpurun_out/bench_v8b.json 2> gpurun_out/bench_v8b.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_v8b.json'))
print(d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['parity_check'])
PY
