#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
for v in 1 0; do
GFX_BANDED_V7=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-records-e2e > gpurun_out/bench_v7_$v.json 2> gpurun_out/bench_v7_$v.err; echo "bench v7=$v rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_v7_$v.json'))
print('v7=$v', d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['parity_check'])
PY
done
