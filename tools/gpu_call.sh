#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for d in csr edges; do
GFX_DESCRIBE=$d timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-records-e2e > gpurun_out/bench_desc_$d.json 2> gpurun_out/bench_desc_$d.err; echo "bench $d rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_desc_$d.json'))
print('$d', d['value'], d['ms_per_step'], d['stage_ms_per_step'], d['roofline']['frac'], d['parity_check']['max_abs'], d['e2e']['value'])
PY
done
