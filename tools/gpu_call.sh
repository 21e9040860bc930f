#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
GFX_SCHED=static timeout 120 python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1 | sed "s/^/static  /"
timeout 120 python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1 | sed "s/^/dynamic /"
done
timeout 120 python tools/fused_trace.py banded > gpurun_out/trace_dyn.log 2>&1; tail -3 gpurun_out/trace_dyn.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q --maxfail=8 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
