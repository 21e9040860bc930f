#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 gpurun_out/pytest_kernels.log
for i in 1 2 3; do
  GFX_LIBRARY=$PWD/ginfinity_b200/libgfx_prev.so timeout 120 python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1 | sed 's/^/prev /'
  timeout 120 python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1 | sed 's/^/new  /'
done
GFX_SCHED=dynamic timeout 120 python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1 | sed 's/^/dyn  /'
timeout 120 python tools/fused_trace.py banded > gpurun_out/trace.log 2>&1; echo "trace rc=$?"
