#!/bin/bash
mkdir -p gpurun_out
{
echo "== previous build"; GFX_LIBRARY=$PWD/ginfinity_b200/libgfx_prev.so timeout 200 python tools/stage_probe.py 100000 fp32
echo "== this build";     timeout 200 python tools/stage_probe.py 100000 fp32
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py tests/test_gpu_reference.py -m gpu -q -x 2>&1 | tail -8
} > gpurun_out/call.log 2>&1
tail -40 gpurun_out/call.log | cut -c1-300
