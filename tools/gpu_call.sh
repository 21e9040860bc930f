#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q -x -k "banded_fp32 or full_precision" 2>&1 | tail -8
echo "== this build";     timeout 200 python tools/stage_probe.py 100000 fp32
echo "== GFX_DESCRIBE=csr (CSR kernel)"; GFX_DESCRIBE=csr timeout 200 python tools/stage_probe.py 100000 fp32
} > gpurun_out/call.log 2>&1
tail -40 gpurun_out/call.log | cut -c1-300
