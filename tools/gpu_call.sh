#!/bin/bash
mkdir -p gpurun_out
{
echo "== previous build"; GFX_LIBRARY=$PWD/ginfinity_b200/libgfx_prev.so timeout 100 python tools/windows_probe.py
echo "== this build";     timeout 100 python tools/windows_probe.py
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
} > gpurun_out/call.log 2>&1
tail -12 gpurun_out/call.log | cut -c1-400
