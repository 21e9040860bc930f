#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1
timeout 120 python tools/fused_trace.py banded 2>&1 | grep "CTAs:" 
timeout 120 python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1
