#!/bin/bash
for i in 1 2 3; do
GFX_LIBRARY=$PWD/ginfinity_b200/libgfx_prev.so python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1 | sed "s/^/prev /"
python tools/fused_probe.py gfx_layer_fused_banded 2>&1 | tail -1 | sed "s/^/new  /"
done
