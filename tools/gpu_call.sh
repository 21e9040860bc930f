#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -2
GFX_LIBRARY=$PWD/ginfinity_b200/libgfx_prev.so timeout 300 python tools/search_bench.py 3 2>&1 | tail -3 | sed 's/^/prev /'
timeout 300 python tools/search_bench.py 3 2>&1 | tail -3 | sed 's/^/new  /'
