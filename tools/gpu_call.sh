#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
