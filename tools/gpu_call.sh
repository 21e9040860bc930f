#!/bin/bash
python tools/c4_probe.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "shard_files" 2>&1 | tail -3
