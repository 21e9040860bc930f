for m in -1 3 2 0; do echo "GFX_FUSED=$m"; GFX_FUSED=$m timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x -k "mixed_full" 2>&1 | grep -E "^E|passed|failed" | head -12; done
