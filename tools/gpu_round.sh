#!/bin/bash
# One GPU session: smoke, tests, bench (both arms), ncu launch list + full capture of the top kernels.
#   tools/gpu_round.sh          everything
#   tools/gpu_round.sh ncu      only the ncu passes
# Every ncu pass is preceded by the same command run plain (B200_PROFILING.md); numbers printed by a
# run under ncu are never used, and the tools that write timing files do not write them under ncu.
mkdir -p gpurun_out
if [ "$1" != "ncu" ]; then
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
NCU_CMD="python bench.py --steps 1 --warmup 1 --records 20000 --no-cpu-baseline --no-extras --no-records-e2e"
timeout 300 $NCU_CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
timeout 300 $NCU_CMD > gpurun_out/ncu_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_banded8|head8|input_umma|edge_classify|edge_describe" -s 4 -c 8 -o gpurun_out/prof_top $NCU_CMD > gpurun_out/ncu_full.log 2>&1
# full_precision path: K1 from row descriptors and the split-fp16 K2 / K3 (first pass of the probe skipped)
FP32_CMD="python tools/stage_probe.py 20000 fp32"
timeout 300 $FP32_CMD > gpurun_out/fp32_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"aggregate_f32_band|split_kernel" -s 18 -c 9 -o gpurun_out/prof_fp32 $FP32_CMD > gpurun_out/ncu_fp32.log 2>&1
if [ "$1" != "ncu" ]; then
SEARCH_CMD="python tools/search_bench.py 3"
timeout 300 $SEARCH_CMD > gpurun_out/search.log 2>&1 && \
SEARCH_BENCH_NO_JSON=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"topk_scan" -s 17 -c 1 -o gpurun_out/prof_search $SEARCH_CMD > gpurun_out/ncu_search.log 2>&1
fi
tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | head -c 1500; tail -3 gpurun_out/bench.err; head -c 600 gpurun_out/bench_ref.json
