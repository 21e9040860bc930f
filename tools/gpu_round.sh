#!/bin/bash
# One GPU session: smoke, tests, bench (both arms), ncu launch list + full capture of the top kernels.
#   tools/gpu_round.sh          everything
#   tools/gpu_round.sh ncu      only the ncu passes (needs gpurun_out/bench.json semantics: re-runs a short bench first)
# The layer kernel is chosen per device by a tuning run; under ncu that run is
# serialised and cold-cache and may choose differently, so the ncu passes pin GFX_FUSED to what the
# plain bench chose on this same board.
mkdir -p gpurun_out
if [ "$1" != "ncu" ]; then
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
else
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-records-e2e > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
cp gpurun_out/bench_short.json gpurun_out/bench_for_pin.json
fi
[ -f gpurun_out/bench_for_pin.json ] || cp gpurun_out/bench.json gpurun_out/bench_for_pin.json
export GFX_FUSED=$(python -c "import json; c = json.load(open('gpurun_out/bench_for_pin.json'))['layer_kernel']['chosen'].lower(); print(3 if 'banded' in c else 2 if 'fused' in c else 0)")
echo "ncu passes pinned to GFX_FUSED=$GFX_FUSED" > gpurun_out/ncu_pin.log
NCU_CMD="python bench.py --steps 1 --warmup 1 --records 20000 --no-cpu-baseline"
timeout 300 $NCU_CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
timeout 300 $NCU_CMD > gpurun_out/ncu_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_banded|fused_pair|umma4_mlp|aggregate_f16|umma2_kernel|input_linear4" -s 7 -c 7 -o gpurun_out/prof_top $NCU_CMD > gpurun_out/ncu_full.log 2>&1
if [ "$1" != "ncu" ]; then
unset GFX_FUSED
SEARCH_CMD="python tools/search_bench.py 3"
timeout 300 $SEARCH_CMD > gpurun_out/search.log 2>&1 && \
SEARCH_BENCH_NO_JSON=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"topk_scan" -s 17 -c 1 -o gpurun_out/prof_search $SEARCH_CMD > gpurun_out/ncu_search.log 2>&1
fi
cat gpurun_out/ncu_pin.log; tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | head -c 1500; tail -3 gpurun_out/bench.err; head -c 600 gpurun_out/bench_ref.json
