#!/bin/bash
bash tools/gpu_round.sh > gpurun_out/round.log 2>&1
tail -12 gpurun_out/round.log | cut -c1-300
