#!/bin/bash
# One GPU-box round: tests, bench, ncu launch list, ncu full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
DT=${1:-f32}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --dtype $DT > gpurun_out/bench_$DT.json 2> gpurun_out/bench_$DT.err; tail -3 gpurun_out/bench_$DT.err; cat gpurun_out/bench_$DT.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$DT.csv python bench.py --steps 1 --warmup 3 --dtype $DT --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
