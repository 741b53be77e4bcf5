#!/bin/bash
# One GPU-box round: tests, bench (bf16 + f32), ncu launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
for DT in bf16 f32; do
  timeout 900 python bench.py --steps 10 --warmup 3 --dtype $DT > gpurun_out/bench_$DT.json 2> gpurun_out/bench_$DT.err; tail -2 gpurun_out/bench_$DT.err; cut -c1-700 gpurun_out/bench_$DT.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bf16.csv python bench.py --steps 1 --warmup 3 --dtype bf16 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
