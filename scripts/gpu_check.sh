#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both scalar cores), ncu launch list.
# Usage (from the repo root): gpurun --timeout 1500 -- bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
timeout 300 python bench.py --steps 200 --warmup 20 --precision 0 --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
if [ "${1:-}" != "quick" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairs_grouped -s 3 -c 2 -o gpurun_out/prof_grouped \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
fi
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_f64.json; tail -3 gpurun_out/bench_f64.err; cat gpurun_out/bench_f32.json
