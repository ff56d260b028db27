#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/dbg_aux.py 2>&1 | grep step
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "engine or fused or pipelined or step or rsgd or rows or joint" > gpurun_out/pytest_step.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_step.log
tail -5 gpurun_out/pytest_step.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 12 --csv --log-file gpurun_out/launches_cfg4.csv \
   python bench.py --workload cfg4 --steps 6 --warmup 3 --pairs 131040 --no-cpu-baseline --no-e2e --rotation 4 > gpurun_out/ncu_cfg4.log 2>&1
grep -o '"void [^"]*"\|"[0-9]*"$' gpurun_out/launches_cfg4.csv | paste - - | tail -8
