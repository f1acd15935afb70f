#!/bin/bash
# First GPU call of the next round: everything that was prepared after the round-1 GPU budget ran out, in one go.
#   gpurun --timeout 900 -- 'bash tools/r2_first_call.sh > gpurun_out/r2_first_call.log 2>&1; tail -40 gpurun_out/r2_first_call.log'
set -x
cd "$(dirname "$0")/.."
PRL_TEST_TASKS=1 timeout 600 python -m pytest tests/test_gpu_task_schedule.py -x -q 2>&1 | tail -5
timeout 300 python tools/leduc_ab.py leduc_b5 200
timeout 300 python tools/ab_twocard.py 134459 5 records-only
timeout 300 python bench.py --workload leduc_b5 --no-cpu-baseline
timeout 300 python bench.py --workload leduc_b5 --no-cpu-baseline --schedule tasks
