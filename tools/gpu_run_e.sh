#!/bin/bash
# round-2 GPU session E: batched BN parameter gradients; final-state launch lists
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/e_tests.log 2>&1; tail -4 $O/e_tests.log
python bench.py > $O/e_bench.log 2> $O/e_bench.err; tail -c 300 $O/e_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file $O/launches_s1e.csv python tools/profile_step.py 128 2 > $O/ncu13.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file $O/launches_s2e.csv python tools/bench_stage2.py 64 1 bf16 eager > $O/ncu14.log 2>&1
