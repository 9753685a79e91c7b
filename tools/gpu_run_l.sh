#!/bin/bash
# round-2 GPU session L: stream priorities as defaults -- tests, bench, PDL re-test
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/l_tests.log 2>&1; tail -4 $O/l_tests.log
python bench.py > $O/l_bench.log 2> $O/l_bench.err; tail -c 300 $O/l_bench.err
SG_PDL_S1=1 python bench.py --no-extras --no-cpu-baseline > $O/l_bench_pdl1.log 2>&1
