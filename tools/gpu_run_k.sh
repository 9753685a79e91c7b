#!/bin/bash
# round-2 GPU session K: stream priorities (main chain / re-pack stream)
set -x
O=gpurun_out
run() { SG_MAIN_PRIO=$1 SG_PACK_PRIO=$2 python bench.py --no-cpu-baseline --steps 20 > $O/k_bench_m$1_p$2.log 2>> $O/k.err; }
run -1 -1
run -1 0
run -1 -2
run -2 -2
run 0 -1
