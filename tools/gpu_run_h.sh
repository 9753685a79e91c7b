#!/bin/bash
# round-2 GPU session H: per-layer re-pack events, uint8 sampler output, bstats threshold 4000 -- tests + bench + final ncu/SASS evidence
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/h_tests.log 2>&1; tail -4 $O/h_tests.log
python bench.py > $O/h_bench.log 2> $O/h_bench.err; tail -c 300 $O/h_bench.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tcp -s 2 -c 1 -o $O/ncu_conv_tcp_final -f python tools/bench_conv.py s1.D1.ds3 3 fprop > $O/ncu20.log 2>&1
ls -la $O/*.ncu-rep
