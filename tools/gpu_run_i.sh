#!/bin/bash
# round-2 GPU session I: final build -- tests, bench, ncu capture of the roofline kernel
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/i_tests.log 2>&1; tail -4 $O/i_tests.log
python bench.py > $O/i_bench.log 2> $O/i_bench.err; tail -c 300 $O/i_bench.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tcp -s 2 -c 1 -o $O/ncu_conv_tcp_final -f python tools/bench_conv.py s1.D1.ds3 3 fprop > $O/ncu21.log 2>&1
python tools/bench_conv.py "" 10 fprop,dgrad > $O/i_conv.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/i_ref.log 2> $O/i_ref.err; tail -c 300 $O/i_ref.log
