#!/bin/bash
# round-2 GPU session D: serial penalty chain restored, conv3 16-row tiles, tanh.approx, overlapped sampling read-back
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/d_tests.log 2>&1; tail -4 $O/d_tests.log
python tools/bench_conv.py "ds0" 10 fprop,dgrad > $O/d_conv.log 2>&1
python tools/bench_conv.py "up3" 10 fprop,dgrad >> $O/d_conv.log 2>&1
python tools/bench_conv.py "up4" 10 fprop,dgrad >> $O/d_conv.log 2>&1
python bench.py > $O/d_bench.log 2> $O/d_bench.err; tail -c 300 $O/d_bench.err
SG_PDL_S1=1 python bench.py --no-extras --no-cpu-baseline > $O/d_bench_pdl1.log 2>&1
