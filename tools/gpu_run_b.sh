#!/bin/bash
# round-2 GPU session B: thin kernels v2, capacity-aware wgrad clusters, fused CA
set -x
O=gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "thin or wgrad or ca_" > $O/b_kernels.log 2>&1; tail -3 $O/b_kernels.log
python tools/bench_conv.py "ds0" 10 fprop,dgrad > $O/b_conv.log 2>&1
python tools/bench_conv.py "up3" 10 fprop,dgrad >> $O/b_conv.log 2>&1
python tools/bench_conv.py "up4" 10 fprop,dgrad >> $O/b_conv.log 2>&1
python tools/bench_conv.py "" 10 wgrad,wgrad_cl > $O/b_wgrad.log 2>&1
python -m pytest tests -m gpu -x -q > $O/b_tests.log 2>&1; tail -3 $O/b_tests.log
python bench.py > $O/b_bench.log 2> $O/b_bench.err; tail -c 600 $O/b_bench.err
