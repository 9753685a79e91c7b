#!/bin/bash
# round-2 GPU session F: BatchNorm-backward statistics in the conv epilogue
set -x
O=gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "bstats or stats or conv_fprop_tcgen05 or conv_dgrad_tcgen05" > $O/f_kernels.log 2>&1; tail -4 $O/f_kernels.log
python -m pytest tests -m gpu -q > $O/f_tests.log 2>&1; tail -4 $O/f_tests.log
python bench.py > $O/f_bench.log 2> $O/f_bench.err; tail -c 300 $O/f_bench.err
python tools/bench_conv.py "s1.D1.ds3" 10 fprop,dgrad > $O/f_conv.log 2>&1
python tools/bench_conv.py "s2.G2.res1" 10 fprop,dgrad >> $O/f_conv.log 2>&1
python tools/bench_conv.py "s2.G2.up1" 10 fprop,dgrad >> $O/f_conv.log 2>&1
