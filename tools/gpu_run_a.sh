#!/bin/bash
# round-2 GPU session A: thin kernels + wgrad clusters (tests, per-layer bench, ncu captures, bench line)
set -x
O=gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "thin or wgrad" > $O/a_kernels.log 2>&1; tail -3 $O/a_kernels.log
python -m pytest tests -m gpu -x -q > $O/a_tests.log 2>&1; tail -3 $O/a_tests.log
python tools/bench_conv.py "" 10 fprop,dgrad > $O/a_conv.log 2>&1
python tools/bench_conv.py "" 10 wgrad,wgrad_cl > $O/a_wgrad.log 2>&1
SG_OPTS=wgrad_mc_odd=0 python tools/bench_conv.py "" 10 wgrad,wgrad_cl > $O/a_wgrad_pairs_only.log 2>&1
python bench.py > $O/a_bench.log 2> $O/a_bench.err; tail -c 600 $O/a_bench.err
# ncu captures (each command ran plainly above)
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tcp -s 2 -c 1 -o $O/ncu_conv_tcp -f python tools/bench_conv.py s1.D1.ds3 3 fprop > $O/ncu1.log 2>&1
$NCU -k regex:wgrad2 -s 2 -c 1 -o $O/ncu_wgrad_res1 -f python tools/bench_conv.py s2.G2.res1 3 wgrad_cl > $O/ncu2.log 2>&1
$NCU -k regex:wgrad2 -s 2 -c 1 -o $O/ncu_wgrad_res3 -f python tools/bench_conv.py s2.G2.res3 3 wgrad_cl > $O/ncu3.log 2>&1
$NCU -k regex:wgrad2 -s 2 -c 1 -o $O/ncu_wgrad_up0 -f python tools/bench_conv.py s2.G2.up0 3 wgrad > $O/ncu4.log 2>&1
$NCU -k regex:"conv3_k4s2|convt3_k4s2" -s 4 -c 2 -o $O/ncu_thin_g2up3 -f python tools/bench_conv.py s2.G2.up3 3 fprop,dgrad > $O/ncu5.log 2>&1
$NCU -k regex:"conv3_k4s2|convt3_k4s2" -s 4 -c 2 -o $O/ncu_thin_d2ds0 -f python tools/bench_conv.py s2.D2.ds0 3 fprop,dgrad > $O/ncu6.log 2>&1
$NCU -k regex:"bn_" -s 6 -c 3 -o $O/ncu_bn -f python tools/probe_bn.py > $O/ncu7.log 2>&1
ls -la $O/*.ncu-rep
