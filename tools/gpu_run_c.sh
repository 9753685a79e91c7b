#!/bin/bash
# round-2 GPU session C: penalty chain on its own stream, cross-boundary checkpoint test, parity policy; ncu of thin v2 + dense kernels
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/c_tests.log 2>&1; tail -6 $O/c_tests.log
python bench.py > $O/c_bench.log 2> $O/c_bench.err; tail -c 400 $O/c_bench.err
NCU="ncu --set full --clock-control none --import-source on"
python tools/profile_step.py > $O/c_pstep.log 2>&1; tail -2 $O/c_pstep.log
$NCU -k regex:"ca_forward|ca_backward|critic_loss|gen_loss|AdamF|adam_tick" -c 12 -o $O/ncu_dense -f python tools/profile_step.py > $O/ncu8.log 2>&1
$NCU -k regex:"conv3_k4s2|convt3_k4s2" -s 4 -c 2 -o $O/ncu_thin2_g2up3 -f python tools/bench_conv.py s2.G2.up3 3 fprop,dgrad > $O/ncu9.log 2>&1
$NCU -k regex:"conv3_k4s2|convt3_k4s2" -s 4 -c 2 -o $O/ncu_thin2_d2ds0 -f python tools/bench_conv.py s2.D2.ds0 3 fprop,dgrad > $O/ncu10.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_s1.csv python tools/profile_step.py 128 2 > $O/ncu11.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_s2.csv python tools/bench_stage2.py 64 1 bf16 eager > $O/ncu12.log 2>&1
ls -la $O/*.ncu-rep $O/*.csv
