#!/bin/bash
# Co-residency experiment: BatchNorm kernels capped at 88 / 80 registers (tools build them as
# imagegenerator_b200/libsgb200_bn{88,80}.so) + the wgrad pipeline at <= 150 KB of shared memory, so that one wgrad CTA fits next to
# two BatchNorm CTAs on an SM.  Prints ms per step of both stages for every combination.
O=gpurun_out
for lib in "" _bn88 _bn80; do
  for kb in 200 150 110; do
    export SG_LIB=$PWD/imagegenerator_b200/libsgb200$lib.so SG_OPTS=wgrad_smem_kb=$kb
    a=$(timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>$O/exp_cores.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])")
    b=$(timeout 300 python tools/bench_stage2.py 64 5 2>>$O/exp_cores.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])")
    echo "lib=libsgb200$lib.so wgrad_smem_kb=$kb stage1_ms=$a stage2_ms=$b"
  done
done
