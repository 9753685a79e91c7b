O=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "masked or bstats" > $O/t_bn.log 2>&1; tail -3 $O/t_bn.log
timeout 900 python -m pytest tests/test_stage1_gpu.py tests/test_stage2_gpu.py tests/test_parity_config_gpu.py tests/test_streams_gpu.py -q -x > $O/t_par.log 2>&1; tail -3 $O/t_par.log
for opt in dbg=0 dbg=0; do
SG_OPTS=$opt timeout 300 python bench.py --no-extras --no-cpu-baseline > $O/bench_$opt.json 2> $O/bench_$opt.err; python -c "import json;d=json.load(open('$O/bench_$opt.json'));print('$opt', d['ms_per_step'],d['value'],d['gpu_launches'],d['e2e']['value'])"
done
timeout 300 python tools/bench_stage2.py 64 5 bf16 > $O/s2.txt 2>&1; tail -1 $O/s2.txt | cut -c1-200
