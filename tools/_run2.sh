O=gpurun_out
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -q > $O/t_dp.log 2>&1; tail -3 $O/t_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -c 300 $O/bench_2gpu.err; python -c "import json;d=json.load(open('$O/bench_2gpu.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d.get('dp_check'),d['stage2']['value'])"
