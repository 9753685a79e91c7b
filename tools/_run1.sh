O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/check_tests.log 2>&1; tail -4 $O/check_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/check_smoke.log 2>&1; tail -1 $O/check_smoke.log
timeout 900 python bench.py > $O/check_bench.log 2> $O/check_bench.err; tail -c 300 $O/check_bench.err; python -c "import json;d=json.load(open('$O/check_bench.log'));print(d['ms_per_step'],d['value'],d['gpu_launches'],d['e2e']['value'],d['roofline']['frac'],d['stage2']['ms_per_step'],d['stage2']['e2e']['value'],d['sampling']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s1.csv python tools/profile_step.py 128 2 > $O/prof_s1.log 2>&1; tail -2 $O/prof_s1.log
