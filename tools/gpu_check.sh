#!/bin/bash
# One GPU-box session: the whole -m gpu suite, smoke(), the bench line.   gpurun --timeout 1200 -- 'bash tools/gpu_check.sh'
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/check_tests.log 2>&1; tail -4 $O/check_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/check_smoke.log 2>&1; tail -1 $O/check_smoke.log
python bench.py > $O/check_bench.log 2> $O/check_bench.err; tail -c 300 $O/check_bench.err
