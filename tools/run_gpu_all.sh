#!/bin/bash
# Full GPU validation: every -m gpu test, smoke(), the default bench line. Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/test_all_gpu.log 2>&1; echo "== pytest -m gpu exit $?"; tail -n 5 gpurun_out/test_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== bench exit $?"; tail -c 3000 gpurun_out/bench_default.log
