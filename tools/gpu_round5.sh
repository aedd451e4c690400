#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/test_all_gpu.log 2>&1; echo "== pytest -m gpu exit $?"; tail -n 8 gpurun_out/test_all_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 1500 python bench.py > gpurun_out/bench4.log 2> gpurun_out/bench4.err; echo "== bench exit $?"; tail -n 2 gpurun_out/bench4.log; tail -n 8 gpurun_out/bench4.err
