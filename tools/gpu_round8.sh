#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/test_all_gpu2.log 2>&1; echo "== pytest -m gpu exit $?"; tail -n 6 gpurun_out/test_all_gpu2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/bench5.log 2> gpurun_out/bench5.err; echo "== bench (PDL) exit $?"; tail -n 2 gpurun_out/bench5.log | cut -c1-400; tail -n 5 gpurun_out/bench5.err
B200MIX_NO_PDL=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/bench5_nopdl.log 2> gpurun_out/bench5_nopdl.err; echo "== bench (no PDL) exit $?"; tail -n 2 gpurun_out/bench5_nopdl.log | cut -c1-400
