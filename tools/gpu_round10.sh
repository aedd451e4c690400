#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_sd3_gpu.py -q -m gpu -x -k "sdpa or llama or sd3" -p no:cacheprovider > gpurun_out/test_a.log 2>&1; echo "== tests exit $?"; tail -n 8 gpurun_out/test_a.log
BNS=0 timeout 600 python tools/gemm_bench.py 2>&1 | tail -n 3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/bench6.log 2> gpurun_out/bench6.err; echo "== bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench6.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['by_kernel_ms'],d['roofline']['by_kernel_achieved'])"
