#!/bin/bash
# Round-end validation on one box: all -m gpu tests, smoke(), default bench line, reference arm (short), ncu traffic pass.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/test_all_gpu.log 2>&1; echo "== pytest -m gpu exit $?"; tail -n 4 gpurun_out/test_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['traffic'],d['cpu_baseline']['value'],d['qwen2vl_prefill']['value'],d['clocks'])"
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:igemm --csv --log-file gpurun_out/igemm_dram.csv python tools/profile_step.py > gpurun_out/ncu_t.log 2>&1; echo "traffic exit $?"
