#!/bin/bash
# first call of a session: full -m gpu suite, decode-path rates, short bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rf > gpurun_out/s2_test_all.log 2>&1; echo "== pytest exit $?"; grep -v PASSED gpurun_out/s2_test_all.log | tail -12
timeout 300 python tools/skinny_probe.py 4 2>&1 | tail -5 | tee gpurun_out/s2_skinny.log
timeout 600 python tools/decode_probe.py > gpurun_out/s2_decode_probe.log 2>&1; echo "probe exit $?"; tail -22 gpurun_out/s2_decode_probe.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/s2_bench.log 2> gpurun_out/s2_bench.err; tail -c 3000 gpurun_out/s2_bench.log
timeout 600 python tools/shape_profile.py > gpurun_out/s2_shape_profile.log 2>&1; tail -60 gpurun_out/s2_shape_profile.log
