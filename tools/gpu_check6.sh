#!/bin/bash
# epilogue operands prefetched one tile ahead + 200-register budget: tests, isolated probe and step A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "linear or conv" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_sd3_gpu.py -q -m "gpu and not slow" -p no:cacheprovider -x 2>&1 | tail -3
for v in main prev; do
  if [ "$v" = main ]; then unset B200MIX_LIB; else export B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so; fi
  echo "== $v"; timeout 300 python tools/ln_fold_probe.py 2>&1 | tail -5
done
for v in main prev main prev; do
  if [ "$v" = main ]; then unset B200MIX_LIB; else export B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-qwen > gpurun_out/s7_bench_$v.log 2> gpurun_out/s7_bench_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s7_bench_$v.log").read().strip().splitlines()[-1])
print("$v", d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:(v["ms"],v["achieved"]) for k,v in d["roofline"]["by_kernel"].items()}, d["clocks"]["sm_mhz"])
PY
done
unset B200MIX_LIB
