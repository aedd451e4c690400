#!/bin/bash
# LayerNorm-epilogue code split into its own instantiations: tests, then fold on / off on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "linear or conv" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_sd3_gpu.py -q -m "gpu and not slow" -p no:cacheprovider -x 2>&1 | tail -3
for fold in 1 0 1 0; do
  B200MIX_FOLD_LN=$fold timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-qwen > gpurun_out/s8_bench_fold$fold.log 2> gpurun_out/s8_bench_fold$fold.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s8_bench_fold$fold.log").read().strip().splitlines()[-1])
print("fold=$fold", d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:(v["ms"],v["achieved"]) for k,v in d["roofline"]["by_kernel"].items()}, d["clocks"]["sm_mhz"])
PY
done
