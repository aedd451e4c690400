#!/bin/bash
# folded LayerNorm: targeted tests, then A/B on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "folded or streamk or act or glu or gate_residual" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_baseline_parity_gpu.py tests/test_fullsize_gpu.py -q -m "gpu" -p no:cacheprovider -x 2>&1 | tail -6
for fold in 1 0 1 0; do
  B200MIX_FOLD_LN=$fold timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-qwen > gpurun_out/s4_bench_fold$fold.log 2> gpurun_out/s4_bench_fold$fold.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s4_bench_fold$fold.log").read().strip().splitlines()[-1])
print("fold=$fold", d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:(v["ms"],v["achieved"]) for k,v in d["roofline"]["by_kernel"].items()}, d["clocks"]["sm_mhz"], d.get("parity"))
PY
done
timeout 600 python tools/shape_profile.py > gpurun_out/s4_shape_profile.log 2>&1; head -40 gpurun_out/s4_shape_profile.log
