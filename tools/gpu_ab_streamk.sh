#!/bin/bash
# stream-K + fused upsample conv: targeted tests, then A/B of the schedule on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "up2x" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_vae_gpu.py tests/test_sd3_gpu.py tests/test_baseline_parity_gpu.py -q -m "gpu and not slow" -p no:cacheprovider -x 2>&1 | tail -4
for sk in 1 0 1 0; do
  B200MIX_STREAMK=$sk timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-qwen > gpurun_out/s3_bench_sk$sk.log 2> gpurun_out/s3_bench_sk$sk.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s3_bench_sk$sk.log").read().strip().splitlines()[-1])
print("streamk=$sk", d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:(v["ms"],v["achieved"]) for k,v in d["roofline"]["by_kernel"].items()}, d["clocks"]["sm_mhz"])
PY
done
B200MIX_STREAMK=1 timeout 600 python tools/shape_profile.py > gpurun_out/s3_shape_profile_sk1.log 2>&1; head -45 gpurun_out/s3_shape_profile_sk1.log
