#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_clip_llava_gpu.py -x -q -m "gpu and not slow" -p no:cacheprovider 2>&1 | tail -15
timeout 1500 python -m pytest tests/ -q -m "gpu and not slow" -p no:cacheprovider --deselect tests/test_vae_gpu.py --deselect tests/test_clip_llava_gpu.py 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen --no-extras > gpurun_out/r2_bench4.log 2> gpurun_out/r2_bench4.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench4.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:(v["ms"],v["achieved"]) for k,v in d["roofline"]["by_kernel"].items()}, d["clocks"])
PY
tail -3 gpurun_out/r2_bench4.err
