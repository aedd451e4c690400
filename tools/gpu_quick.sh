#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m "gpu and not slow" -p no:cacheprovider -rf -x 2>&1 | grep -v PASSED | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/r2_bench6.log 2> gpurun_out/r2_bench6.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench6.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:(v["ms"],v["achieved"]) for k,v in d["roofline"]["by_kernel"].items()}, d["clocks"])
print({k: (d[k].get("value"), d[k].get("ms_per_step"), d[k].get("error")) for k in ("sdxl_strong","sd3_b32","stdit2_b4") if k in d})
PY
tail -3 gpurun_out/r2_bench6.err
timeout 600 python tools/shape_profile.py > gpurun_out/r2_shape_profile6.log 2>&1; grep "^ln \|^gn \|attn \|total" gpurun_out/r2_shape_profile6.log | head -20
