#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m "gpu and not slow" -p no:cacheprovider -rf 2>&1 | grep -v PASSED | tail -12
timeout 1800 python -m pytest tests/ -q -m "gpu and slow" -p no:cacheprovider -rf 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/r2_bench5.log 2> gpurun_out/r2_bench5.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench5.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:(v["ms"],v["achieved"]) for k,v in d["roofline"]["by_kernel"].items()}, d["clocks"])
print({k: (d[k].get("value"), d[k].get("ms_per_step")) for k in ("sdxl_strong","sd3_b32","stdit2_b4") if k in d}, d["job"])
PY
tail -3 gpurun_out/r2_bench5.err
