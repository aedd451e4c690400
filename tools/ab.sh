#!/bin/bash
# usage: tools/ab.sh NAME1 NAME2 ...   (variants built by tools/build_variant.sh; "main" = the in-tree library)
# Runs the SDXL bench for each variant, twice, interleaved, on the same box.
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" = main ]; then unset B200MIX_LIB; else export B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so; fi
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen --no-extras > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err || { echo "$v FAILED"; tail -n 5 gpurun_out/ab_$v.err; }
    python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads(open(f"gpurun_out/ab_{v}.log").read().strip().splitlines()[-1])
print(f"{v:12s} {d['ms_per_step']:.3f} ms  { {k: x['ms'] for k, x in d['roofline']['by_kernel'].items()} }  sm_mhz={d['clocks'].get('sm_mhz')}", flush=True)
PY
  done
done
