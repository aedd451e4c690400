#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "folded" 2>&1 | tail -3
for v in main lnp1 lnp2; do  # variants: tools/build_variant.sh lnp1 -DLN_PROBE=1, lnp2 -DLN_PROBE=2 (skipped when absent)
  if [ "$v" = main ]; then unset B200MIX_LIB; else export B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so; [ -f "$B200MIX_LIB" ] || continue; fi
  echo "== $v"; timeout 300 python tools/ln_fold_probe.py 2>&1 | tail -6
done
unset B200MIX_LIB
for fold in 1 0; do
  B200MIX_FOLD_LN=$fold timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-qwen > gpurun_out/s5_bench_fold$fold.log 2> gpurun_out/s5_bench_fold$fold.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s5_bench_fold$fold.log").read().strip().splitlines()[-1])
print("fold=$fold", d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:(v["ms"],v["achieved"]) for k,v in d["roofline"]["by_kernel"].items()}, d["clocks"]["sm_mhz"])
PY
done
timeout 600 python tools/shape_profile.py > gpurun_out/s5_shape_profile.log 2>&1; grep "^lin" gpurun_out/s5_shape_profile.log | head -12
