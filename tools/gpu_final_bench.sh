#!/bin/bash
# round-end validation, part 2: GroupNorm / UNet checks of the last kernel edit, ncu evidence of the final build (launch list,
# --set full per family, igemm DRAM traffic), then the driver's command (default bench.py) reading the fresh traffic figure
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py -q -m "gpu and not slow" -p no:cacheprovider -x -k "groupnorm or unet or vae" 2>&1 | tail -3 | tee gpurun_out/final_gn_check.log
grep -q "failed\|error" gpurun_out/final_gn_check.log && { echo "GN CHECK FAILED - stopping"; exit 3; }
N="ncu --profile-from-start off --clock-control none"
$N --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:igemm --csv --log-file gpurun_out/r02_igemm_dram.csv python tools/profile_step.py > gpurun_out/ncu_t.log 2>&1; echo "traffic exit $?"
python tools/ncu_summary.py traffic gpurun_out/r02_igemm_dram.csv > gpurun_out/r02_igemm_dram_traffic.json && cp gpurun_out/r02_igemm_dram_traffic.json profiles/r02_igemm_dram_traffic.json; cat profiles/r02_igemm_dram_traffic.json
$N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py > gpurun_out/ncu_l.log 2>&1; echo "launch list exit $?"
$N --set full --import-source on -k regex:igemm_kernel -s 200 -c 6 -f -o gpurun_out/r02_igemm python tools/profile_step.py > gpurun_out/ncu_f1.log 2>&1; echo "igemm exit $?"
$N --set full --import-source on -k "regex:gn_" -s 20 -c 4 -f -o gpurun_out/r02_norm python tools/profile_step.py > gpurun_out/ncu_f3.log 2>&1; echo "norm exit $?"
timeout 1500 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['traffic'],d['roofline'].get('unfused_layernorm'),d['cpu_baseline']['value'],d['qwen2vl_prefill']['value'],d['clocks'],d.get('parity',{}).get('cosine'))
print({k: (d[k].get('value'), d[k].get('ms_per_step'), d[k].get('error')) for k in ('sdxl_strong','sd3_b32','stdit2_b4') if k in d})"
tail -3 gpurun_out/bench_default.err
timeout 600 python tools/shape_profile.py > gpurun_out/r02_shape_profile.log 2>&1; head -8 gpurun_out/r02_shape_profile.log
