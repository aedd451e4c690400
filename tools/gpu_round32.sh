timeout 900 python tools/bench_models.py > gpurun_out/bench_models32.log 2>&1; cat gpurun_out/bench_models32.log | tail -n 6
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench32.log 2> gpurun_out/bench32.err; python -c "
import json;d=json.loads(open('gpurun_out/bench32.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['qwen2vl_prefill']['value'],d['qwen2vl_prefill']['ms_per_prefill'],d['qwen2vl_prefill']['e2e'])"
