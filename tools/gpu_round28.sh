timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_stdit2_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/bench28.log 2> gpurun_out/bench28.err; echo "== bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench28.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['by_kernel_ms'],d['clocks'])"
