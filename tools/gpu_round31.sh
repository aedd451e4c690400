timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "small_cin" 2>&1 | tail -n 3
timeout 300 python tools/misc_probe.py 2>&1 | head -n 2
timeout 600 python -m pytest tests/test_unet_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/bench31.log 2> gpurun_out/bench31.err; echo "== bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench31.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['by_kernel_ms'],d['clocks'])"
