timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "linear or conv or glu or act" 2>&1 | tail -n 4
timeout 400 python tools/epi_probe.py 2>&1 | tail -n 16
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/bench25.log 2> gpurun_out/bench25.err; echo "== bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench25.log').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['by_kernel_ms'],d['clocks'])"
