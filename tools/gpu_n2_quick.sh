#!/bin/bash
# 2-GPU check of the final build: C-ABI collective test, weak-scaling line with the job (steps + all-gather + D2H)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_collective_gpu2.py -x -q -p no:cacheprovider 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err; echo "bench n2 exit $?"; tail -c 3000 gpurun_out/r2_bench_n2.log; tail -3 gpurun_out/r2_bench_n2.err
