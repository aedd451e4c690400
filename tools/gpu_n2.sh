#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_collective_gpu2.py -x -q -p no:cacheprovider 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err; echo "bench n2 exit $?"; tail -c 2500 gpurun_out/r2_bench_n2.log; tail -5 gpurun_out/r2_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n2.log 2> gpurun_out/r2_bench_ref_n2.err; echo "ref n2 exit $?"; tail -c 600 gpurun_out/r2_bench_ref_n2.log
