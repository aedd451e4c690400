#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/mufu_bench.py > gpurun_out/mufu_bench.log 2>&1; echo "== mufu exit $?"; cat gpurun_out/mufu_bench.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "== bench n2 exit $?"; tail -n 2 gpurun_out/bench_n2.log; tail -n 6 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.log 2> gpurun_out/bench_ref_n2.err; echo "== ref n2 exit $?"; tail -n 2 gpurun_out/bench_ref_n2.log; tail -n 3 gpurun_out/bench_ref_n2.err
