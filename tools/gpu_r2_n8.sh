#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-library > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -c 1500 gpurun_out/r02_bench_n8.json; tail -3 gpurun_out/r02_bench_n8.err
