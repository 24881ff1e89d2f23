#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -c 2500 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; tail -c 300 gpurun_out/r02_bench_ref_n2.json
