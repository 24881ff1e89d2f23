#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -12 | cut -c1-600
timeout 300 python tools/train_bench.py --train-tc 0 > gpurun_out/r02_train_bench_tc0.json 2>&1; tail -1 gpurun_out/r02_train_bench_tc0.json
timeout 300 python tools/train_bench.py --train-tc 1 > gpurun_out/r02_train_bench_tc1.json 2>&1; tail -1 gpurun_out/r02_train_bench_tc1.json
