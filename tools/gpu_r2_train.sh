#!/bin/bash
# training path: GPU gradient tests, the step time, the launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/train_bench.py > gpurun_out/r02_train_bench.json 2> gpurun_out/train_bench.err; tail -c 420 gpurun_out/r02_train_bench.json
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_train.csv python tools/train_bench.py --steps 2 --warmup 1 > gpurun_out/ncu_train.log 2>&1
