#!/bin/bash
# GPU session for the training path: all GPU tests, training-step bench, ncu launch list + full captures of its top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -8 gpurun_out/pytest.log
timeout 600 python tools/train_bench.py --batch 8 --seconds 5 --steps 3 > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; tail -c 1500 gpurun_out/train_bench.json; tail -3 gpurun_out/train_bench.err
timeout 300 python tools/train_bench.py --batch 2 --seconds 1 --steps 3 --cpu 0 > gpurun_out/train_bench_small.json 2>> gpurun_out/train_bench.err; tail -c 600 gpurun_out/train_bench_small.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --batch 8 --seconds 5 --steps 1 --warmup 0 --cpu 0 > gpurun_out/ncu_t1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lstm_train_fwd|lstm_train_bwd|outer_kernel|rowgemm' -s 4 -c 8 -o gpurun_out/prof_train python tools/train_bench.py --batch 4 --seconds 2 --steps 1 --warmup 0 --cpu 0 > gpurun_out/ncu_t2.log 2>&1
ls -la gpurun_out | grep -E "train|pytest"
