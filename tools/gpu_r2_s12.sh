#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_train.csv python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_train.log 2>&1; tail -2 gpurun_out/ncu_train.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_wgrad_tc -s 20 -c 1 -o gpurun_out/r02_prof_wgrad_tc python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_wgrad.log 2>&1; tail -2 gpurun_out/ncu_wgrad.log | cut -c1-200
