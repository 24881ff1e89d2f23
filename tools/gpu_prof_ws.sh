#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ws -s 12 -c 1 -o gpurun_out/prof_ws_m2 python tools/profile_run.py --mode streaming --chunks 4 --graph 0 > gpurun_out/ncu_m2.log 2>&1
tail -2 gpurun_out/ncu_m2.log
