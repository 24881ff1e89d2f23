#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_wsp -s 6 -c 1 -o gpurun_out/prof_wsp python tools/profile_run.py --mode streaming --chunks 3 --graph 0 --intra-algo 8 > gpurun_out/ncu_wsp.log 2>&1
tail -3 gpurun_out/ncu_wsp.log
