#!/bin/bash
# round-2 session 2: grouped sweep on the TMA kernel with 8 / 32 hardware queues; full ncu capture of lstm_tcp_kernel
mkdir -p gpurun_out
timeout 600 python tools/group_sweep.py 8,8 8,16 16,8 16,16 25,8 25,16 32,8 > gpurun_out/r02_group_sweep_tcp.txt 2>&1; cat gpurun_out/r02_group_sweep_tcp.txt
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python tools/group_sweep.py 8,8 8,16 8,32 16,8 16,16 16,32 25,8 25,16 25,25 32,16 > gpurun_out/r02_group_sweep_tcp_conn32.txt 2>&1; cat gpurun_out/r02_group_sweep_tcp_conn32.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tcp -s 12 -c 2 -o gpurun_out/r02_prof_tcp python tools/profile_run.py --mode offline --batch 32 --frames 625 --intra-algo 9 --inter-algo 9 > gpurun_out/ncu_tcp.log 2>&1; tail -3 gpurun_out/ncu_tcp.log
