#!/bin/bash
# round-2 exploration: grouped-chunk sweep (per-chunk feed, stateless intra ranges)
mkdir -p gpurun_out
timeout 900 python tools/group_sweep.py 1,8 4,8 8,8 4,16 8,16 4,32 16,8 > gpurun_out/r02_group_sweep2.txt 2>&1; cat gpurun_out/r02_group_sweep2.txt
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 900 python tools/group_sweep.py 1,8 8,8 4,16 8,16 4,32 > gpurun_out/r02_group_sweep2_conn32.txt 2>&1; cat gpurun_out/r02_group_sweep2_conn32.txt
