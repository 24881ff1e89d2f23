#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/group_sweep.py 64,16 64,24 64,10 128,8 128,16 48,16 > gpurun_out/r02_group_sweep_tcr2.txt 2>&1; tail -7 gpurun_out/r02_group_sweep_tcr2.txt | cut -c1-200
