#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -6
timeout 600 python tools/group_sweep.py 32,16 32,24 32,32 64,16 16,32 > gpurun_out/r02_group_sweep_tcr.txt 2>&1; tail -7 gpurun_out/r02_group_sweep_tcr.txt | cut -c1-200
