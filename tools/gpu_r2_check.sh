#!/bin/bash
# short confirmation run: the GPU suite and a bench line (no CPU / library / training legs)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.txt 2>&1; tail -2 gpurun_out/r02_smoke.txt
timeout 300 python tools/conv_in_check.py > gpurun_out/r02_conv_in_tc.txt 2>&1; tail -3 gpurun_out/r02_conv_in_tc.txt | cut -c1-200
