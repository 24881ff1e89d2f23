#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/tc_check.py parity > gpurun_out/tc_parity.txt 2>&1; echo "rc=$?" >> gpurun_out/tc_parity.txt; cat gpurun_out/tc_parity.txt | tail -20
timeout 180 python tools/tc_check.py golden > gpurun_out/tc_golden.txt 2>&1; echo "rc=$?" >> gpurun_out/tc_golden.txt; cat gpurun_out/tc_golden.txt | tail -12
timeout 300 python tools/lstm_bench.py tc > gpurun_out/lstm_bench_tc.txt 2>&1; grep -E "tc |tile " gpurun_out/lstm_bench_tc.txt | tail -20
nvidia-smi --query-gpu=name,clocks.sm --format=csv | tail -1
