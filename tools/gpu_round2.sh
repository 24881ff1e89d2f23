#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1
tail -15 gpurun_out/pytest.log
./tools/ubench > gpurun_out/ubench.txt 2>&1; cat gpurun_out/ubench.txt
timeout 900 python tools/lstm_bench.py > gpurun_out/lstm_bench.txt 2>&1; cat gpurun_out/lstm_bench.txt
