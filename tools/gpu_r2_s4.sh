#!/bin/bash
# round-2 session 4: conv_in with four bins per thread, attention gradients on the B200, full GPU suite, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_s4.txt 2>&1; tail -12 gpurun_out/r02_pytest_gpu_s4.txt
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu --no-library --no-train > gpurun_out/r02_bench_s4.json 2> gpurun_out/r02_bench_s4.err; tail -c 2500 gpurun_out/r02_bench_s4.json; tail -5 gpurun_out/r02_bench_s4.err
