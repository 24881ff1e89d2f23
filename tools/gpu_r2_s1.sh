#!/bin/bash
# round-2 session 1: state check of the restored tree (tests, tcp kernel, grouped sweep, bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02_pytest_gpu.txt
timeout 300 python tools/tcp_check.py parity > gpurun_out/r02_tcp_parity.txt 2>&1; tail -12 gpurun_out/r02_tcp_parity.txt
timeout 300 python tools/tcp_check.py time > gpurun_out/r02_tcp_time.txt 2>&1; tail -12 gpurun_out/r02_tcp_time.txt
timeout 900 python tools/group_sweep.py 1,8 4,8 8,8 8,16 16,8 > gpurun_out/r02_group_sweep.txt 2>&1; cat gpurun_out/r02_group_sweep.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_s1.json 2> gpurun_out/r02_bench_s1.err; tail -c 3000 gpurun_out/r02_bench_s1.json; tail -5 gpurun_out/r02_bench_s1.err
