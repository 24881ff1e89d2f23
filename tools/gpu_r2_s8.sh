#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tcp_check.py cell7 > gpurun_out/r02_tcp_cell7_v3.txt 2>&1; cat gpurun_out/r02_tcp_cell7_v3.txt | tail -6
timeout 300 python tools/tcp_check.py parity > gpurun_out/r02_tcp_parity_v3.txt 2>&1; tail -8 gpurun_out/r02_tcp_parity_v3.txt
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu --no-library --no-train --no-strong > gpurun_out/r02_bench_s8.json 2> gpurun_out/r02_bench_s8.err; tail -c 1300 gpurun_out/r02_bench_s8.json | head -c 900; tail -5 gpurun_out/r02_bench_s8.err
