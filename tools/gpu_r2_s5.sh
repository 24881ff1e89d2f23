#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tcp_check.py cell7 > gpurun_out/r02_tcp_cell7.txt 2>&1; cat gpurun_out/r02_tcp_cell7.txt | tail -8
timeout 300 python tools/tcp_check.py parity > gpurun_out/r02_tcp_parity_cell7.txt 2>&1; tail -8 gpurun_out/r02_tcp_parity_cell7.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "headline or golden or tensor" -s 2>&1 | grep -E "^(grouped|per_chunk|offline)|passed|failed" | cut -c1-330
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu --no-library --no-train --no-strong > gpurun_out/r02_bench_s5.json 2> gpurun_out/r02_bench_s5.err; tail -c 1500 gpurun_out/r02_bench_s5.json; tail -5 gpurun_out/r02_bench_s5.err
