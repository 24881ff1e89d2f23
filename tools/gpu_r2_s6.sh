#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tcp_check.py cell7 > gpurun_out/r02_tcp_cell7_lds.txt 2>&1; cat gpurun_out/r02_tcp_cell7_lds.txt | tail -6
timeout 300 python tools/tcp_check.py parity > gpurun_out/r02_tcp_parity_lds.txt 2>&1; tail -8 gpurun_out/r02_tcp_parity_lds.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu --no-library --no-train --no-strong > gpurun_out/r02_bench_s6.json 2> gpurun_out/r02_bench_s6.err; tail -c 1300 gpurun_out/r02_bench_s6.json; tail -5 gpurun_out/r02_bench_s6.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tcp -s 12 -c 2 -o gpurun_out/r02_prof_tcp_v2 python tools/profile_run.py --mode offline --batch 32 --frames 625 --intra-algo 9 --inter-algo 9 > gpurun_out/ncu_tcp2.log 2>&1; tail -2 gpurun_out/ncu_tcp2.log
