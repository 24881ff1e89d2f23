#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_s3.json 2> gpurun_out/r02_bench_s3.err; tail -c 7000 gpurun_out/r02_bench_s3.json; tail -5 gpurun_out/r02_bench_s3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 1200 gpurun_out/r02_bench_ref.json; tail -3 gpurun_out/r02_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-library --no-strong --no-parity > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
