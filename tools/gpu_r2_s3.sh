#!/bin/bash
# round-2 session 3: the new bench line (grouped default, parity, roofline, baselines), headline parity tests, attention backward
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "headline" -s > gpurun_out/r02_pytest_headline.txt 2>&1; tail -8 gpurun_out/r02_pytest_headline.txt
timeout 600 python -m pytest tests/test_zz_experimental_gpu.py -m gpu -q --runxfail > gpurun_out/r02_attn_bwd_runxfail.txt 2>&1; tail -30 gpurun_out/r02_attn_bwd_runxfail.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_s3.json 2> gpurun_out/r02_bench_s3.err; tail -c 6000 gpurun_out/r02_bench_s3.json; tail -5 gpurun_out/r02_bench_s3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 900 gpurun_out/r02_bench_ref.json; tail -3 gpurun_out/r02_bench_ref.err
