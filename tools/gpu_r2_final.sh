#!/bin/bash
# round-2 closing session on one B200: tests, smoke, bench (+ reference arm), ncu launch list of the bench command, full capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
nproc > gpurun_out/r02_host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r02_host.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.txt 2>&1; tail -3 gpurun_out/r02_smoke.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 1500 gpurun_out/r02_bench.json; tail -5 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 600 gpurun_out/r02_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-library --no-strong --no-parity > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tcr -s 12 -c 2 -o gpurun_out/r02_prof_tcr python tools/profile_run.py --mode offline --batch 32 --frames 64 --intra-algo 7 --inter-algo 7 > gpurun_out/ncu_tcr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_in|deconv_spec|istft_ola|stft_features' -s 4 -c 4 -o gpurun_out/r02_prof_frontback_v2 python tools/profile_run.py --mode offline --batch 32 --frames 32 > gpurun_out/ncu_fb.log 2>&1
timeout 300 python tools/variants_bench.py > gpurun_out/r02_variants_bench.txt 2>&1; cat gpurun_out/r02_variants_bench.txt | tail -8
ls -la gpurun_out | tail -20
