#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch lists + one full capture of the dominant kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1
tail -30 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --pdl 1 --no-cpu > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; tail -c 1500 gpurun_out/bench_pdl.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_streaming.csv python tools/profile_run.py --mode streaming --chunks 12 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_offline.csv python tools/profile_run.py --mode offline > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_lane -s 12 -c 2 -o gpurun_out/prof_lane python tools/profile_run.py --mode streaming --chunks 4 --graph 0 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tile -s 2 -c 2 -o gpurun_out/prof_tile python tools/profile_run.py --mode offline > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out
