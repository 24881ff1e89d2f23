#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 600 python tools/lstm_bench.py > gpurun_out/lstm_bench.txt 2>&1; grep -E "T=  1" gpurun_out/lstm_bench.txt | grep -E "ws|tile"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --pdl 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --pdl 1 --no-cpu > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; tail -c 600 gpurun_out/bench_pdl.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ws -s 12 -c 1 -o gpurun_out/prof_ws python tools/profile_run.py --mode streaming --chunks 4 --graph 0 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_in|backend_small|stft_features|lstm_tile' -s 9 -c 4 -o gpurun_out/prof_small python tools/profile_run.py --mode streaming --chunks 4 --graph 0 > gpurun_out/ncu5.log 2>&1
