#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1
tail -8 gpurun_out/pytest.log
timeout 900 python tools/lstm_bench.py > gpurun_out/lstm_bench.txt 2>&1; cat gpurun_out/lstm_bench.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --pdl 1 --no-cpu > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; tail -c 2500 gpurun_out/bench_pdl.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_streaming.csv python tools/profile_run.py --mode streaming --chunks 12 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ws -s 12 -c 2 -o gpurun_out/prof_ws python tools/profile_run.py --mode streaming --chunks 4 --graph 0 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tile -s 2 -c 1 -o gpurun_out/prof_tile python tools/profile_run.py --mode offline --batch 8 --frames 200 --intra-algo 1 --inter-algo 1 > gpurun_out/ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_in|backend_small|stft_features' -s 9 -c 3 -o gpurun_out/prof_small python tools/profile_run.py --mode streaming --chunks 4 --graph 0 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out
