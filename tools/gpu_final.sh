#!/bin/bash
# One GPU session: parity tests, smoke, bench (+reference arm), ncu launch lists, full captures of the dominant kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1
tail -6 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_streaming.csv python tools/profile_run.py --mode streaming --chunks 12 --graph 0 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_offline.csv python tools/profile_run.py --mode offline > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ws -s 12 -c 2 -o gpurun_out/prof_ws python tools/profile_run.py --mode streaming --chunks 4 --graph 0 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ws -s 12 -c 2 -o gpurun_out/prof_ws2 python tools/profile_run.py --mode streaming --chunks 4 --graph 0 --intra-algo 8 > gpurun_out/ncu3b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 2 -c 1 -o gpurun_out/prof_tc python tools/profile_run.py --mode offline --batch 32 --frames 625 > gpurun_out/ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tile -s 2 -c 1 -o gpurun_out/prof_tile python tools/profile_run.py --mode offline --batch 8 --frames 200 --intra-algo 1 --inter-algo 1 > gpurun_out/ncu6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_in|backend_small|stft_features' -s 9 -c 3 -o gpurun_out/prof_small python tools/profile_run.py --mode streaming --chunks 4 --graph 0 > gpurun_out/ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prepare_batch -s 1 -c 1 -o gpurun_out/prof_prepare python tools/prepare_bench.py > gpurun_out/ncu7.log 2>&1
timeout 300 python tools/prepare_bench.py > gpurun_out/prepare_bench.txt 2>&1; tail -3 gpurun_out/prepare_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_core_tc -s 2 -c 1 -o gpurun_out/prof_attn_tc python tools/profile_attn_offline.py > gpurun_out/ncu8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_stream_pv -s 2 -c 1 -o gpurun_out/prof_attn_pv python tools/profile_attn.py > gpurun_out/ncu9.log 2>&1
timeout 300 python tools/variants_bench.py > gpurun_out/variants_bench.txt 2>&1; cat gpurun_out/variants_bench.txt
timeout 600 python tools/lstm_bench.py > gpurun_out/lstm_bench.txt 2>&1; tail -4 gpurun_out/lstm_bench.txt
ls -la gpurun_out | head -50
