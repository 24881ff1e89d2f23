#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_core_tc -s 2 -c 1 -o gpurun_out/prof_attn_tc python tools/profile_attn_offline.py > gpurun_out/ncu8.log 2>&1; tail -1 gpurun_out/ncu8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_stream_pv -s 2 -c 1 -o gpurun_out/prof_attn_pv python tools/profile_attn.py > gpurun_out/ncu9.log 2>&1; tail -1 gpurun_out/ncu9.log
timeout 300 python tools/variants_bench.py > gpurun_out/variants_bench.txt 2>&1; cat gpurun_out/variants_bench.txt
