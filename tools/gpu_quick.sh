#!/bin/bash
# quick check of a kernel change: GPU tests, LSTM micro-benchmark (ws lines), pipelined bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 600 python tools/lstm_bench.py > gpurun_out/lstm_bench.txt 2>&1; grep -E " ws " gpurun_out/lstm_bench.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "in_order", round(d["in_order"]["value"]), "ms", round(d["ms_per_step"],1), d["streaming_vs_offline_maxabs"], d["roofline"]["avg_launch_us"], d["roofline"]["stage_us_per_chunk"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_quick.err").read()[-2000:])
PY
