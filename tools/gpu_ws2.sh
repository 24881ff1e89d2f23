#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -5 gpurun_out/pytest.log
timeout 600 python tools/lstm_bench.py > gpurun_out/lstm_bench.txt 2>&1; grep -E " ws | ws2 " gpurun_out/lstm_bench.txt
for a in 8 5; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --pipe-intra-algo $a > gpurun_out/bench_ws2_a$a.json 2> gpurun_out/bench_ws2_a$a.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ws2_a$a.json").read().strip().splitlines()[-1])
    print("intra_algo=$a", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "in_order", round(d["in_order"]["value"]), "ms", round(d["ms_per_step"],1), d["streaming_vs_offline_maxabs"], "offline", round(d["offline"]["value"]), d["offline"]["stage_ms"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_ws2_a$a.err").read()[-2000:])
PY
done
