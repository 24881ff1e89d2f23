#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
for cfg in "5 8" "8 8" "8 12" "8 16"; do
set -- $cfg; a=$1; d=$2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --pipe-intra-algo $a --depth $d > gpurun_out/bench_w_$a_$d.json 2> gpurun_out/bench_w.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_w_$a_$d.json").read().strip().splitlines()[-1])
    print("intra_algo=$a depth=$d", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],1), d["streaming_vs_offline_maxabs"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_w.err").read()[-1500:])
PY
done
