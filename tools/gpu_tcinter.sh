#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "tensor_core or forced or golden_through_c_abi" > gpurun_out/pytest_tc.log 2>&1; tail -3 gpurun_out/pytest_tc.log
timeout 300 python tools/lstm_bench.py warm > gpurun_out/lstm_bench_warm.txt 2>&1; grep -E "inter (tile|tc)" gpurun_out/lstm_bench_warm.txt
for a in 1 0; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --pipe-inter-algo $a > gpurun_out/bench_tci_a$a.json 2> gpurun_out/bench_tci_a$a.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_tci_a$a.json").read().strip().splitlines()[-1])
    print("inter_algo=$a", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "in_order", round(d["in_order"]["value"]), "ms", round(d["ms_per_step"],1), d["streaming_vs_offline_maxabs"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_tci_a$a.err").read()[-2000:])
PY
done
