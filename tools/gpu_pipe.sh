#!/bin/bash
mkdir -p gpurun_out
for cfg in "6 4" "6 5" "6 6" "6 8" "4 4" "3 6"; do
set -- $cfg; r=$1; d=$2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --ranges $r --depth $d > gpurun_out/bench_pipe_r${r}_d$d.json 2> gpurun_out/bench_pipe_r${r}_d$d.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pipe_r${r}_d$d.json").read().strip().splitlines()[-1])
    print("ranges=$r depth=$d", "enq_ms", round(d["config"]["host_enqueue_ms_per_step"],1), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "in_order", round(d["in_order"]["value"]), "ms", round(d["ms_per_step"],1), d["streaming_vs_offline_maxabs"])
except Exception as e:
    print("ranges=$r depth=$d failed", e); print(open("gpurun_out/bench_pipe_r${r}_d$d.err").read()[-2000:])
PY
done
