#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 8" "14 6" "14 8" "14 10" "14 12" "14 16"; do
set -- $cfg; r=$1; d=$2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --ranges $r --depth $d > gpurun_out/bench_s2_r${r}_d$d.json 2> gpurun_out/bench_s2_r${r}_d$d.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_s2_r${r}_d$d.json").read().strip().splitlines()[-1])
    print("$r $d", round(d["config"]["host_enqueue_ms_per_step"],1), round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],1))
except Exception as e:
    print("ranges=$r depth=$d failed", e); print(open("gpurun_out/bench_s2_r${r}_d$d.err").read()[-1500:])
PY
done
