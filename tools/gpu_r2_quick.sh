#!/bin/bash
# pipe / streaming tests and a short bench line (no CPU / library / training / strong legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "headline or pipe or stream or group or sliced" 2>&1 | tail -4
timeout 1200 python bench.py --steps 6 --warmup 3 --no-cpu --no-library --no-train --no-strong > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_quick.json').read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d["parity"]["rms"], d["parity"]["e2e_vs_device_maxabs"], "enqueue", d["config"]["host_enqueue_ms_per_step"])
PY
tail -3 gpurun_out/r02_bench_quick.err
