#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
timeout 1200 python bench.py --steps 6 --warmup 3 --no-cpu --no-library --no-train --no-strong > gpurun_out/r02_bench_s27.json 2> gpurun_out/r02_bench_s27.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_s27.json').read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d["parity"]["rms"], "offline", d["offline"]["value"], d["roofline"]["stage_us_per_group"])
PY
tail -3 gpurun_out/r02_bench_s27.err
