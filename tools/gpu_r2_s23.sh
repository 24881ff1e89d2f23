#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu --no-library --no-train --no-strong > gpurun_out/r02_bench_s23.json 2> gpurun_out/r02_bench_s23.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_s23.json').read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"], d["roofline"]["stage_us_per_group"], "offline", d["offline"]["value"], "enqueue", d["config"]["host_enqueue_ms_per_step"])
PY
tail -3 gpurun_out/r02_bench_s23.err
