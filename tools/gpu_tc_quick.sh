#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "tensor_core or forced or golden_through_c_abi or pipelined" > gpurun_out/pytest_tc.log 2>&1; tail -3 gpurun_out/pytest_tc.log
timeout 300 python tools/lstm_bench.py tc 2>&1 | grep " tc "
timeout 300 python tools/lstm_bench.py warm 2>&1 | grep -E "inter (tile|tc)"
