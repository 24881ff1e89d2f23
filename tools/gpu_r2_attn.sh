#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_tc_check.py > gpurun_out/r02_attn_tc_check.txt 2>&1; tail -12 gpurun_out/r02_attn_tc_check.txt | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x -k "attn" 2>&1 | tail -4
timeout 300 python tools/variants_bench.py > gpurun_out/r02_variants_bench.txt 2>&1; tail -5 gpurun_out/r02_variants_bench.txt | cut -c1-200
