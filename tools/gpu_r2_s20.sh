#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tcp_check.py cw16 > gpurun_out/r02_tcr_cw16.txt 2>&1; tail -8 gpurun_out/r02_tcr_cw16.txt
