#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_tcr -s 6 -c 1 -o gpurun_out/r02_prof_tcr python tools/tcp_check.py pipe > gpurun_out/ncu_tcr.log 2>&1; tail -3 gpurun_out/ncu_tcr.log | cut -c1-200
