#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tcq -s 14 -c 1 -o gpurun_out/r02_prof_tcq python tools/tcp_check.py tcq > gpurun_out/ncu_tcq.log 2>&1; tail -3 gpurun_out/ncu_tcq.log | cut -c1-200
