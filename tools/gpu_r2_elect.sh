#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/tcp_check.py tcq > gpurun_out/r02_tcq_elect.txt 2>&1; tail -16 gpurun_out/r02_tcq_elect.txt | cut -c1-200
timeout 240 python tools/tcp_check.py parity > gpurun_out/r02_tcp_parity_elect.txt 2>&1; tail -7 gpurun_out/r02_tcp_parity_elect.txt
