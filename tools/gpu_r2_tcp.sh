#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tcp_check.py parity > gpurun_out/r02_tcp_parity.txt 2>&1; tail -20 gpurun_out/r02_tcp_parity.txt
timeout 300 python tools/tcp_check.py time > gpurun_out/r02_tcp_time.txt 2>&1; tail -20 gpurun_out/r02_tcp_time.txt
timeout 300 python tools/tcp_check.py golden > gpurun_out/r02_tcp_golden.txt 2>&1; tail -8 gpurun_out/r02_tcp_golden.txt
