#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tcp_check.py pipe > gpurun_out/r02_tcp_pipe.txt 2>&1; tail -8 gpurun_out/r02_tcp_pipe.txt
timeout 150 python tools/tcp_check.py parity > gpurun_out/r02_tcp_parity_pipe.txt 2>&1; tail -8 gpurun_out/r02_tcp_parity_pipe.txt
