#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tcp_check.py cell7 > gpurun_out/r02_tcp_n256.txt 2>&1; cat gpurun_out/r02_tcp_n256.txt | tail -6
timeout 300 python tools/tcp_check.py parity > gpurun_out/r02_tcp_parity_n256.txt 2>&1; tail -8 gpurun_out/r02_tcp_parity_n256.txt
