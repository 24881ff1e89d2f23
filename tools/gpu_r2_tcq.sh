#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/tcp_check.py tcq > gpurun_out/r02_tcq_check.txt 2>&1; tail -22 gpurun_out/r02_tcq_check.txt | cut -c1-220
