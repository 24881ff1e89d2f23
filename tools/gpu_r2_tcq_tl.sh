#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/tcq_timeline.py > gpurun_out/r02_tcq_timeline.txt 2>&1; head -150 gpurun_out/r02_tcq_timeline.txt
