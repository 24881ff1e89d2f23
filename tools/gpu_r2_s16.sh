#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tcq_timeline.py tcr > gpurun_out/r02_tcr_timeline.txt 2>&1; head -120 gpurun_out/r02_tcr_timeline.txt
