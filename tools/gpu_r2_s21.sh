#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/conv_in_check.py dbg > gpurun_out/r02_conv_in_tc.txt 2>&1; tail -8 gpurun_out/r02_conv_in_tc.txt
