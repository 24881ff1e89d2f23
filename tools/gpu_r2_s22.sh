#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_in_tc -s 14 -c 1 -o gpurun_out/r02_prof_conv_in_tc python tools/conv_in_check.py > gpurun_out/ncu_cit.log 2>&1; tail -3 gpurun_out/ncu_cit.log | cut -c1-200
