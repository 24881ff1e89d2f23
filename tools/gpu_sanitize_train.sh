#!/bin/bash
# compute-sanitizer passes over the training kernels at small shapes
mkdir -p gpurun_out
K='test_recurrent_path_gradients or test_module_gradients or test_variants or first_version'
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -q -x -k "$K" > gpurun_out/sanitizer_memcheck_train.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck_train.log
tail -4 gpurun_out/sanitizer_memcheck_train.log
K2='test_module_gradients or first_version'
timeout 170 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -q -x -k "$K2" > gpurun_out/sanitizer_racecheck_train.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck_train.log
tail -4 gpurun_out/sanitizer_racecheck_train.log
