#!/bin/bash
# compute-sanitizer passes over the kernels at small shapes (the reference has no race / memory checking).
mkdir -p gpurun_out
K='test_stft_features or test_conv_in or test_film or test_intra_lstm or test_inter_lstm or test_backend or test_intra_convlstm or test_attention or tensor_core or pipelined or streaming_session or sliced or kernel_on_the_gpu'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_batching.py -q -x -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log
tail -4 gpurun_out/sanitizer_memcheck.log
K2='test_intra_lstm or test_inter_lstm or test_conv_in or test_backend or test_stft_features or test_attention'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$K2" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck.log
tail -4 gpurun_out/sanitizer_racecheck.log
grep -c "ERROR SUMMARY" gpurun_out/sanitizer_*.log
