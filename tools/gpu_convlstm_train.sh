#!/bin/bash
# conv-LSTM training path on the GPU: all GPU tests, memcheck over the conv-LSTM gradient tests, training step of the RPI model
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -q -x -k "grad_rpi or grad_syn_convlstm" > gpurun_out/sanitizer_memcheck_convlstm_train.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck_convlstm_train.log
tail -3 gpurun_out/sanitizer_memcheck_convlstm_train.log
timeout 300 python tools/train_bench.py --config rpi --batch 8 --seconds 5 --steps 3 --cpu 0 > gpurun_out/train_bench_rpi.json 2> gpurun_out/train_bench_rpi.err; tail -c 700 gpurun_out/train_bench_rpi.json; tail -2 gpurun_out/train_bench_rpi.err
