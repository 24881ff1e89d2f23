#!/bin/bash
# One GPU session for the training path (the pieces this round ran as separate short calls):
#   GPU tests, training-step bench (TFG_S + Raspberry-Pi model, A/Bs of the kernel options), launch list, full ncu capture
#   of one block's forward + backward kernels, compute-sanitizer over the gradient tests.  2 GPUs: add tools/train_ddp.py.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python tools/train_bench.py --batch 8 --seconds 5 --steps 3 > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; tail -c 900 gpurun_out/train_bench.json
timeout 300 python tools/train_bench.py --batch 8 --seconds 5 --steps 3 --cpu 0 --one-row 1 > gpurun_out/train_bench_one_row.json 2>> gpurun_out/train_bench.err
timeout 300 python tools/train_bench.py --batch 8 --seconds 5 --steps 3 --cpu 0 --ffma2 0 > gpurun_out/train_bench_ffma2_0.json 2>> gpurun_out/train_bench.err
timeout 300 python tools/train_bench.py --config rpi --batch 8 --seconds 5 --steps 3 --cpu 0 > gpurun_out/train_bench_rpi.json 2>> gpurun_out/train_bench.err; tail -c 400 gpurun_out/train_bench_rpi.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --batch 8 --seconds 5 --steps 1 --warmup 0 --cpu 0 > gpurun_out/ncu_t1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lstm_train_bwd|lstm_train_fwd|outer_kernel|ln_bwd|rowgemm' -s 20 -c 22 -o gpurun_out/prof_train python tools/train_bench.py --batch 4 --seconds 2 --steps 1 --warmup 0 --cpu 0 > gpurun_out/ncu_t2.log 2>&1
K='test_recurrent_path_gradients or test_module_gradients or test_variants or first_version'
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -q -x -k "$K" > gpurun_out/sanitizer_memcheck_train.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck_train.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -q -x -k "test_module_gradients or first_version" > gpurun_out/sanitizer_racecheck_train.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck_train.log
ls -la gpurun_out | grep -E "train|pytest|smoke"
