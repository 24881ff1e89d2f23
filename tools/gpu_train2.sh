#!/bin/bash
# 2-GPU session: data-parallel training step over NCCL + full ncu capture of one block's backward kernels (GPU 0 only)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_ddp.py --batch-per-gpu 4 --seconds 5 --steps 3 > gpurun_out/train_ddp_n2.json 2> gpurun_out/train_ddp_n2.err; tail -c 700 gpurun_out/train_ddp_n2.json; tail -3 gpurun_out/train_ddp_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/train_ddp.py --batch-per-gpu 4 --seconds 5 --steps 3 > gpurun_out/train_ddp_n1.json 2>> gpurun_out/train_ddp_n2.err; tail -c 500 gpurun_out/train_ddp_n1.json
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lstm_train_bwd|outer_kernel|ln_bwd|rowgemm' -s 12 -c 18 -o gpurun_out/prof_train_bwd python tools/train_bench.py --batch 4 --seconds 2 --steps 1 --warmup 0 --cpu 0 > gpurun_out/ncu_t3.log 2>&1
ls -la gpurun_out | grep -E "train"
