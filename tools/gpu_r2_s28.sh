#!/bin/bash
mkdir -p gpurun_out
cd sound_bubble_b200/csrc/build
OBJS=$(ls *.o | grep -v tcp_soft | grep -v sb_lstm_tcp.o)
for n in 1 2; do
  nvcc -shared -o ../../libsoundbubble_sm100a.so $OBJS tcp_soft$n.o -gencode arch=compute_100a,code=sm_100a || exit 1
  cd ../../..
  echo "== SB_SOFT_EX2=$n"
  timeout 120 python tools/tcp_check.py pipe 2>&1 | tail -5 | cut -c1-150
  timeout 150 python tools/tcp_check.py parity 2>&1 | tail -7 | cut -c1-150
  cd sound_bubble_b200/csrc/build
done
