#!/bin/sh
# TEST INFRASTRUCTURE ONLY: compiles the SIMT kernels of sound_bubble_b200/csrc for the host (see cuda_emu.h).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT="$HERE/libsoundbubble_emu.so"
mkdir -p "$HERE/build"
OBJS=""
for f in sb_host sb_lstm sb_lstm_tc sb_lstm_tcp sb_frontend sb_frontend_tc sb_backend sb_convlstm sb_attn sb_attn_tc sb_net sb_pipe sb_prepare sb_train sb_train_tc; do
  s="$ROOT/sound_bubble_b200/csrc/$f.cu"
  o="$HERE/build/$f.o"
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ "$ROOT/sound_bubble_b200/csrc/sb_common.cuh" -nt "$o" ] || [ "$ROOT/include/soundbubble.h" -nt "$o" ] || [ "$HERE/cuda_emu.h" -nt "$o" ] || [ "$ROOT/sound_bubble_b200/csrc/sb_lstm.cuh" -nt "$o" ]; then
    g++ -std=c++20 -O2 -fPIC -DSB_EMU -x c++ -I"$HERE" -I"$ROOT/include" -I"$ROOT/sound_bubble_b200/csrc" -Wno-unknown-pragmas -c "$s" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
o="$HERE/build/cuda_emu.o"
g++ -std=c++20 -O2 -fPIC -DSB_EMU -I"$HERE" -c "$HERE/cuda_emu.cpp" -o "$o"
g++ -shared -o "$OUT" $OBJS "$o" -lpthread
echo "built $OUT"
