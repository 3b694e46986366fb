#!/bin/bash
# PAWCS launch-configuration sweep on the GPU box: tools/sweep_pawcs.sh NAME... (variants built by tools/sweep_variants.sh build)
cd "$(dirname "$0")/.."
for v in "$@"; do
  so=litiv_b200/liblitiv_b200.so; [ "$v" != base ] && so=exp_build/lib_$v.so
  echo "$v vga   $(LVB_SO=$PWD/$so python tools/exp_pawcs.py 640 480 60 2>&1 | tail -1)"
  echo "$v 1080p $(LVB_SO=$PWD/$so python tools/exp_pawcs.py 1920 1080 60 2>&1 | tail -1)"
done
