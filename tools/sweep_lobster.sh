#!/bin/bash
cd "$(dirname "$0")/.."
for v in "$@"; do
  so=litiv_b200/liblitiv_b200.so; [ "$v" != base ] && so=exp_build/lib_$v.so
  for cfg in "320x240 1" "1920x1080 1" "1920x1080 3"; do set -- $cfg; echo "$v $1 c$2 $(LVB_SO=$PWD/$so python tools/bench_streams.py --algo lobster --streams 1 --size $1 --channels $2 --steps 300 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['fps_per_stream']), 'fps', round(d['ms_per_round']*1e3,1),'us')")"; done
done
