#!/bin/bash
# Builds tuning variants of liblitiv_b200.so into exp_build/ (here, cross-compiled) and, on the GPU box, times each with bench.py.
#   tools/sweep_variants.sh build "NAME:-DFLAG=1 -DOTHER=2" ...      (run in the container)
#   tools/sweep_variants.sh run NAME[@ENV=VAL,...] ...                (run under gpurun; prints scan / feedback / frame times)
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -shared -Xcompiler -fPIC -diag-suppress 550"
if [ "$mode" = build ]; then
  for v in "$@"; do
    name=${v%%:*}; defs=${v#*:}
    nvcc $FLAGS $defs -o exp_build/lib_$name.so litiv_b200/csrc/litiv_b200.cu &
  done
  wait
  ls -la exp_build/
else
  for v in "$@"; do
    name=${v%%@*}; envs=""
    if [[ "$v" == *@* ]]; then envs=$(echo "${v#*@}" | tr ',' ' '); fi
    so=litiv_b200/liblitiv_b200.so
    [ "$name" != base ] && so=exp_build/lib_$name.so
    out=$(env $envs LVB_SO=$PWD/$so python bench.py --steps 40 --warmup 5 --repeats 5 --no-cpu-baseline --no-streams64 --no-other-configs 2>&1 | tail -1)
    echo "$v $(echo "$out" | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('frame_ms=%.4f scan_ms=%.4f tail_ms=%.4f fb_ms=%.4f e2e=%.0f sync=%.0f frame_frac=%.3f sbar=%.2f' % (d['ms_per_step'], r['avg_launch_ms'], r['other_kernels'][0]['avg_launch_ms'], r['other_kernels'][1]['avg_launch_ms'], d['e2e']['value'], d['e2e']['synchronous_apply_value'], r['frame']['frac'], r['scan_depth']))
except Exception as e: print('FAILED', e)
")"
  done
fi
