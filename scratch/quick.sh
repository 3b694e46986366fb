for n in 3 5 6; do echo "MIN_BLOCKS=$n"; LVB_SO=$PWD/scratch/lib_mb$n.so python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'])"; done
