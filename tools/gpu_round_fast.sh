#!/bin/bash
mkdir -p gpurun_out
for v in "" len12 len16 ""; do
  if [ -n "$v" ]; then export BXS_LIB_PATH=$PWD/boxinstseg_b200/lib/libboxseg_b200_$v.so; else unset BXS_LIB_PATH; fi
  python -m pytest tests/test_boxinst_gpu.py -x -q -k "single_pass or full_size" 2>&1 | tail -1
  python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v] step us', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3))"
done
unset BXS_LIB_PATH
python tools/debug_bfs.py 2>&1 | tail -12
python -m pytest tests/test_losses_gpu.py -x -q 2>&1 | tail -2
