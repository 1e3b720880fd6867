#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_boxinst_gpu.py tests/test_config_a_gpu.py tests/test_match_cost_gpu.py -x -q 2>&1 | tail -3
for v in "" s3r16 s3r24 s2r16; do
  if [ -n "$v" ]; then export BXS_LIB_PATH=$PWD/boxinstseg_b200/lib/libboxseg_b200_$v.so; else unset BXS_LIB_PATH; fi
  python -m pytest tests/test_boxinst_gpu.py -x -q -k "single_pass or full_size" 2>&1 | tail -1
  python bench.py --steps 400 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('variant [$v] step us', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'fwd', round(k['single_pass_forward(onepass_main+onepass_finalize)']['us'],2))"
done
