#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck --print-limit 6 python tools/sanitize_small.py > gpurun_out/sanitize_synccheck.log 2>&1; grep -v "^=========     \|^  File\|^    " gpurun_out/sanitize_synccheck.log | head -40
timeout 600 python -m pytest tests/test_boxinst_gpu.py tests/test_config_a_gpu.py tests/test_tree_filter_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('step us', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'])"; done
