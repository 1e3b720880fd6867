#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tree_filter_gpu.py tests/test_reference_ext_gpu.py tests/test_mask_loss_heads_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/debug_bfs.py 2>&1 | grep -A2 "^2x200x256\|^2x96x96" 
for c in D E; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; python - $c <<'PY'
import json,sys
c=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2_bench_{c}.json').read().strip().split('\n')[-1])
print(c, d['value'], d['ms_per_step'], d['gpu_reference']['value'], d['gpu_reference'].get('loss_reference'), d['gpu_reference'].get('loss_b200'))
PY
done
