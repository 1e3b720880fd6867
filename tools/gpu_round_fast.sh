#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_losses_gpu.py tests/test_mask_loss_heads_gpu.py tests/test_detector_glue_gpu.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --config C --steps 10 --warmup 3 > gpurun_out/r2_bench_C.json 2> gpurun_out/r2_bench_C.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_C.json').read().strip().split('\n')[-1])
print('C', d['value'], d['ms_per_step'], d['gpu_reference']['value'], d['gpu_reference'].get('loss_reference'), d['gpu_reference'].get('loss_b200'))
PY
