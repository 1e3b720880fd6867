#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ['metric','value','unit','n_gpus','steps','warmup','ms_per_step','gpu_launches','clocks']}, d['roofline']['frac'], d['roofline']['traffic'], d['e2e'], d['cpu_baseline']['value'])
PY
