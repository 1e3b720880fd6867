#!/bin/bash
# A/B of the two single-pass schedules on ONE box: warp-granular queue (default) vs CTA-granular (BXS_ONEPASS_CTA=1)
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/ab_wq.json 2> gpurun_out/ab_wq.err
BXS_ONEPASS_CTA=1 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/ab_cta.json 2> gpurun_out/ab_cta.err
python - <<'PY'
import json
for tag in ('wq', 'cta'):
    try:
        d = json.loads(open(f'gpurun_out/ab_{tag}.json').read().strip().splitlines()[-1])
        k = d['roofline']['kernels']
        print(tag, 'step us', d['ms_per_step'] * 1e3, 'frac', d['roofline']['frac'], 'fwd us', k['single_pass_forward(onepass_main+onepass_finalize)']['us'],
              'step cabi', k['single_pass_step_c_abi']['us'], 'e2e', d['e2e']['value'])
    except Exception as e:
        print(tag, 'FAILED', e)
        print(open(f'gpurun_out/ab_{tag}.err').read()[-2000:])
PY
