#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dynconv_gpu.py -x -q 2>&1 | tail -4
timeout 300 python tools/bench_dynconv.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if '{' in l:
        name=l.split(' ',1)[0]; d=json.loads(l.split(' ',1)[1])
        print(name, {k: round(v,1) for k,v in d.items() if k in ('fwd_us','bwd_feat_us','bwd_kernel_us','cublas_tf32_fwd_us','cublas_tf32_bwd_feat_us','cublas_tf32_bwd_kernel_us')})
"
