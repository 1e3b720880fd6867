#!/bin/bash
# one GPU visit: the GPU test suite, the per-warp timeline of the headline kernel, A/B bench of the two schedules, op timings
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/trace_wq.py 2>&1 | tail -36
bash tools/ab_onepass.sh 2>&1 | tail -6
python tools/bench_ops.py > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err; tail -30 gpurun_out/bench_ops.json
