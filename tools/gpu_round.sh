#!/bin/bash
# one GPU visit: the GPU test suite, the per-warp timeline of the headline kernel, A/B bench of the two schedules
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/trace_wq.py 2>&1 | tail -36
bash tools/ab_onepass.sh 2>&1 | tail -6
