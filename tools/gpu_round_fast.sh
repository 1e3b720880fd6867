#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_boxinst_gpu.py tests/test_config_a_gpu.py -x -q 2>&1 | tail -3
python tools/trace_wq.py 2>&1 | tail -36
bash tools/ab_onepass.sh 2>&1 | tail -6
