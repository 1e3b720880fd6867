#!/bin/bash
# one GPU visit for the headline kernel: parity tests, timeline, A/B bench, ncu full capture of the main kernel
mkdir -p gpurun_out
python -m pytest tests/test_boxinst_gpu.py tests/test_config_a_gpu.py -x -q 2>&1 | tail -5
python tools/trace_wq.py 2>&1 | tail -40
bash tools/ab_onepass.sh 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wq_ -s 20 -c 2 -o gpurun_out/wq_full -f python tools/raw_loop1.py 16 > gpurun_out/ncu_wq.log 2>&1
tail -3 gpurun_out/ncu_wq.log
