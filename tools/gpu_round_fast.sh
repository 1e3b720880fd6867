#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_tree_filter_gpu.py tests/test_reference_ext_gpu.py tests/test_mask_loss_heads_gpu.py tests/test_losses_gpu.py -x -q 2>&1 | tail -3
python tools/debug_bfs.py 2>&1 | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:refine_updown -s 2 -c 1 -o gpurun_out/refine_full -f python tools/debug_bfs.py > gpurun_out/ncu_refine.log 2>&1; tail -2 gpurun_out/ncu_refine.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bfs_grid -s 1 -c 1 -o gpurun_out/bfs_full -f python tools/debug_bfs.py > gpurun_out/ncu_bfs.log 2>&1; tail -2 gpurun_out/ncu_bfs.log
for c in D C; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1]); print('$c', d['value'], d['ms_per_step'], d['config']['launch'])"; done
