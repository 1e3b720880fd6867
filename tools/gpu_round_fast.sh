#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_boxinst_gpu.py tests/test_config_a_gpu.py tests/test_tree_filter_gpu.py tests/test_reference_ext_gpu.py -x -q 2>&1 | tail -3
python tools/trace_wq.py 2>&1 | tail -40
bash tools/ab_onepass.sh 2>&1 | tail -4
python tools/bench_ops.py 2>/dev/null | grep -E "mst_|bfs_|treefilter|tree_levels"
for c in D E C; do timeout 300 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -c 1500 gpurun_out/bench_$c.json; tail -3 gpurun_out/bench_$c.err; done
