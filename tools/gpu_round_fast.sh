#!/bin/bash
timeout 600 python -m pytest tests/test_tree_filter_gpu.py tests/test_reference_ext_gpu.py tests/test_mask_loss_heads_gpu.py -x -q 2>&1 | tail -3
BXS_LIB_PATH=boxinstseg_b200/lib/libboxseg_b200_treetrace.so timeout 300 python tools/trace_bfs.py 2>&1 | tail -8
timeout 300 python tools/debug_bfs.py 2>&1 | tail -14
