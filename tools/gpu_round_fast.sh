#!/bin/bash
BXS_LIB_PATH=boxinstseg_b200/lib/libboxseg_b200_treetrace.so timeout 300 python tools/trace_bfs.py 2>&1 | tail -10
