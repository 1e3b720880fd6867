#!/bin/bash
timeout 600 python -m pytest tests/test_solo_targets_gpu.py -x -q 2>&1 | grep -v "^  File\|site-packages" | tail -40
