#!/bin/bash
timeout 600 python -m pytest tests/test_mask_loss_heads_gpu.py -x -q 2>&1 | tail -15
