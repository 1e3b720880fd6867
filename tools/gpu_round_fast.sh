#!/bin/bash
timeout 600 python tools/probe_d_levels.py 2>&1 | grep -v Warning | tail -8
