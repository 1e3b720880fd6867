#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -15 > gpurun_out/sanitize_memcheck.log; tail -8 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -25 > gpurun_out/sanitize_racecheck.log; tail -25 gpurun_out/sanitize_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -12 > gpurun_out/sanitize_synccheck.log; tail -8 gpurun_out/sanitize_synccheck.log
