#!/bin/bash
# Round-2 evidence: bench lines of every config, the ncu launch list of the bench command, ncu --set full captures of the
# dominant kernels (exported to CSV on the box: gpurun_out/ may not exceed 64 MiB), the per-warp timeline.
# Run under gpurun; copy what should be judged from gpurun_out/ into profiles/.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python bench.py --steps 200 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
for c in C D E; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; done
python tools/trace_wq.py > gpurun_out/r2_wq_timeline.txt 2>&1
rm -f gpurun_out/wq_trace.npy
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for c in C D E; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
python tools/bench_dynconv.py > gpurun_out/r2_bench_dynconv.log 2>&1; cp gpurun_out/bench_dynconv.json gpurun_out/r2_bench_dynconv.json
python tools/bench_ops.py > gpurun_out/r2_bench_ops.json 2> gpurun_out/r2_bench_ops.err
python tools/bench_f_rows.py > gpurun_out/r2_bench_f_rows.json 2> gpurun_out/r2_bench_f_rows.err    # f2 / f4 rows vs the reference's composition
# full captures -> CSV (raw page = every metric per kernel), reports deleted except the headline one
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wq_|onepass_backward' -s 24 -c 3 -o gpurun_out/r2_wq_full -f python tools/raw_loop1.py 16 > gpurun_out/ncu_a.log 2>&1
ncu -i gpurun_out/r2_wq_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_wq.csv 2>/dev/null
for cfg in "D refine_updown|refine_bwd_weight|bfs_grid|mst_cluster|lsf_forward|lsf_backward|edge_weight|prj_" "E dynconv_tf32|dynconv_wgrad|lcm_fused|resize_" "C mf_step_global|dynconv_tf32"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none -k regex:"$2" -c 14 -o gpurun_out/r2_cfg$1_full -f python bench.py --config $1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/r2_cfg$1_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_cfg$1.csv 2>/dev/null
  rm -f gpurun_out/r2_cfg$1_full.ncu-rep
done
du -sh gpurun_out; ls -la gpurun_out | tail -30
