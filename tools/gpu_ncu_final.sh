#!/bin/bash
# final build: launch list of the default bench command + one full-set capture of k_bfs_fused
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bfs_s24_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sweep none > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bfs_fused -s 3 -c 1 -o gpurun_out/r02_k_bfs_fused_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sweep none --no-parity > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_bfs_s24_launches_final.csv
tail -2 gpurun_out/ncu_a.log | cut -c1-300
