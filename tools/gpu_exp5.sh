#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_compat.py tests/test_gpu_fragment.py -q -m gpu -k "lcc or engine or batch or wcc_opt or save or vertex_map" > gpurun_out/exp5_pytest.log 2>&1
tail -3 gpurun_out/exp5_pytest.log
B="python bench.py --no-cpu-baseline --sweep none --steps 5 --warmup 2"
$B > gpurun_out/exp5_bfs.json 2>&1
for v in 1 0; do
  GL_L2_PERSIST=$v $B --app pagerank > gpurun_out/exp5_pr_push_l2$v.json 2>&1
  GL_L2_PERSIST=$v $B --app pagerank --pr-pull > gpurun_out/exp5_pr_pull64_l2$v.json 2>&1
  GL_L2_PERSIST=$v $B --app pagerank --pr-f32 > gpurun_out/exp5_pr_pull32hub_l2$v.json 2>&1
  GL_L2_PERSIST=$v $B --app sssp > gpurun_out/exp5_sssp_l2$v.json 2>&1
  GL_L2_PERSIST=$v $B --app wcc > gpurun_out/exp5_wcc_l2$v.json 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/exp5_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('exp5_')[1], 'ms/query %.4f'%d['ms_per_step'], 'e2e %.3f'%d['e2e']['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['config']['ms_per_superstep'][:6])
    except Exception as e:
        print(f, 'ERR', open(f).read()[-300:])
PY
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size if hasattr(p,'L2_cache_size') else None)
import ctypes
PY
# ncu: launch list of the default bench + full sets of the kernels the roofline lines name
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bfs_s24_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sweep none > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bfs_fused -c 1 -o gpurun_out/r02_k_bfs_fused python bench.py --steps 1 --warmup 1 --no-cpu-baseline --sweep none > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pr_pull_hub -s 2 -c 1 -o gpurun_out/r02_k_pr_pull_hub python bench.py --app pagerank --pr-f32 --steps 1 --warmup 1 --no-cpu-baseline --sweep none > gpurun_out/ncu_c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_sssp_s24_launches.csv python bench.py --app sssp --steps 1 --warmup 1 --no-cpu-baseline --sweep none > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
