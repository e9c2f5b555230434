#!/bin/bash
# experiment: two fragment ranks on ONE GPU (time-sliced contexts + CUDA IPC), then a baseline bench
mkdir -p gpurun_out
(GL_ONE_DEVICE=1 GL_APPS=bfs_step,wcc,sssp timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py 11 > gpurun_out/onedev_a.log 2>&1; echo rc=$? >> gpurun_out/onedev_a.log)
(GL_ONE_DEVICE=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/mgpu_worker.py 11 > gpurun_out/onedev_b.log 2>&1; echo rc=$? >> gpurun_out/onedev_b.log)
nvidia-smi --query-gpu=name,compute_mode --format=csv > gpurun_out/smi.log
python bench.py > gpurun_out/bench_base.log 2>&1
tail -4 gpurun_out/onedev_a.log gpurun_out/onedev_b.log gpurun_out/bench_base.log
