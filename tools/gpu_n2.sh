#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 5 --warmup 2 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo bench rc=$?
tail -c 600 gpurun_out/n2_bench.err
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "parity and 2" > gpurun_out/n2_pytest.log 2>&1
tail -3 gpurun_out/n2_pytest.log
