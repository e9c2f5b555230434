#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "parity and 2 or one_device" > gpurun_out/n2_pytest.log 2>&1
tail -3 gpurun_out/n2_pytest.log
bash tools/gpu_n8.sh 2
