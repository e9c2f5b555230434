#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_compat.py tests/test_gpu_multi.py tests/test_gpu_apps.py -q -m gpu -x > gpurun_out/exp2_pytest.log 2>&1
echo rc=$? >> gpurun_out/exp2_pytest.log
tail -15 gpurun_out/exp2_pytest.log
