#!/bin/bash
# round 2, step B: the VMC-loop test (reference vs b200 backend) + the previously failing mirror test
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_vmc_gpu.py tests/test_gpu_parity.py -q -m gpu -x -k "vmc or mirror_known" > gpurun_out/pytest_vmc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_vmc.log
tail -60 gpurun_out/pytest_vmc.log
