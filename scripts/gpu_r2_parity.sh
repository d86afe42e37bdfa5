#!/bin/bash
# Round-2 parity visit (run with gpurun --gpus 2): every -m gpu test incl. the edge cases and the 2-GPU NCCL tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs "$@" > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest_gpu.log
tail -n 80 gpurun_out/r2a_pytest_gpu.log
