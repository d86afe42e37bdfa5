#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 240 python -m pytest tests/test_gpu_umma.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/umma.log 2>&1
echo "umma exit $?" >> gpurun_out/umma.log
grep -E "max\|d\||passed|failed|Error|error|exit" gpurun_out/umma.log | head -60
