#!/bin/bash
# Round-2 GPU visit: parity tests (all, no -x), then the bench line + optional launch list.  Usage: gpu_r2_visit.sh TAG [pytest args]
TAG=${1:-r2c}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs -s "$@" > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|error|FAILED|ERROR|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -40
grep -E "best MRR|dropout .*best HR" gpurun_out/${TAG}_pytest_gpu.log | head -20
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 400 gpurun_out/${TAG}_bench.err
python scripts/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -60
