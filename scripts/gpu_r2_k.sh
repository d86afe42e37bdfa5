#!/bin/bash
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_convergence.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-220
timeout -s KILL 300 python bench.py --no-gather-probe --workload cfg1k3 --also "" 2>gpurun_out/${TAG}_bench.err | python scripts/show_bench.py /dev/stdin | cut -c1-300
tail -3 gpurun_out/${TAG}_bench.err | cut -c1-300
