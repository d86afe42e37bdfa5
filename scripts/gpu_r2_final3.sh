#!/bin/bash
# Evidence visit of the FINAL build: smoke, all parity tests, default bench line (every workload with the reference beside it)
TAG=${1:-r2s}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|error|FAILED|ERROR|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -10
SESSREC_GRAPH_DEBUG=1 timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
echo "bench exit $?"; grep "auto:" gpurun_out/${TAG}_bench_default.err | head -5
python scripts/show_bench.py gpurun_out/${TAG}_bench_default.json 2>/dev/null | cut -c1-300
