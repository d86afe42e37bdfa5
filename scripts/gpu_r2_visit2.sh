#!/bin/bash
# Round-2 GPU visit: all parity tests, bench (cfg1 headline + cfg2 + cfg4), launch lists of cfg1 / cfg2.
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|error|FAILED|ERROR|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -30
timeout -s KILL 900 python bench.py --also cfg2,cfg4 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 400 gpurun_out/${TAG}_bench.err
python scripts/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -60
for wl in cfg1 cfg2; do
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches_${wl}.csv \
    python bench.py --workload $wl --also '' --steps 2 --warmup 3 --no-cpu-baseline --no-gather-probe > gpurun_out/${TAG}_ncu_bench_${wl}.log 2>&1
echo "launch list $wl exit $?"
done
