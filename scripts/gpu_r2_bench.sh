#!/bin/bash
# Round-2 measurement visit (1 GPU): bench line (cfg1 headline + cfg2 block, reference arm from oracle/_ref), launch lists.
TAG=${1:-r2b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/${TAG}_bench.err
python scripts/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -60
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
    python bench.py --workload cfg2 --also '' --steps 2 --warmup 3 --no-cpu-baseline --no-gather-probe > gpurun_out/${TAG}_ncu_bench_cfg2.log 2>&1
echo "launch list cfg2 exit $?"
