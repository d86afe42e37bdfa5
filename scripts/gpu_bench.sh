#!/bin/bash
# bench line + ncu launch list of the same command (cold-cache, serialised: compare SHARES only)
mkdir -p gpurun_out
TAG=${1:-r1}
shift
python bench.py --steps 30 --warmup 5 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -n 5 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather-probe "$@" > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_bench_$TAG.log
