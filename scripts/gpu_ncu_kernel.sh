#!/bin/bash
# ncu --set full capture of one kernel (regex $1) from the bench run; report -> gpurun_out/$2.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-3} -c ${4:-1} -f -o gpurun_out/$2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather-probe > gpurun_out/ncu_$2.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/$2.ncu-rep
