#!/bin/bash
# one `ncu --set full` capture of the catalog GEMM kernel (3 launches = fwd, dS, dE of one step)
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 9 -c 3 -f -o gpurun_out/umma_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather-probe > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/umma_$TAG.ncu-rep
