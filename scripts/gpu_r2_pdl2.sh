#!/bin/bash
# PDL footprint thresholds at cfg1 / cfg2, stage timing, dead-layer diagnosis
TAG=${1:-r2x}
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --workload ${WL:-cfg1} --also "" 2>gpurun_out/${TAG}_bench.err | python scripts/show_bench.py /dev/stdin 2>/dev/null | head -3 | cut -c1-220; grep -A12 "step timing\|stage" gpurun_out/${TAG}_bench.err | tail -${TL:-0}; }
for wl in cfg1 cfg2; do
export WL=$wl
run SESSREC_PDL=1 SESSREC_PDL_MAX_THREADS=303104 SESSREC_PDL_MAX_SMEM_KB=8192
run SESSREC_PDL=1 SESSREC_PDL_MAX_THREADS=151552 SESSREC_PDL_MAX_SMEM_KB=8192
run SESSREC_PDL=1 SESSREC_PDL_MAX_THREADS=75776 SESSREC_PDL_MAX_SMEM_KB=4096
run SESSREC_PDL=1 SESSREC_PDL_MAX_THREADS=37888 SESSREC_PDL_MAX_SMEM_KB=2048
done
WL=cfg2 run SESSREC_PDL=0 SESSREC_BENCH_NO_DEAD=1
WL=cfg2 run SESSREC_PDL=1 SESSREC_BENCH_NO_DEAD=1
WL=cfg2 SESSREC_PDL=0 SESSREC_STEP_TIMING=2 timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --workload cfg2 --also "" --steps 6 --warmup 3 2>&1 >/dev/null | grep -v "^\[bench\]" | tail -30
