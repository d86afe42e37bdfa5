#!/bin/bash
# single GPU: plain launches vs backward-half graph replay (with / without programmatic edges inside the graph)
mkdir -p gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo; uptime
run() { echo "== $*"; env "$@" timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --workload ${WL:-cfg1} --also "" 2>gpurun_out/ab3.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']), 'e2e', round(d['e2e']['value']), 'enq', d['host_enqueue_ms_per_step'], d.get('graph'))"; }
for rep in 1 2; do
run A=plain
run SESSREC_GRAPH=1
run SESSREC_GRAPH=1 SESSREC_PDL=2
done
WL=cfg2 run A=plain
WL=cfg2 run SESSREC_GRAPH=1 SESSREC_PDL=2
