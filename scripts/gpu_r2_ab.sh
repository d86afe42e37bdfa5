#!/bin/bash
# A/B of single switches at cfg1 on ONE box (box-to-box variance is ~5 %): each line = one bench run
mkdir -p gpurun_out
run() { env "$@" timeout -s KILL 300 python bench.py --also '' --no-cpu-baseline --no-gather-probe 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], round(d['value']), 'e2e', round(d['e2e']['value']), 'enq', d['host_enqueue_ms_per_step'])"; }
for rep in 1 2; do
run A=default
run SESSREC_GATHER_TMA=0
run SESSREC_SCATTER_ATOMICS=1
run SESSREC_GATHER_TMA=0 SESSREC_SCATTER_ATOMICS=1
done
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "scatter or gather" 2>&1 | tail -3
python scripts/gs_probe.py 2>/dev/null | python -c "
import sys,json
for r in json.loads(sys.stdin.read()): print(r['kernel'][:40], r['shape'][:12], r['achieved'], r['frac'], r['ms'])"
