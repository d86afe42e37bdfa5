#!/bin/bash
# accumulating dE form of the fused-head backward (one [V, d] buffer instead of per-tile partial tables): parity + A/B
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_flash_ce.py tests/test_gpu_convergence.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/deat_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/deat_pytest.log | cut -c1-220
run() { echo "== $*"; env "$@" timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --workload ${WL:-cfg2} --also "" 2>gpurun_out/deat.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']), 'e2e', round(d['e2e']['value']), 'enq', d['host_enqueue_ms_per_step'], d['roofline']['ms_per_kernel'])"; }
for rep in 1 2; do
for wl in ${WLS:-cfg1 cfg3}; do
WL=$wl run A=atomic
WL=$wl run SESSREC_FCE_DE_ATOMIC=0
done
done
