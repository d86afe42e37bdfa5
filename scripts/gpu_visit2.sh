#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r1k}
for pm in 0 2 3; do echo "L2 promo $pm"; SESSREC_FCE_L2PROMO=$pm python scripts/head_probe.py 2>&1 | grep "flushed"; done
timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_kernel'], 'cpu', d.get('cpu_baseline'))
PY
SESSREC_STREAMS=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gather-probe > gpurun_out/bench_${TAG}_nostreams.json 2> gpurun_out/bench_${TAG}_nostreams.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}_nostreams.json'))
print('NO STREAMS', {k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step')}, 'e2e', d['e2e']['value'])
PY
