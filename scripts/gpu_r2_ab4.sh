#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/ab4_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/ab4_pytest.log | cut -c1-200
run() { echo "== $*"; env "$@" timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --workload ${WL:-cfg1} --also "" 2>gpurun_out/ab4.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']), 'e2e', round(d['e2e']['value']), 'enq', d['host_enqueue_ms_per_step'], d.get('graph'))"; grep "auto:" gpurun_out/ab4.err | sort | uniq -c | head -4; }
run SESSREC_GRAPH_DEBUG=1
run SESSREC_GRAPH=0
run SESSREC_GRAPH=1
WL=cfg2 run SESSREC_GRAPH_DEBUG=1
WL=cfg1k3 run SESSREC_GRAPH_DEBUG=1
# a loaded host: 14 busy threads beside the benchmark
for i in $(seq 14); do (timeout 100 python -c "while True: pass" &) ; done
sleep 1
run SESSREC_GRAPH_DEBUG=1 LOADED=1
run SESSREC_GRAPH=0 LOADED=1
