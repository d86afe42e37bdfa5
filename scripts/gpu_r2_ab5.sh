#!/bin/bash
# emulate a slow host: the benchmark shares ONE core with a busy loop; then the unloaded machine
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout -s KILL 300 $PIN python bench.py --no-cpu-baseline --no-gather-probe --workload ${WL:-cfg1} --also "" 2>gpurun_out/ab5.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']), 'e2e', round(d['e2e']['value']), 'enq', d['host_enqueue_ms_per_step'], d.get('graph'))"; grep "auto:" gpurun_out/ab5.err | sort | uniq -c | head -4; }
PIN="" run SESSREC_GRAPH_DEBUG=1 UNLOADED=1
(taskset -c 3 timeout 120 python -c "while True: pass" &)
sleep 1
PIN="taskset -c 3" run SESSREC_GRAPH_DEBUG=1
PIN="taskset -c 3" run SESSREC_GRAPH=0
timeout -s KILL 300 python -m pytest tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -k "native or graph or train" 2>&1 | tail -2
