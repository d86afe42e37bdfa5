#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
run() { env "$@" timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --no-gather-probe 2>gpurun_out/ab2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], round(d['value']), 'e2e', round(d['e2e']['value']), 'enq', d['host_enqueue_ms_per_step'])"; grep -c "update pass failed" gpurun_out/ab2.err; grep "sessrec graph" gpurun_out/ab2.err | sort | uniq -c | head -5; }
for rep in 1 2; do
run SESSREC_GRAPH_DEBUG=1 A=native
run SESSREC_GRAPH_DEBUG=1 SESSREC_NATIVE_COMM=0
done
run SESSREC_GRAPH=0 A=native_nograph
run SESSREC_GRAPH=0 SESSREC_NATIVE_COMM=0
