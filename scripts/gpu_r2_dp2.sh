#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --no-gather-probe 2>gpurun_out/dp2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value']), 'e2e', round(d['e2e']['value']), 'enq', d['host_enqueue_ms_per_step'], d.get('graph'))"; grep "sessrec graph" gpurun_out/dp2.err | sort | uniq -c | sort -rn | head -4; }
run SESSREC_GRAPH_DEBUG=1
run SESSREC_GRAPH=0
run SESSREC_PDL=0 SESSREC_GRAPH_DEBUG=1
run SESSREC_PDL=0 SESSREC_GRAPH=0
run SESSREC_PDL=2
