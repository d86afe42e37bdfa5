#!/bin/bash
# DP attribution at N ranks: native comm (all-reduce behind the graph), native comm without graph replay, torch.distributed
TAG=${1:-r2n}; N=${2:-2}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-gather-probe > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err; echo "$name N=$N exit $?"; python scripts/show_bench.py gpurun_out/${TAG}_${name}_${N}gpu.json 2>/dev/null | cut -c1-260; }
run dp_native A=1
run dp_native_nograph SESSREC_GRAPH=0
run dp_torch SESSREC_NATIVE_COMM=0
