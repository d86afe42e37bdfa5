#!/bin/bash
# Final-build multi-GPU lines at N ranks: data parallel cfg1 (default path and without graph replay), NISER cfg3, catalog-sharded cfg4
TAG=${1:-r2q}; N=${2:-8}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout -s KILL 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-gather-probe ${EXTRA} > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err; echo "$name N=$N exit $?"; python scripts/show_bench.py gpurun_out/${TAG}_${name}_${N}gpu.json 2>/dev/null | cut -c1-260; }
run dp_cfg1 A=1
run dp_cfg1_nograph SESSREC_GRAPH=0
EXTRA="--workload cfg3" run dp_cfg3 A=1
EXTRA="--parallelism shard" run shard_cfg4 A=1
