#!/bin/bash
# Multi-GPU measurement visit: gpurun --gpus N -- bash scripts/gpu_r2_scale.sh TAG N
TAG=${1:-r2s}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
run() {   # name, extra bench args...
  name=$1; shift
  timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py \
      --gpus $N --no-gather-probe "$@" > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err
  echo "$name N=$N exit $?"; tail -c 200 gpurun_out/${TAG}_${name}_${N}gpu.err | tr '\n' ' '; echo
  python scripts/show_bench.py gpurun_out/${TAG}_${name}_${N}gpu.json 2>/dev/null | head -3
}
run dp_cfg1
run dp_cfg3 --workload cfg3
run shard_cfg4 --parallelism shard
SESSREC_NATIVE_COMM=0 run dp_cfg1_torchcomm --workload cfg1
