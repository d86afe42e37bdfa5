#!/bin/bash
# Round-2 multi-GPU visit (gpurun --gpus 2): NCCL parity tests, then DP / catalog-sharded bench lines at N=2.
TAG=${1:-r2d}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short -p no:cacheprovider -rs > gpurun_out/${TAG}_pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_dist.log
tail -n 40 gpurun_out/${TAG}_pytest_dist.log
for mode in dp shard; do
  for wl in default; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py \
        --gpus $N --parallelism $mode --no-gather-probe > gpurun_out/${TAG}_bench_${mode}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${mode}_${N}gpu.err
    echo "bench $mode exit $?"; tail -c 300 gpurun_out/${TAG}_bench_${mode}_${N}gpu.err
    python scripts/show_bench.py gpurun_out/${TAG}_bench_${mode}_${N}gpu.json 2>/dev/null | head
  done
done
# the 1-GPU anchors of both workloads on the same box
timeout 600 python bench.py --no-gather-probe --no-cpu-baseline --also cfg4 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
python scripts/show_bench.py gpurun_out/${TAG}_bench_1gpu.json 2>/dev/null | head
