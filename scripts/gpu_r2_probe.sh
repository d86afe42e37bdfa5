#!/bin/bash
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_readout_fused.py -m gpu -q --tb=short -p no:cacheprovider -k "scatter or readout or gemm_forms" 2>&1 | tail -5
timeout -s KILL 300 python -m pytest tests/test_gpu_models.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5
for wl in cfg2 cfg1; do
  timeout -s KILL 200 python scripts/step_probe.py --workload $wl 2>&1 | tail -5
  SESSREC_STEP_TIMING=2 timeout -s KILL 200 python scripts/step_probe.py --workload $wl --steps 3 2>&1 | grep "step timing" | tail -3
done
SESSREC_GRAPH=1 timeout -s KILL 300 python bench.py --also '' --no-cpu-baseline --no-gather-probe > gpurun_out/${TAG}_bench_graph1.json 2> gpurun_out/${TAG}_bench_graph1.err
python scripts/show_bench.py gpurun_out/${TAG}_bench_graph1.json
timeout -s KILL 300 python bench.py --also '' --no-cpu-baseline --no-gather-probe > gpurun_out/${TAG}_bench_graph_auto.json 2> gpurun_out/${TAG}_bench_graph_auto.err
python scripts/show_bench.py gpurun_out/${TAG}_bench_graph_auto.json
