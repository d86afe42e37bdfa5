#!/bin/bash
# A/B of the single-thread issue guard (SRK_ISSUE_MODE, csrc/umma.cuh): parity of the tensor-core kernels on the shipped build,
# then head / GEMM timings for the shipped build and for rebuilt variants (nvcc on the box; the box copy is thrown away).
TAG=${1:-r2u}
mkdir -p gpurun_out
probe() {
  for shape in "512 43097 96" "2048 17000 256" "512 100000 128"; do timeout -s KILL 200 python scripts/head_probe.py $shape 2>&1 | grep "L2 flushed"; done
  timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --also cfg2 2>/dev/null | python scripts/show_bench.py /dev/stdin 2>/dev/null | head -4
}
timeout -s KILL 900 python -m pytest tests/test_gpu_flash_ce.py tests/test_gpu_umma.py tests/test_gpu_models.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
echo "== shipped build (mode 2)"; probe
for m in ${MODES:-0 1}; do
  touch sessionrec-pytorch_b200/csrc/flash_ce.cu sessionrec-pytorch_b200/csrc/umma_gemm.cu
  SESSREC_NVCC_EXTRA=-DSRK_ISSUE_MODE=$m python sessionrec-pytorch_b200/build.py > /dev/null 2>gpurun_out/${TAG}_build_$m.err || { echo "build mode $m failed"; continue; }
  echo "== mode $m"; probe
done
