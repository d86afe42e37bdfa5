#!/bin/bash
# Second evidence visit of the final build (after the order-K native step): all parity tests, default bench line, memcheck over the
# code that changed after r2p (elect issue, PDL, launch diet, pre-split operands, whole-step graph, order-K step)
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|error|FAILED|ERROR|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -10
timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
echo "bench exit $?"; tail -c 300 gpurun_out/${TAG}_bench_default.err
python scripts/show_bench.py gpurun_out/${TAG}_bench_default.json 2>/dev/null | cut -c1-300
OUT=gpurun_out/${TAG}_sanitizer.md
echo "# compute-sanitizer --tool memcheck (B200, final build of round 2)" > $OUT
run() {
  echo "" >> $OUT; echo "\`$*\`:" >> $OUT
  timeout -s KILL 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "$@" -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|FAILED" | tail -5 >> $OUT
}
run tests/test_gpu_flash_ce.py -k "128-128-256 or 100-1000-256 or 130-129-256 or topk or outside or 300-1000-96"
run tests/test_gpu_umma.py -k "tc_gemm and (333 or 500 or 100-36)"
run tests/test_gpu_models.py -k "order_k and (2-32 or 3-32) or graph_replay and not 3-1 or native_srgnn_step_with_tensor_core"
cat $OUT
