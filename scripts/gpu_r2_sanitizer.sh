#!/bin/bash
# compute-sanitizer memcheck over the kernels that are new in round 2 (small shapes: memcheck is ~50x slower)
mkdir -p gpurun_out
OUT=gpurun_out/r2p_sanitizer.md
echo "# compute-sanitizer --tool memcheck (B200, round-2 final build)" > $OUT
run() {
  echo "" >> $OUT; echo "\`$*\`:" >> $OUT
  timeout -s KILL 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "$@" -m gpu -q -p no:cacheprovider -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | tail -4 >> $OUT
}
run tests/test_gpu_flash_ce.py -k "128-128-256 or 100-1000-256 or 130-129-256 or topk or outside"
run tests/test_gpu_kernels.py -k "gather_scatter or deterministic and not 70000 or topk_rows"
run tests/test_gpu_umma.py -k "tc_gemm and (333 or 500 or 100-36)"
run tests/test_gpu_models.py -k "native_srgnn_step_matches or optimizer_state or edge_case"
cat $OUT
# scatter chunk sweep on the stress shape (experiment switch)
for c in 4 8 16; do
  SESSREC_SCATTER_CHUNK=$c python scripts/gs_probe.py --stress-only 2>/dev/null | python -c "
import sys,json
for r in json.loads(sys.stdin.read()): print('chunk $c', r['kernel'][:34], r['achieved'], r['frac'], r['ms'])"
done
