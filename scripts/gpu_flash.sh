#!/bin/bash
# GPU visit for the fused head: (1) flash CE kernel tests under a short timeout (a hang must not eat the box),
# (2) the whole GPU suite (with the flash head if (1) passed, else with it disabled), (3) smoke + bench line.
mkdir -p gpurun_out
TAG=${1:-r1j}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 10 300 python -m pytest tests/test_gpu_flash_ce.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/flash.log 2>&1
FL=$?
echo "flash exit $FL" >> gpurun_out/flash.log
grep -E "fwd B=|bwd d|passed|failed|Error|error|exit|relative error map" gpurun_out/flash.log | head -80
if [ $FL -ne 0 ]; then export SESSREC_NO_FLASH_CE=1; echo "FLASH CE DISABLED for the rest of this visit"; fi
timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_flash_ce.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -n 3 gpurun_out/bench_$TAG.err
# in-pipeline stage times of the native step (debug aid: synchronises every step, NOT a bench number)
SESSREC_STEP_TIMING=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gather-probe > /dev/null 2> gpurun_out/stages_$TAG.err
grep "step timing" gpurun_out/stages_$TAG.err | tail -n 3
