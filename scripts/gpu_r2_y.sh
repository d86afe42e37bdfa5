#!/bin/bash
TAG=${1:-r2y}
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest.log
run() { echo "== $*"; env "$@" timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --workload ${WL:-cfg1} --also "${ALSO:-}" 2>gpurun_out/${TAG}_bench.err | python scripts/show_bench.py /dev/stdin 2>/dev/null | head -3 | cut -c1-230; }
ALSO=cfg2,cfg4 run A=default
WL=cfg2 SESSREC_STEP_TIMING=2 timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --workload cfg2 --also "" --steps 3 --warmup 3 2>&1 >/dev/null | grep "step timing" | tail -3
WL=cfg1 SESSREC_STEP_TIMING=2 timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --workload cfg1 --also "" --steps 3 --warmup 3 2>&1 >/dev/null | grep "step timing" | tail -3
