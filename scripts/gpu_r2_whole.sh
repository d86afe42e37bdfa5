#!/bin/bash
# whole-step graph: parity of the replay, then the bench with the same batch every step (upper bound of a fixed-shape replay)
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_models.py -m gpu -q --tb=short -p no:cacheprovider -k "graph_replay or native" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/${TAG}_pytest.log
run() { echo "== $*"; env "$@" timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --also "" 2>gpurun_out/${TAG}_bench.err | python scripts/show_bench.py /dev/stdin 2>/dev/null | head -3; grep "sessrec graph" gpurun_out/${TAG}_bench.err | sort | uniq -c | head -5; }
run A=plain
run SESSREC_BENCH_BATCHES=1 A=plain_same_batch
run SESSREC_GRAPH_WHOLE=1 SESSREC_GRAPH_DEBUG=1
run SESSREC_GRAPH_WHOLE=1 SESSREC_BENCH_BATCHES=1 SESSREC_GRAPH_DEBUG=1
