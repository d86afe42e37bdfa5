#!/bin/bash
# programmatic dependent launch + launch diet: full parity suite, then A/B of the bench line
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/${TAG}_pytest.log
run() { echo "== $*"; env "$@" timeout -s KILL 300 python bench.py --no-cpu-baseline --no-gather-probe --also "${ALSO:-}" 2>gpurun_out/${TAG}_bench.err | python scripts/show_bench.py /dev/stdin 2>/dev/null | head -3; grep "host profile" gpurun_out/${TAG}_bench.err; grep "sessrec graph" gpurun_out/${TAG}_bench.err | sort | uniq -c | sort -rn | head -3; }
ALSO=cfg2 run SESSREC_PDL=1 SESSREC_HOST_PROFILE=1
ALSO=cfg2 run SESSREC_PDL=0
run SESSREC_PDL=1 SESSREC_STREAMS=0
run SESSREC_PDL=2 SESSREC_GRAPH_WHOLE=1 SESSREC_GRAPH_DEBUG=1
run SESSREC_PDL=2 SESSREC_GRAPH_WHOLE=1 SESSREC_BENCH_BATCHES=1 SESSREC_GRAPH_DEBUG=1
run SESSREC_PDL=1 SESSREC_GRAPH_WHOLE=1 SESSREC_BENCH_BATCHES=1
