#!/bin/bash
# Round-2 final evidence visit (1 GPU): smoke, all parity tests, the default bench line (+ the reference arm), launch lists of
# cfg1 / cfg2, ncu --set full of the gather / scatter kernels (stress shape), gather / scatter probe.
TAG=${1:-r2q}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|error|FAILED|ERROR|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -10
timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
echo "bench exit $?"; tail -c 300 gpurun_out/${TAG}_bench_default.err
python scripts/show_bench.py gpurun_out/${TAG}_bench_default.json 2>/dev/null | cut -c1-300
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "reference arm exit $?"; head -c 400 gpurun_out/${TAG}_bench_reference.json; echo
for wl in cfg1 cfg2; do
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches_${wl}.csv \
    python bench.py --workload $wl --also '' --steps 2 --warmup 3 --no-cpu-baseline --no-gather-probe > gpurun_out/${TAG}_ncu_bench_${wl}.log 2>&1
echo "launch list $wl exit $?"
done
for k in gather_tma_kernel scatter_bwd_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/gs_probe.py --stress-only > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"
done
python scripts/ncu_summary.py gpurun_out/gather_tma_kernel_$TAG.ncu-rep gpurun_out/scatter_bwd_kernel_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_gather_scatter.json 2>/dev/null
python scripts/gs_probe.py > gpurun_out/${TAG}_gs_probe.json 2> gpurun_out/${TAG}_gs_probe.err; head -c 1500 gpurun_out/${TAG}_gs_probe.json
