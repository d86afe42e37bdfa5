#!/bin/bash
# Round-2 evidence visit (1 GPU): all parity tests, the default bench line, ncu --set full captures of the head kernels
# (cfg1 narrow, cfg2 wide) and of the gather / scatter kernels (stress shape), gather/scatter probe.
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|error|FAILED|ERROR|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -30
timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -c 400 gpurun_out/${TAG}_bench.err
python scripts/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -60
if [ "$2" != "nocap" ]; then
for k in fce_fwd_kernel fce_bwd_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/head_probe.py --iters 2 > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"
done
for k in fce_fwd_wide_kernel fce_bwd_wide_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/head_probe.py 2048 17000 256 --iters 2 > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"
done
for k in gather_tma_kernel scatter_bwd_kernel; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/gs_probe.py --stress-only > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"
done
python scripts/ncu_summary.py gpurun_out/fce_fwd_kernel_$TAG.ncu-rep gpurun_out/fce_bwd_kernel_$TAG.ncu-rep gpurun_out/fce_fwd_wide_kernel_$TAG.ncu-rep gpurun_out/fce_bwd_wide_kernel_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_flash_ce.json 2>/dev/null
python scripts/ncu_summary.py gpurun_out/gather_tma_kernel_$TAG.ncu-rep gpurun_out/scatter_bwd_kernel_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_gather_scatter.json 2>/dev/null
head -c 3000 gpurun_out/${TAG}_ncu_full_flash_ce.json
fi
python scripts/gs_probe.py > gpurun_out/${TAG}_gs_probe.json 2> gpurun_out/${TAG}_gs_probe.err; head -c 2500 gpurun_out/${TAG}_gs_probe.json
