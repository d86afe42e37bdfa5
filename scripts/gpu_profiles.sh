#!/bin/bash
# Evidence for profiles/: launch list of the bench command, ncu --set full of the fused head kernels and of the
# gather / scatter-add kernels (stress shape).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
TAG=${1:-r1n}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gather-probe > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list exit $?"
for k in fce_fwd_kernel fce_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/head_probe.py --iters 2 > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"
done
for k in gather_fwd_kernel scatter_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/gs_probe.py --stress-only > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"
done
python scripts/gs_probe.py > gpurun_out/gs_probe_$TAG.json 2> gpurun_out/gs_probe_$TAG.err; cat gpurun_out/gs_probe_$TAG.json | head -c 1500
ls -la gpurun_out/*.ncu-rep
