#!/bin/bash
# ncu --set full of the fused head kernels (one launch each) + event timings of the same probe, not under the profiler
mkdir -p gpurun_out
TAG=${1:-r1j}
python scripts/head_probe.py > gpurun_out/head_probe_$TAG.txt 2>&1; cat gpurun_out/head_probe_$TAG.txt | tail -n 6
for k in fce_fwd_kernel fce_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/head_probe.py --iters 2 > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"; ls -la gpurun_out/${k}_$TAG.ncu-rep
done
