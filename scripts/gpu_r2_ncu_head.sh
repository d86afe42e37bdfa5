#!/bin/bash
# ncu --set full of the four fused-head kernels (narrow at the cfg1 shape, wide at the cfg2 shape) + summary
TAG=${1:-r2aa}
mkdir -p gpurun_out
for k in fce_fwd_kernel fce_bwd_kernel; do
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/head_probe.py --iters 2 > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"
done
for k in fce_fwd_wide_kernel fce_bwd_wide_kernel; do
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${k}_$TAG \
      python scripts/head_probe.py 2048 17000 256 --iters 2 > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "ncu $k exit $?"
done
python scripts/ncu_summary.py gpurun_out/fce_fwd_kernel_$TAG.ncu-rep gpurun_out/fce_bwd_kernel_$TAG.ncu-rep gpurun_out/fce_fwd_wide_kernel_$TAG.ncu-rep gpurun_out/fce_bwd_wide_kernel_$TAG.ncu-rep > gpurun_out/${TAG}_ncu_full_flash_ce.json 2>/dev/null
python - <<PY
import json
for r in json.load(open('gpurun_out/${TAG}_ncu_full_flash_ce.json')):
    print(r['kernel'][:28], 'us', round(r.get('time_us',0),1), 'tensor% act', r.get('tensor_pipe_pct_active'), 'elapsed', r.get('tensor_pipe_pct_elapsed'), 'dram MB r/w', round(r.get('dram_bytes_read',0)/1e6,1), round(r.get('dram_bytes_write',0)/1e6,1), 'regs', r.get('regs'))
PY
