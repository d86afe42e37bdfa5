#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3 4 6; do
  SESSREC_HEAD_CHUNKS=$c python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gather-probe > gpurun_out/sweep_chunks_$c.json 2>gpurun_out/sweep_chunks_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/sweep_chunks_$c.json'))
print('chunks', $c, 'ms/step', d['ms_per_step'], 'sessions/s', d['value'], 'e2e', d['e2e']['value'], 'enqueue ms', d['host_enqueue_ms_per_step'])
PY
done
