#!/usr/bin/env python
"""Wall-clock and stage-level look at ONE workload's training step: back-to-back (device events) vs one step at a time
(synchronised after every step, as the end-to-end loop does).  SESSREC_STEP_TIMING=2 adds the per-stage times of the native
step on stderr."""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import build, load_package  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='cfg2')
ap.add_argument('--steps', type=int, default=20)
args = ap.parse_args()
build()
pkg = load_package()
from sessionrec_pytorch_b200.synthetic import CONFIGS, SessionSampler  # noqa: E402
cfg = dict(CONFIGS[args.workload])
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
m = bench.build_model(cfg, dev)
m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
smp = SessionSampler(cfg['V'], seed=123)
host = [pkg.SessionBatch.build_flat(*smp.batch(cfg['B']), bench.kind_of(cfg), cfg['order'], pin=True) for _ in range(8)]
res = [b.to(dev) for b in host]
for i in range(8):
    m.train_step(res[i % 8])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(args.steps):
    m.train_step(res[i % 8])
torch.cuda.synchronize()
print(f'{args.workload}: back to back, resident batches: {1e3 * (time.perf_counter() - t0) / args.steps:.3f} ms/step')
for mode in ('resident + sync', 'H2D + item'):
    ts = []
    for i in range(args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = res[i % 8] if mode.startswith('resident') else host[i % 8].to(dev, non_blocking=True)
        t1 = time.perf_counter()
        loss = m.train_step(b)
        t2 = time.perf_counter()
        if mode.startswith('resident'):
            torch.cuda.synchronize()
        else:
            loss.item()
        t3 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2))
    n = len(ts)
    print(f'{args.workload}: {mode}: copy call {1e3 * sum(t[0] for t in ts) / n:.3f} ms, train_step call {1e3 * sum(t[1] for t in ts) / n:.3f} ms, '
          f'wait {1e3 * sum(t[2] for t in ts) / n:.3f} ms, total {1e3 * sum(sum(t) for t in ts) / n:.3f} ms/step')
