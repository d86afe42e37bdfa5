"""Does the CUDA-graph replay of the native step's backward half hold at a full-size configuration (tensor-core GAT
projections on)?  Prints graph launches / fallbacks over a few steps with varying batches.  SESSREC_GRAPH_DEBUG=1 makes
the library say why a capture or an update pass was refused."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('SESSREC_GRAPH_DEBUG', '1')
import torch  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='cfg1')
ap.add_argument('--steps', type=int, default=12)
args = ap.parse_args()
pkg = load_package()
from sessionrec_pytorch_b200._lib import lib  # noqa: E402
from sessionrec_pytorch_b200.msgifsr import MSGIFSR  # noqa: E402
from sessionrec_pytorch_b200.synthetic import CONFIGS, SessionSampler  # noqa: E402

cfg = CONFIGS[args.workload]
L = lib().functions
torch.manual_seed(0)
m = MSGIFSR(cfg['V'], 'probe', cfg['d'], cfg['layers'], dropout=cfg['dropout'], order=1, extra=False, fusion=False).to('cuda')
m.train()
m.configure_optimizer()
smp = SessionSampler(cfg['V'], seed=1)
batches = [pkg.SessionBatch.build_flat(*smp.batch(cfg['B']), 'ccs', 1).to('cuda') for _ in range(4)]
for mode in (1, 0):
    L['srk_set_graph_mode'](mode)
    g0, f0 = L['srk_graph_launches'](), L['srk_graph_fallbacks']()
    losses = [float(m.train_step(batches[i % 4])) for i in range(args.steps)]
    print(f'graph mode {mode}: {args.steps} steps, graph launches {L["srk_graph_launches"]() - g0}, fallbacks '
          f'{L["srk_graph_fallbacks"]() - f0}, loss {losses[0]:.4f} -> {losses[-1]:.4f}', flush=True)
