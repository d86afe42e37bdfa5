"""Converged-run goldens: HR@20 / MRR@20 per epoch of the UNMODIFIED reference `TrainRunner.train`
(`/root/reference/src/utils/train.py:84-127`) on `datasets/sample`, written to `tests/golden/convergence_golden.json`.

TEST INFRASTRUCTURE; build container only (needs /root/reference):   python -m oracle.make_convergence_golden [names...]

Per model the reference is trained exactly as its scripts do (`src/scripts/main_msgifsr.py:133-186`, `main_niser.py`:
AugmentedDataset of the sample sessions, sequential train order, Adam lr 1e-3, L2 1e-4 through `fix_weight_decay`,
StepLR(3, 0.1), early stop on MRR and HR) - the model code, collate code and the training loop are the reference's own
files, imported as they are over `oracle/dgl_shim`.  Runs recorded:

  * `p0`        dropout 0 (deterministic): the run the GPU parity test reproduces epoch by epoch;
  * `p0_t1`     the same run on ONE host thread: MKL picks other reduction orders, so the difference to `p0` is the
                reference's own floating-point spread - the floor of any cross-implementation tolerance;
  * `stock_s*`  the script's default dropout with three seeds: the seed-to-seed spread of the metric the north star
                bounds by +-0.001.

`evaluate` is wrapped (not modified) to record its return values at full precision; nothing else is touched.
"""
import contextlib
import io
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch as th

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_import  # noqa: E402

GOLD = ROOT / 'tests' / 'golden'

# name: (model, embedding dim, layers, batch size, stock dropout, patience, max epochs)
RUNS = {
    'srgnn_cfg0': ('SRGNN', 256, 1, 32, 0.1, 3, 8),            # BASELINE.json configs[0]
    'niser': ('NISER', 64, 2, 128, 0.5, 2, 12),                # main_niser.py defaults
    'msgifsr_k1': ('MSGIFSR', 96, 1, 512, 0.1, 3, 12),         # start.sh: --order 1 --num-layers 1, cfg1's d
    'msgifsr_k3': ('MSGIFSR', 96, 1, 512, 0.1, 3, 12, 3),      # the argparse default order (main_msgifsr.py:84), no --extra / --fusion
}


def loaders(R, name, bs, train_sessions, test_sessions, order=1):
    if name == 'MSGIFSR':
        fn = R.collate.collate_fn_factory_ccs((R.collate.seq_to_ccs_graph,), order=order)
    else:
        fn = R.collate.collate_fn_factory(R.collate.seq_to_session_graph)
    out = []
    for sess in (train_sessions, test_sessions):
        ds = R.AugmentedDataset(np.array(sess, dtype=object))
        # SequentialSampler order (`main_msgifsr.py:156`); metrics do not depend on the test order
        out.append([fn([ds[i] for i in range(lo, min(lo + bs, len(ds)))]) for lo in range(0, len(ds), bs)])
    return out


def reseed_init(m, seed):
    """Weights as a function of the seed alone: the reference's own `reset_parameters()` (uniform(-1/sqrt(d), 1/sqrt(d)) in
    `parameters()` order, `srgnn.py:126-129`, `msgifsr.py:224-227`) re-run right after seeding, then the constructor's
    alpha / beta constants (`msgifsr.py:213-216`).  A drop-in with the same parameter order reproduces it bit for bit without
    replaying the RNG draws of the sub-module constructors (tests/test_gpu_convergence.py does exactly this)."""
    th.manual_seed(seed)
    m.reset_parameters()
    if hasattr(m, 'alpha'):
        m.alpha.data = th.zeros(m.order)
        m.alpha.data[0] = 1.0
        m.beta.data = th.tensor(1.0)


def one_run(R, name, d, L, bs, p, patience, epochs, seed, threads, train_l, test_l, V, order=1):
    th.set_num_threads(threads)
    if name == 'MSGIFSR':
        m = R.MSGIFSR(V, 'sample', d, L, dropout=p, order=order, extra=False, fusion=False)
    else:
        m = getattr(R, name)(V, d, L, p)
    reseed_init(m, 123)                 # every run starts from the same weights
    th.manual_seed(seed)                # dropout stream of this run
    np.random.seed(seed)
    runner = R.train.TrainRunner('sample', m, train_l, test_l, th.device('cpu'), lr=1e-3, weight_decay=1e-4, patience=patience)
    rec = []
    real_eval = R.train.evaluate

    def recording_eval(*a, **k):
        r = real_eval(*a, **k)
        rec.append([float(r[0]), float(r[1])])
        return r
    R.train.evaluate = recording_eval
    buf = io.StringIO()
    t0 = time.time()
    try:
        with contextlib.redirect_stdout(buf):
            mrr, hit = runner.train(epochs, 100)
    finally:
        R.train.evaluate = real_eval
    losses = [float(l.split('Loss = ')[1].split(',')[0]) for l in buf.getvalue().splitlines() if l.startswith('Batch ')]
    return dict(seed=seed, threads=threads, dropout=p, evals=rec, best_mrr=float(mrr), best_hit=float(hit),
                logged_losses=losses, seconds=round(time.time() - t0, 1))


def main():
    R = ref_import.load()
    root = ref_import.REFERENCE_ROOT / 'datasets' / 'sample'
    train_s, test_s = ref_import.read_sessions(root / 'train.txt'), ref_import.read_sessions(root / 'test.txt')
    V = int((root / 'num_items.txt').read_text().split()[0])
    names = [a for a in sys.argv[1:] if a in RUNS] or list(RUNS)
    path = GOLD / 'convergence_golden.json'
    out = json.loads(path.read_text()) if path.exists() else {}
    for key in names:
        name, d, L, bs, p, patience, epochs = RUNS[key][:7]
        order = RUNS[key][7] if len(RUNS[key]) > 7 else 1
        train_l, test_l = loaders(R, name, bs, train_s, test_s, order)
        init = None
        runs = {}
        for tag, pp, seed, thr in (('p0', 0.0, 123, 4), ('p0_t1', 0.0, 123, 1), ('stock_s123', p, 123, 4),
                                   ('stock_s124', p, 124, 4), ('stock_s125', p, 125, 4)):
            runs[tag] = one_run(R, name, d, L, bs, pp, patience, epochs, seed, thr, train_l, test_l, V, order)
            e = runs[tag]['evals']
            print(f'{key:12s} {tag:11s} epochs {len(e) - 1:2d}  best MRR {runs[tag]["best_mrr"]:.5f} HR {runs[tag]["best_hit"]:.5f}  '
                  f'({runs[tag]["seconds"]} s)', flush=True)
            out[key] = dict(model=name, V=V, d=d, layers=L, order=order, batch_size=bs, stock_dropout=p, patience=patience,
                            max_epochs=epochs, lr=1e-3, weight_decay=1e-4, init_seed=123, runs=runs,
                            note='evals[0] is the evaluation before epoch 0; evals[i] after epoch i-1: [MRR@20, HR@20]')
            path.write_text(json.dumps(out, indent=1))
    print('done ->', path)


if __name__ == '__main__':
    main()
