"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: data-parallel gradient averaging reproduces the
full-batch gradient of the oracle, and the catalog-sharded cross-entropy reproduces the full soft-max."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return dict(ret)


def _load():
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    from __graft_entry__ import load_package
    load_package()
    from sessionrec_pytorch_b200 import parallel
    return parallel


def _dp_job(rank, world):
    par = _load()
    from oracle import models as OM
    from tests.util import golden, oracle_batch, oracle_params
    c = golden('models_golden.pt')['niser']
    seqs = [s for s, _ in c['samples']]
    labels = [l for _, l in c['samples']]
    s_r, l_r = par.shard_batch(seqs, labels, rank, world)
    p = oracle_params(c['params'])
    ob = oracle_batch(list(zip(s_r, l_r)), 'session', 1)
    loss = OM.nll(OM.srgnn_forward(p, ob, niser=True), ob['labels'])
    loss.backward()
    names = [n for n, v in p.items() if v.grad is not None]
    flat = torch.cat([p[n].grad.reshape(-1) for n in names])
    scale = par.allreduce_mean_grads(flat, weight=len(s_r))          # weighted: shards may be unequal
    flat *= scale
    ref = torch.cat([c['grads'][n].reshape(-1) for n in names])
    return float((flat - ref).abs().max() / ref.abs().max())


def test_data_parallel_gradient_equals_full_batch_gradient():
    out = _run(_dp_job)
    assert len(out) == 2 and max(out.values()) < 2e-5, out


def _sharded_job(rank, world):
    par = _load()
    g = torch.Generator().manual_seed(0)
    B, V, d = 37, 1001, 16
    s = torch.nn.functional.normalize(torch.randn(B, d, generator=g), dim=-1)
    E = torch.nn.functional.normalize(torch.randn(V, d, generator=g), dim=-1)
    labels = torch.randint(0, V, (B,), generator=g)
    lo, hi = par.shard_slice(V, rank, world)
    res = {}
    for bound, scale in ((12.0, 12.0), (None, 1.7)):
        sr = s.clone().requires_grad_(True)
        z = scale * (sr @ E.t())
        full = torch.nn.functional.cross_entropy(z, labels)
        full.backward()
        z_loc = scale * (s @ E[lo:hi].t())
        loss, dz, lse = par.sharded_ce(z_loc, labels, lo, hi, bound=bound)
        ds = scale * (dz @ E[lo:hi])
        dist.all_reduce(ds)
        res[str(bound)] = (abs(float(loss) - float(full)), float((ds - sr.grad).abs().max()),
                           float((lse - torch.logsumexp(z.detach(), 1)).abs().max()))
    return res


def test_catalog_sharded_cross_entropy_equals_full_softmax():
    out = _run(_sharded_job)
    for r in out.values():
        for k, (dl, dg, dlse) in r.items():
            assert dl < 2e-6 and dg < 1e-6 and dlse < 2e-5, (k, dl, dg, dlse)


def test_shard_slice_covers_everything():
    par = _load()
    for n in (0, 1, 7, 512, 43097):
        for w in (1, 2, 3, 8):
            spans = [par.shard_slice(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
