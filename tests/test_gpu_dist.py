"""Two-GPU NCCL tests (skipped on a single-GPU box): data-parallel train_step and the catalog-sharded scoring head
against the single-GPU result on the same global batch."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import sys
        from pathlib import Path
        sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
        from __graft_entry__ import load_package
        pkg = load_package()
        from sessionrec_pytorch_b200 import parallel
        from tests.test_gpu_models import TRAINS, make_model
        from tests.util import golden
        name = 'msgifsr_k1' if mode != 'shard_srgnn' else 'srgnn'
        c = golden('models_golden.pt')[name]
        dev = f'cuda:{rank}'
        m = make_model(pkg, c).to(dev)
        m.train()
        seqs = [s for s, _ in c['samples']]
        labels = [l for _, l in c['samples']]
        kind = 'session' if c['model'] in ('SRGNN', 'NISER') else 'ccs'
        if mode == 'dp':
            s_r, l_r = parallel.shard_batch(seqs, labels, rank, world)
            b = pkg.SessionBatch.build(s_r, l_r, kind, c['K']).to(dev)
            m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
            m.train_step(b, dist.group.WORLD)
            ret[rank] = {n: p.detach().cpu() for n, p in m.named_parameters()}
        else:
            b = pkg.SessionBatch.build(seqs, labels, kind, c['K']).to(dev)
            m.shard_catalog(dist.group.WORLD)
            loss = m.loss(b)
            loss.backward()
            ret[rank] = dict(loss=float(loss), grads={n: (None if p.grad is None else p.grad.cpu()) for n, p in m.named_parameters()})
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


def _run(mode, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    return dict(ret)


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')


@needs2
@pytest.mark.parametrize('mode', ['shard', 'shard_srgnn'])
def test_catalog_sharded_head_matches_single_gpu(pkg, mode):
    from tests.util import assert_grad_close, golden
    out = _run(mode)
    c = golden('models_golden.pt')['msgifsr_k1' if mode == 'shard' else 'srgnn']
    for r in (0, 1):
        assert abs(out[r]['loss'] - c['loss']) <= 1e-4 * abs(c['loss']), (r, out[r]['loss'], c['loss'])
        for n, g in c['grads'].items():
            assert_grad_close(f'rank{r}.{n}', out[r]['grads'][n], g)


@needs2
def test_data_parallel_step_matches_full_batch_step(pkg):
    """Two ranks with half the batch each + one all-reduce == one rank with the full batch (batch halves are equal)."""
    from tests.test_gpu_models import make_batch, make_model
    from tests.util import assert_close, golden
    out = _run('dp')
    c = golden('models_golden.pt')['msgifsr_k1']
    m = make_model(pkg, c)
    m.train()
    m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
    b, _ = make_batch(pkg, c)
    m.train_step(b)
    for n, p in m.named_parameters():
        assert_close(f'dp.{n}', out[0][n], p, rtol=1e-4, floor=0.5)
        assert torch.equal(out[0][n], out[1][n]), f'replicas diverged on {n}'
