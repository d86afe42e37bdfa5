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
        if mode.startswith('dp'):
            # dp: even halves, gradient all-reduce through torch.distributed between two C calls; dp_native: the step enqueues
            # the all-reduce itself on this library's own NCCL communicator; dp_uneven: rank 0 holds two thirds of the batch,
            # B_global is all-reduced; dp_empty: rank 1's shard is empty and it still joins the collectives
            if mode in ('dp_native', 'dp_uneven_native'):
                parallel.init_comm(dist.group.WORLD)
                m.dp_allreduce_inside = True
            else:
                m.dp_allreduce_inside = False
            n = len(seqs)
            cut = {'dp': n // 2, 'dp_native': n // 2, 'dp_uneven': 2 * n // 3, 'dp_uneven_native': 2 * n // 3, 'dp_empty': n}[mode]
            s_r, l_r = (seqs[:cut], labels[:cut]) if rank == 0 else (seqs[cut:], labels[cut:])
            b = pkg.SessionBatch.build(s_r, l_r, kind, c['K']).to(dev) if len(s_r) else None
            m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
            for _ in range(4 if 'native' in mode else 1):       # > 2 steps: the third captures the graph, the fourth replays it
                m.train_step(b, dist.group.WORLD, global_batch=n if mode in ('dp', 'dp_native') else None)
            ret[rank] = {n_: p.detach().cpu() for n_, p in m.named_parameters()}
        elif mode == 'shard_step':
            # catalog-sharded training step, all exchanges enqueued by the native step (csrc/comm.cu)
            parallel.init_comm(dist.group.WORLD)
            b = pkg.SessionBatch.build(seqs, labels, kind, c['K']).to(dev)
            m.shard_catalog(dist.group.WORLD)
            m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
            losses = [float(m.train_step(b)) for _ in range(4)]
            ret[rank] = dict(losses=losses, params={n_: p.detach().cpu() for n_, p in m.named_parameters()})
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
        p.join(180)
    hung = [p for p in procs if p.is_alive()]
    for p in hung:                       # a rank stuck in a collective must not outlive the test
        p.kill()
    assert not hung, f'{len(hung)} rank(s) still running after 180 s (killed)'
    for p in procs:
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
@pytest.mark.parametrize('mode', ['dp', 'dp_native', 'dp_uneven', 'dp_uneven_native', 'dp_empty'])
def test_data_parallel_step_matches_full_batch_step(pkg, mode):
    """Two ranks with a shard of the batch each + one all-reduce == one rank with the full batch: equal halves, uneven
    shards (the gradients are weighted by B_local / B_global), an empty shard; through torch.distributed between two C calls
    and (native) with the all-reduce enqueued by the step itself, graph replay included."""
    from tests.test_gpu_models import make_batch, make_model
    from tests.util import assert_close, golden
    out = _run(mode)
    c = golden('models_golden.pt')['msgifsr_k1']
    m = make_model(pkg, c)
    m.train()
    m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
    b, _ = make_batch(pkg, c)
    for _ in range(4 if 'native' in mode else 1):
        m.train_step(b)
    for n, p in m.named_parameters():
        assert_close(f'{mode}.{n}', out[0][n], p, rtol=1e-4, floor=0.5)
        assert torch.equal(out[0][n], out[1][n]), f'replicas diverged on {n}'


@needs2
def test_catalog_sharded_native_step_matches_single_gpu(pkg):
    """MSGIFSR training steps with the catalog rows sharded over two ranks - [2, B] soft-max statistics and dS all-reduced,
    rank-local Adam on the owned rows, owners broadcast their updated rows - against the same steps on one GPU."""
    from tests.test_gpu_models import make_batch, make_model
    from tests.util import assert_close, golden
    out = _run('shard_step')
    c = golden('models_golden.pt')['msgifsr_k1']
    m = make_model(pkg, c)
    m.train()
    m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
    b, _ = make_batch(pkg, c)
    losses = [float(m.train_step(b)) for _ in range(4)]
    for r in (0, 1):
        for a, b_ in zip(out[r]['losses'], losses):
            assert abs(a - b_) <= 1e-5 * abs(b_), (r, out[r]['losses'], losses)
        for n, p in m.named_parameters():
            assert_close(f'shard_step.rank{r}.{n}', out[r]['params'][n], p, rtol=1e-4, floor=0.5)
    for n in out[0]['params']:
        assert torch.equal(out[0]['params'][n], out[1]['params'][n]), f'replicas diverged on {n}'
