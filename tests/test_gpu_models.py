"""End-to-end parity of the drop-in modules (all arithmetic in the sm_100a kernels) against the golden vectors of
the unmodified reference and against the CPU oracle on seeded inputs.  Tolerances: north_star's 1e-4 relative on
fp32 log-probs / loss; gradients 1e-4 of each parameter's max-norm; integer outputs (top-k ids) exact."""
import os

import numpy as np
import pytest
import torch

from oracle import models as OM
from tests.util import (RTOL, assert_close, assert_close_after_adam, assert_grad_close, assert_grad_close_robust, golden, oracle_batch,
                        oracle_params, run_oracle)

pytestmark = pytest.mark.gpu
DEV = 'cuda'
MODELS = {**golden('models_golden.pt'), **golden('models_extra_golden.pt')}     # the second file: REnorm head (--extra)
TRAINS = {**golden('train_golden.pt'), **golden('train_extra_golden.pt')}
BUILT = list(MODELS)


def make_model(pkg, c, dropout=0.0):
    from sessionrec_pytorch_b200.msgifsr import MSGIFSR
    from sessionrec_pytorch_b200.srgnn import NISER, SRGNN
    if c['model'] == 'MSGIFSR':
        m = MSGIFSR(c['V'], 'golden', c['d'], c.get('L', 1), dropout=dropout, order=c['K'], extra=c.get('extra', False),
                    fusion=c.get('fusion', False))
    else:
        m = {'SRGNN': SRGNN, 'NISER': NISER}[c['model']](c['V'], c['d'], c.get('L', 1), dropout)
    missing = m.load_state_dict(c['params'], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.to(DEV)


def make_batch(pkg, c, samples=None):
    samples = c['samples'] if samples is None else samples
    kind = 'session' if c['model'] in ('SRGNN', 'NISER') else 'ccs'
    b = pkg.SessionBatch.build([s for s, _ in samples], [l for _, l in samples], kind, c['K'])
    return b.to(DEV), torch.tensor([l for _, l in samples], dtype=torch.long, device=DEV)


def test_state_dict_keys_match_reference(pkg):
    for name, c in MODELS.items():
        from sessionrec_pytorch_b200.msgifsr import MSGIFSR
        from sessionrec_pytorch_b200.srgnn import NISER, SRGNN
        if c['model'] == 'MSGIFSR':
            m = MSGIFSR(c['V'], 'x', c['d'], c['L'], order=c['K'], extra=c.get('extra', False), fusion=c['fusion'])
        else:
            m = {'SRGNN': SRGNN, 'NISER': NISER}[c['model']](c['V'], c['d'], c['L'])
        assert list(m.state_dict().keys()) == list(c['params'].keys()), name
        for k, v in m.state_dict().items():
            assert tuple(v.shape) == tuple(c['params'][k].shape), (name, k)


@pytest.mark.parametrize('name', BUILT)
def test_forward_backward_vs_reference_golden(pkg, name):
    """forward() -> (B, V) log-probs, nll_loss, backward: the unmodified TrainRunner contract."""
    c = MODELS[name]
    m = make_model(pkg, c)
    m.train()
    b, labels = make_batch(pkg, c)
    out = m(b)
    assert out.shape == (len(c['samples']), c['V']) and out.dtype == torch.float32
    loss = torch.nn.functional.nll_loss(out, labels)
    loss.backward()
    assert_close(f'{name}.logp', out, c['out'])
    assert abs(float(loss) - c['loss']) <= RTOL * abs(c['loss'])
    top_ref, top_got = c['out'].topk(20)[1], out.detach().cpu().topk(20)[1]
    gap = (c['out'].topk(21)[0][:, 19] - c['out'].topk(21)[0][:, 20])
    safe = gap > 1e-3                      # ids are compared wherever the 20/21 boundary is not a near-tie
    assert torch.equal(top_ref[safe].sort(-1)[0], top_got[safe].sort(-1)[0])
    params = dict(m.named_parameters())
    for n, g in c['grads'].items():
        assert_grad_close(f'{name}.grad[{n}]', params[n].grad, g)
    for n, w in c['params_after_forward'].items():
        assert_close(f'{name}.{n} after forward (max_norm renorm)', m.state_dict()[n], w, rtol=1e-6)


@pytest.mark.parametrize('name', BUILT)
def test_fused_loss_path_vs_reference_golden(pkg, name):
    c = MODELS[name]
    m = make_model(pkg, c)
    m.train()
    b, _ = make_batch(pkg, c)
    loss = m.loss(b)
    loss.backward()
    assert abs(float(loss) - c['loss']) <= RTOL * abs(c['loss'])
    params = dict(m.named_parameters())
    for n, g in c['grads'].items():
        assert_grad_close(f'{name}.grad[{n}]', params[n].grad, g)


@pytest.mark.parametrize('name', ['srgnn', 'niser', 'msgifsr_k1', 'msgifsr_k1_inflate_L2', 'msgifsr_k2', 'msgifsr_k3',
                                  'msgifsr_k2_fusion', 'msgifsr_k1_extra', 'msgifsr_k2_extra_fusion'])
@pytest.mark.parametrize('p', [0.2, 0.5])
def test_dropout_with_injected_masks_vs_oracle(pkg, name, p):
    """Training mode with dropout: the oracle consumes the same counter-based masks the kernels regenerate."""
    c = MODELS[name]
    # seed picked so that no near-tie (< 1e-6) occurs in the max over GAT heads: with 20260101 + 2 the order-2 model
    # has one, the two sides pick different heads for one node and every upstream gradient moves by ~1e-3 relative
    seed = 20260111 + int(p * 10)
    m = make_model(pkg, c, dropout=p)
    m.train()
    m.set_dropout_seed(seed)
    b, labels = make_batch(pkg, c)
    loss = m.loss(b)
    loss.backward()
    prm = oracle_params(c['params'])
    kind = 'session' if c['model'] in ('SRGNN', 'NISER') else 'ccs'
    ob = oracle_batch(c['samples'], kind, c['K'])
    ref = run_oracle(c['model'], prm, ob, c['L'], c['fusion'], drop=OM.Dropout(p, True, seed), extra=c.get('extra', False))
    rl = OM.nll(ref, ob['labels'])
    rl.backward()
    assert abs(float(loss) - float(rl)) <= RTOL * abs(float(rl)), (float(loss), float(rl))
    params = dict(m.named_parameters())
    for n in c['grads']:
        assert_grad_close_robust(f'{name}.p{p}.grad[{n}]', params[n].grad, prm[n].grad)
    m.eval()
    with torch.no_grad():
        assert_close(f'{name}.eval logp', m(b), run_oracle(c['model'], oracle_params(c['params'], False), ob, c['L'], c['fusion'],
                                                                 extra=c.get('extra', False)))


EDGE = golden('models_edge_golden.pt')


@pytest.mark.parametrize('name', sorted(EDGE))
@pytest.mark.parametrize('mode', ['forward', 'loss', 'train_step'])
def test_edge_case_batches_vs_reference_golden(pkg, name, mode):
    """Only single-click sessions (no edge in the ccs graph, dummy k-gram nodes, SRGNN self-loops) and a two-session batch."""
    c = EDGE[name]
    m = make_model(pkg, c)
    m.train()
    b, labels = make_batch(pkg, c)
    if mode == 'train_step':
        m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
        loss = m.train_step(b)
        assert abs(float(loss) - c['loss']) <= RTOL * abs(c['loss'])
        return
    loss = torch.nn.functional.nll_loss(m(b), labels) if mode == 'forward' else m.loss(b)
    loss.backward()
    assert abs(float(loss) - c['loss']) <= RTOL * abs(c['loss'])
    params = dict(m.named_parameters())
    for n, g in c['grads'].items():
        assert_grad_close(f'{name}.grad[{n}]', params[n].grad, g)


@pytest.mark.parametrize('head', ['flash', 'tf32'])
@pytest.mark.parametrize('name', sorted(TRAINS))
def test_fused_train_step_trajectory_vs_reference(pkg, name, head):
    """train_step (fused fwd + CE + bwd + our Adam kernel) reproduces the reference TrainRunner's losses, final
    embedding table and evaluate() metrics.  head: 'flash' = fused bf16 x 3 scoring + CE head (gradients ~1e-5 relative),
    'tf32' = materialised logits on the 3xTF32 GEMMs (~1e-6)."""
    c = TRAINS[name]
    m = make_model(pkg, c)
    m.flash_ce = head == 'flash'
    m.train()
    m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
    for it in range(c['steps']):
        b, _ = make_batch(pkg, c, c['samples'][it * c['bs']:(it + 1) * c['bs']])
        loss = m.train_step(b)
        assert abs(float(loss) - c['losses'][it]) <= RTOL * abs(c['losses'][it]), (it, float(loss), c['losses'][it])
    emb = 'embeddings.weight' if c['model'] == 'MSGIFSR' else 'embedding.weight'
    if head == 'tf32':
        assert_close(f'{name}.final_embedding', m.state_dict()[emb], c['final_embedding'], rtol=RTOL, floor=0.1)   # Adam steps are lr-sized: 1e-5 abs
    else:
        assert_close_after_adam(f'{name}.final_embedding', m.state_dict()[emb], c['final_embedding'], 1e-3, c['steps'])
    # every other tensor of the reference's state_dict after the same steps: parameters its forward never reaches keep their
    # initial values bit for bit (torch.optim.Adam skips a parameter whose grad is None; so does the fused step)
    sd = m.state_dict()
    untouched = 0
    for n_, ref_t in c['final_state'].items():
        if torch.equal(ref_t, c['params'][n_]):
            untouched += 1
            assert torch.equal(sd[n_].cpu(), ref_t), f'{name}: {n_} must stay at its initial value'
        elif head == 'tf32':
            assert_close(f'{name}.final[{n_}]', sd[n_], ref_t, rtol=RTOL, floor=0.1)
        else:
            assert_close_after_adam(f'{name}.final[{n_}]', sd[n_], ref_t, 1e-3, c['steps'])
    assert untouched >= 5, 'the goldens hold parameters the reference never updates (dead GGNN layers, lint / linq / link ...)'
    m.eval()
    mrr = hit = n = 0
    with torch.no_grad():
        for i in range(0, len(c['test_samples']), c['bs']):
            b, labels = make_batch(pkg, c, c['test_samples'][i:i + c['bs']])
            r, h = OM.topk_metrics(m(b).cpu(), labels.cpu().numpy())
            mrr, hit, n = mrr + r, hit + h, n + b.B
    assert abs(hit / n - c['hit']) <= 1e-3 and abs(mrr / n - c['mrr']) <= 1e-3


@pytest.mark.parametrize('name', ['srgnn', 'msgifsr_k1', 'msgifsr_k1_extra'])
def test_unmodified_style_training_loop_with_torch_adam(pkg, name):
    """The reference loop body verbatim (optimizer.zero_grad / model(*inputs) / nll_loss / backward / step) with
    torch.optim.Adam + fix_weight_decay groups runs on the drop-in module and follows the golden trajectory."""
    c = TRAINS[name]
    m = make_model(pkg, c)
    m.train()
    named = list(m.named_parameters())
    dec, no = OM.decay_split([n for n, _ in named])
    pm = dict(named)
    opt = torch.optim.Adam([{'params': [pm[n] for n in dec]}, {'params': [pm[n] for n in no], 'weight_decay': 0}],
                           lr=1e-3, weight_decay=1e-4)
    for it in range(c['steps']):
        b, labels = make_batch(pkg, c, c['samples'][it * c['bs']:(it + 1) * c['bs']])
        opt.zero_grad()
        scores = m(b)
        assert not torch.isnan(scores).any()
        loss = torch.nn.functional.nll_loss(scores, labels)
        loss.backward()
        opt.step()
        assert abs(loss.item() - c['losses'][it]) <= RTOL * abs(c['losses'][it]), (it, loss.item(), c['losses'][it])


@pytest.mark.parametrize('cfg', ['cfg1', 'cfg2', 'cfg3', 'cfg4'])
def test_full_size_configs_vs_oracle(pkg, cfg):
    """BASELINE.json shapes (synthetic sessions): loss and gradients against the CPU oracle, plus size-independent
    properties: rows of exp(logp) sum to 1, loss() == nll(forward()), catalog renorm is idempotent."""
    from sessionrec_pytorch_b200.synthetic import CONFIGS, SessionSampler
    k = CONFIGS[cfg]
    torch.manual_seed(5)
    c = dict(model=k['model'], V=k['V'], d=k['d'], L=k['layers'], K=1, fusion=False)
    from sessionrec_pytorch_b200.msgifsr import MSGIFSR
    from sessionrec_pytorch_b200.srgnn import NISER, SRGNN
    if k['model'] == 'MSGIFSR':
        m = MSGIFSR(k['V'], 'x', k['d'], k['layers'], dropout=0.0, order=1, extra=False, fusion=False)
        with torch.no_grad():
            m.embeddings.weight[::5] *= 2.0            # exercise max_norm on a fifth of the catalog
    else:
        m = {'SRGNN': SRGNN, 'NISER': NISER}[k['model']](k['V'], k['d'], k['layers'], 0.0)
    sd = {n: v.clone() for n, v in m.state_dict().items()}
    m = m.to(DEV).train()
    B = k['B']                                  # full BASELINE batch (cfg2: 2048 sessions)
    seqs, labels = SessionSampler(k['V'], seed=123).sessions(B)
    kind = 'session' if k['model'] in ('SRGNN', 'NISER') else 'ccs'
    b = pkg.SessionBatch.build(seqs, labels, kind, 1).to(DEV)
    lab = torch.tensor(labels, dtype=torch.long, device=DEV)
    out = m(b)
    loss = torch.nn.functional.nll_loss(out, lab)
    loss.backward()
    assert_close('rows of exp(logp) sum to 1', out.detach().double().exp().sum(-1).float(), torch.ones(B), rtol=2e-5)
    grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    fused = m.loss(b)
    assert abs(float(fused) - float(loss)) <= 1e-6 * abs(float(loss))
    prm = oracle_params(sd)
    ob = oracle_batch(list(zip(seqs, labels)), kind, 1)
    torch.set_num_threads(8)
    ref = run_oracle(k['model'], prm, ob, k['layers'])
    rl = OM.nll(ref, ob['labels'])
    rl.backward()
    assert abs(float(loss) - float(rl)) <= RTOL * abs(float(rl)), (float(loss), float(rl))
    assert_close(f'{cfg}.logp', out, ref)
    for n, g in grads.items():
        if prm[n].grad is not None:
            assert_grad_close(f'{cfg}.grad[{n}]', g, prm[n].grad)


@pytest.mark.parametrize('p', [0.0, 0.3])
def test_native_step_matches_staged_composition(pkg, p):
    """csrc/step.cu (one C call per training step) against the stage-by-stage Python composition of the same
    kernels: same losses and same parameters after a few Adam steps (dropout masks included)."""
    c = TRAINS['msgifsr_k1']
    models = []
    for native in (True, False):
        m = make_model(pkg, c, dropout=p)
        m.train()
        m.native_step = native
        m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
        losses = []
        for it in range(4):
            m.set_dropout_seed(777 + it)
            b, _ = make_batch(pkg, c, c['samples'][it * c['bs']:(it + 1) * c['bs']])
            losses.append(float(m.train_step(b)))
        models.append((m, losses))
    (m1, l1), (m2, l2) = models
    for a, b_ in zip(l1, l2):
        assert abs(a - b_) <= 2e-6 * abs(b_), (l1, l2)
    for (n, q1), (_, q2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert_close(f'native vs staged {n}', q1, q2, rtol=1e-4, floor=0.5)   # Adam turns 1e-9 gradient noise (atomics) into lr-sized steps


@pytest.mark.parametrize('name', ['srgnn', 'niser'])
@pytest.mark.parametrize('p', [0.0, 0.3])
@pytest.mark.parametrize('head', ['flash', 'tf32'])
def test_native_srgnn_step_matches_staged_composition(pkg, name, p, head):
    """csrc/step_srgnn.cu (one C call per SRGNN / NISER training step, dead GGNN layers on a side stream) against the
    stage-by-stage Python composition of the same kernels: same losses, same parameters after a few Adam steps."""
    c = TRAINS[name]
    models = []
    for native in (True, False):
        m = make_model(pkg, c, dropout=p)
        m.train()
        m.native_step, m.flash_ce = native, head == 'flash'
        m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
        losses = []
        for it in range(4):
            m.set_dropout_seed(999 + it)
            b, _ = make_batch(pkg, c, c['samples'][it * c['bs']:(it + 1) * c['bs']])
            losses.append(float(m.train_step(b)))
        models.append((m, losses))
    (m1, l1), (m2, l2) = models
    for a, b_ in zip(l1, l2):
        assert abs(a - b_) <= 2e-6 * abs(b_), (l1, l2)
    for (n, q1), (_, q2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert_close(f'native vs staged {n}', q1, q2, rtol=1e-4, floor=0.5)


@pytest.mark.parametrize('name', ['srgnn', 'msgifsr_k1'])
def test_optimizer_state_survives_reflattening_and_checkpointing(pkg, name):
    """`load_state_dict(assign=True)` (or model.to / p.data = ...) invalidates the flat parameter views; the next
    train_step must carry lr / weight decay / step count / both Adam moments over instead of silently starting a default
    optimizer, and optimizer_state_dict() / load_optimizer_state_dict() must resume a run bit for bit."""
    c = TRAINS[name]
    batches = [make_batch(pkg, c, c['samples'][it * c['bs']:(it + 1) * c['bs']])[0] for it in range(4)]

    def fresh():
        m = make_model(pkg, c)
        m.train()
        m.configure_optimizer(lr=3e-3, weight_decay=2e-4)
        return m
    ref = fresh()
    for b in batches:
        ref.train_step(b)
    # (1) re-flatten in the middle of the run
    m1 = fresh()
    m1.train_step(batches[0])
    m1.train_step(batches[1])
    m1.load_state_dict({k: v.clone() for k, v in m1.state_dict().items()}, assign=True)      # parameters are new tensors now
    m1.train_step(batches[2])
    m1.train_step(batches[3])
    assert m1._opt['lr'] == 3e-3 and m1._opt['weight_decay'] == 2e-4 and m1._opt['step'] == 4
    # (2) checkpoint after two steps, resume in a new module
    m2 = fresh()
    m2.train_step(batches[0])
    m2.train_step(batches[1])
    osd, msd = m2.optimizer_state_dict(), {k: v.clone() for k, v in m2.state_dict().items()}
    m3 = make_model(pkg, c)
    m3.train()
    m3.load_state_dict(msd)
    m3.load_optimizer_state_dict(osd)
    m3.train_step(batches[2])
    m3.train_step(batches[3])
    for (n, p), (_, q1), (_, q3) in zip(ref.named_parameters(), m1.named_parameters(), m3.named_parameters()):
        assert_close(f'reflatten {n}', q1, p, rtol=1e-4, floor=0.5)
        assert_close(f'resume {n}', q3, p, rtol=1e-4, floor=0.5)


@pytest.mark.parametrize('d', [256, 128])
def test_native_msgifsr_step_wide_embedding(pkg, d):
    """BASELINE configs[4] embedding width (d = 256: scores materialised on the 3xTF32 GEMM, GAT projections on the
    CUDA-core kernel) and d = 128 (largest width of the fused head): native step == staged composition, dropout included."""
    from sessionrec_pytorch_b200.msgifsr import MSGIFSR
    from sessionrec_pytorch_b200.synthetic import SessionSampler
    V, B = 2500, 384
    res = []
    for native in (True, False):
        torch.manual_seed(4)
        m = MSGIFSR(V, 'x', d, 1, dropout=0.2, order=1, extra=False, fusion=False).to(DEV).train()
        m.native_step = native
        m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
        smp = SessionSampler(V, seed=12)
        losses = []
        for it in range(3):
            seqs, labels = smp.sessions(B)
            m.set_dropout_seed(41 + it)
            losses.append(float(m.train_step(pkg.SessionBatch.build(seqs, labels, 'ccs', 1).to(DEV))))
        res.append((m, losses))
    (m1, l1), (m2, l2) = res
    for a, b_ in zip(l1, l2):
        assert abs(a - b_) <= 5e-6 * abs(b_), (l1, l2)
    for (n, q1), (_, q2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert_close(f'native vs staged {n}', q1, q2, rtol=1e-4, floor=0.5)


@pytest.mark.parametrize('K,d,L,p', [(2, 32, 1, 0.0), (3, 96, 1, 0.2), (3, 32, 2, 0.2), (4, 64, 1, 0.2), (2, 256, 1, 0.2)])
def test_native_order_k_step_matches_staged_composition(pkg, K, d, L, p):
    """MSGIFSR of order K > 1 (k-gram node types, SemanticExpander, intra / inter relations, multi-order read-out): the general
    native step (csrc/step_k.cu, ONE C call) == the staged composition of the same kernels, dropout masks included; the staged
    composition itself is pinned by the reference's goldens (msgifsr_k2 / msgifsr_k3 trajectories)."""
    from sessionrec_pytorch_b200.msgifsr import MSGIFSR
    from sessionrec_pytorch_b200.synthetic import SessionSampler
    V, B = 2500, 300
    res = []
    for native in (True, False):
        torch.manual_seed(5)
        m = MSGIFSR(V, 'x', d, L, dropout=p, order=K, extra=False, fusion=False).to(DEV).train()
        m.native_step = native
        m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
        smp = SessionSampler(V, seed=13)
        losses = []
        for it in range(3):
            seqs, labels = smp.sessions(B)
            if it == 2:
                seqs = [q[:1] for q in seqs[:B // 2]] + seqs[B // 2:]       # many single-click sessions: dummy k-gram nodes
            m.set_dropout_seed(51 + it)
            b = pkg.SessionBatch.build(seqs, labels, 'ccs', K).to(DEV)
            assert m._native_k_ok(b) or not native
            losses.append(float(m.train_step(b)))
        res.append((m, losses))
    (m1, l1), (m2, l2) = res
    for a, b_ in zip(l1, l2):
        assert abs(a - b_) <= 5e-6 * abs(b_), (l1, l2)
    # The two paths are not bitwise identical (atomics in the one-launch scatter-add and the split-K products; at d = 256 the
    # staged composition sends its big products to the 3xTF32 tensor-core GEMM, the native step keeps them in fp32) and Adam
    # turns a 1e-9 difference of a near-zero gradient into a visible step: see assert_close_after_adam
    for (n, q1), (_, q2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert_close_after_adam(f'native vs staged {n}', q1, q2, 1e-3, 3, rtol=1e-4, floor=0.5, max_frac=1e-4)


@pytest.mark.parametrize('model,d,L', [('SRGNN', 256, 1), ('NISER', 64, 2), ('SRGNN', 96, 2)])
def test_native_srgnn_step_with_tensor_core_projections(pkg, model, d, L):
    """Shapes where the read-out / GGNN projections of the native step go to the tcgen05 GEMM (srk_tc_gemm) and, for d = 256,
    the head runs on the materialised 3xTF32 scores: native step == staged composition, dropout masks included."""
    from sessionrec_pytorch_b200.srgnn import NISER, SRGNN
    from sessionrec_pytorch_b200.synthetic import SessionSampler
    V, B = 3000, 640
    smp = SessionSampler(V, seed=11)
    res = []
    for native in (True, False):
        torch.manual_seed(3)
        m = {'SRGNN': SRGNN, 'NISER': NISER}[model](V, d, L, 0.2).to(DEV).train()
        m.native_step = native
        m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
        smp2 = SessionSampler(V, seed=11)
        losses = []
        for it in range(3):
            seqs, labels = smp2.sessions(B)
            m.set_dropout_seed(31 + it)
            losses.append(float(m.train_step(pkg.SessionBatch.build(seqs, labels, 'session', 1).to(DEV))))
        res.append((m, losses))
    (m1, l1), (m2, l2) = res
    for a, b_ in zip(l1, l2):
        assert abs(a - b_) <= 5e-6 * abs(b_), (l1, l2)
    for (n, q1), (_, q2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert_close(f'native vs staged {n}', q1, q2, rtol=1e-4, floor=0.5)


@pytest.mark.parametrize('p,inject,whole', [(0.0, 0, 0), (0.2, 0, 0), (0.2, 3, 0), (0.0, 0, 1), (0.2, 0, 1), (0.2, 3, 1)])
def test_native_step_graph_replay_matches_plain_launches(pkg, p, inject, whole):
    """The native step with its backward half replayed as ONE CUDA graph (kernel-node parameters rewritten per batch:
    different batch shapes, pointers, dropout seeds) against the same steps issued as plain launches: same losses, same
    parameters.  inject > 0: one replay hits a (forced) kernel-sequence mismatch half-way and must finish the step with
    plain launches.  whole: forward + backward + optimizer as ONE graph (srk_set_graph_whole)."""
    from sessionrec_pytorch_b200._lib import lib
    L = lib().functions
    c = TRAINS['msgifsr_k1']
    nsteps = 12
    res = []
    try:
        for graphs in (1, 0):
            L['srk_set_graph_mode'](graphs)
            L['srk_set_graph_whole'](1 if (graphs and whole) else 0)
            L['srk_graph_inject_mismatch'](inject if graphs else 0)      # one replay must fall back half-way
            g0, f0 = L['srk_graph_launches'](), L['srk_graph_fallbacks']()
            m = make_model(pkg, c, dropout=p)
            m.train()
            m.configure_optimizer(lr=1e-3, weight_decay=1e-4)
            losses = []
            for it in range(nsteps):
                m.set_dropout_seed(4242 + it)
                lo = (it % 3) * c['bs']
                # batches of different sizes: the graph is captured on one shape and replayed on others
                b, _ = make_batch(pkg, c, c['samples'][lo:lo + c['bs'] - 3 * (it % 2)])
                losses.append(float(m.train_step(b)))
            res.append((m, losses, L['srk_graph_launches']() - g0, L['srk_graph_fallbacks']() - f0))
    finally:
        L['srk_set_graph_mode'](2)                      # back to the default (auto: always for data-parallel steps, measured for a single rank)
        L['srk_set_graph_whole'](2)
        L['srk_graph_inject_mismatch'](0)
    (m1, l1, n1, fb1), (m2, l2, n2, _) = res
    # two batch sizes = two graphs; each takes two warm-up steps, then capture (+ launch) and replays
    assert n1 >= nsteps - 4 - (1 if inject else 0) and n2 == 0, (n1, n2)
    assert fb1 == (1 if inject else 0), fb1
    for a, b_ in zip(l1, l2):
        assert abs(a - b_) <= 2e-6 * abs(b_), (l1, l2)
    for (n, q1), (_, q2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert_close(f'graph vs plain {n}', q1, q2, rtol=1e-4, floor=0.5)


def test_native_step_two_layers_and_sgemm_head(pkg):
    c = MODELS['msgifsr_k1_inflate_L2']
    outs = []
    for native, tc in ((True, True), (False, True), (True, False)):
        m = make_model(pkg, c, dropout=0.2)
        m.train()
        m.native_step, m.use_tensor_cores = native, tc
        m.set_dropout_seed(5)
        m.configure_optimizer()
        b, _ = make_batch(pkg, c)
        loss = float(m.train_step(b))
        outs.append((loss, m.embeddings.weight.detach().clone()))
    for loss, w in outs[1:]:
        assert abs(loss - outs[0][0]) <= 1e-5 * abs(outs[0][0])
        # (True, False) compares the fused bf16 x 3 head with the fp32 FFMA head: one Adam step, see assert_close_after_adam
        assert_close_after_adam('embedding after 1 step', w, outs[0][1], 1e-3, 1, rtol=1e-4, floor=0.5)


@pytest.mark.parametrize('name', ['srgnn', 'niser', 'msgifsr_k1', 'msgifsr_k3'])
def test_fused_topk_matches_reference_topk_ids(pkg, name):
    """model.topk (fused scoring + radix top-k, no log-prob materialisation) returns exactly the ids of
    `logits.topk(20)` on the reference's golden output (rows with a near-tie at the 20/21 boundary excluded)."""
    from sessionrec_pytorch_b200.train import evaluate
    c = MODELS[name]
    m = make_model(pkg, c)
    m.eval()
    b, labels = make_batch(pkg, c)
    got = m.topk(b, k=20).cpu()
    vals, ref = c['out'].topk(21)
    gaps = (vals[:, :-1] - vals[:, 1:]).min(-1)[0]
    safe = gaps > 1e-4
    assert safe.sum() >= 5
    assert torch.equal(got[safe], ref[safe][:, :20])
    mrr, hit = evaluate(m, [([b], labels)], DEV)
    r, h = OM.topk_metrics(c['out'], labels.cpu().numpy())
    assert abs(hit - h / b.B) < 1e-9 and abs(mrr - r / b.B) < 1e-6
