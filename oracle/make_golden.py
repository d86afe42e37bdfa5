"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference's `src/models/*`, `src/utils/data/collate.py`, `src/utils/train.py` are imported as they are
(over `oracle/dgl_shim`, see `oracle/ref_import.py`); everything written here is an OUTPUT of that code.
Also cross-checks the restatement in `oracle/models.py` / `oracle/collate.py` against the reference on
the way and fails loudly on a mismatch, so a stale fixture can not be produced silently.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch as th

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import collate as OC  # noqa: E402
from oracle import models as OM  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLD = ROOT / 'tests' / 'golden'
V_SMALL, D_SMALL, N_SESS = 512, 16, 300


def flat_from_dgl(bg, kind, K):
    """Batched (shim) DGLGraph produced by the reference's collate -> flat batch layout of
    `oracle/collate.py` (labels are filled by the caller)."""
    out = dict(K=K, kind=kind, iid={}, seg={}, last={}, rel={})
    for k in range(1, K + 1):
        nt = '_N' if kind == 'session' else f's{k}'
        cnt = bg.batch_num_nodes(nt).numpy()
        seg = np.zeros(len(cnt) + 1, np.int64)
        np.cumsum(cnt, out=seg[1:])
        fr = bg._nframes[nt]
        out['seg'][k] = seg
        out['iid'][k] = fr['iid'].numpy().astype(np.int64)
        out['last'][k] = np.nonzero(fr['last'].numpy() == 1)[0].astype(np.int64)
        out['B'] = len(cnt)
    for (s, e, t), (src, dst) in bg._rels.items():
        if kind == 'session':
            name, st, dt = 'intra1', 1, 1
        else:
            st, dt = int(s[1:]), int(t[1:])
            name = e if e != 'inter' else (f'inter1_{dt}' if st == 1 else f'inter{st}_1')
        out['rel'][name] = (st, dt, src.numpy().astype(np.int64), dst.numpy().astype(np.int64))
    if kind == 'session':
        out['w'] = bg.edata['w'].numpy().astype(np.int64)
    return out


def ref_collate(R, samples, kind, K):
    if kind == 'session':
        fn = R.collate.collate_fn_factory(R.collate.seq_to_session_graph)
    else:
        fn = R.collate.collate_fn_factory_ccs((R.collate.seq_to_ccs_graph,), order=K)
    inputs, labels = fn(samples)
    flat = flat_from_dgl(inputs[0], kind, K)
    flat['labels'] = labels.numpy().astype(np.int64)
    return inputs, labels, flat


def _jsonable(flat):
    j = dict(B=int(flat['B']), K=int(flat['K']), kind=flat['kind'], labels=flat['labels'].tolist(),
             iid={str(k): v.tolist() for k, v in flat['iid'].items()},
             seg={str(k): v.tolist() for k, v in flat['seg'].items()},
             last={str(k): v.tolist() for k, v in flat['last'].items()},
             rel={n: [int(st), int(dt), s.tolist(), d.tolist()] for n, (st, dt, s, d) in flat['rel'].items()})
    if 'w' in flat and flat['w'] is not None:
        j['w'] = flat['w'].tolist()
    return j


def assert_same_flat(a, b, what):
    assert a['B'] == b['B'] and a['K'] == b['K'], what
    np.testing.assert_array_equal(a['labels'], b['labels'], what)
    for k in a['iid']:
        np.testing.assert_array_equal(a['iid'][k], b['iid'][k], f'{what} iid{k}')
        np.testing.assert_array_equal(a['seg'][k], b['seg'][k], f'{what} seg{k}')
        np.testing.assert_array_equal(a['last'][k], b['last'][k], f'{what} last{k}')
    assert sorted(a['rel']) == sorted(b['rel']), (what, sorted(a['rel']), sorted(b['rel']))
    for n in a['rel']:
        for i in (2, 3):
            np.testing.assert_array_equal(a['rel'][n][i], b['rel'][n][i], f'{what} {n}')
    if a.get('w') is not None:
        np.testing.assert_array_equal(a['w'], b['w'], what)


def params_of(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def oracle_params(sd, grad=True):
    p = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if grad and t.is_floating_point():
            t.requires_grad_(True)
        p[k] = t
    return p


def model_case(R, name, samples, V, d, L, K=1, fusion=False, inflate=False, seed=123, extra=False):
    kind = 'session' if name in ('SRGNN', 'NISER') else 'ccs'
    inputs, labels, flat = ref_collate(R, samples, kind, K)
    th.manual_seed(seed)
    if name == 'MSGIFSR':
        m = R.MSGIFSR(V, 'golden', d, L, dropout=0.0, order=K, extra=extra, fusion=fusion)
    else:
        m = getattr(R, name)(V, d, L, 0.0)
    if inflate:   # push some rows over norm 1 so that the max_norm renorm path is exercised
        with th.no_grad():
            w = (m.embeddings if name == 'MSGIFSR' else m.embedding).weight
            w[::3] *= 2.5
            if name == 'MSGIFSR':
                m.alpha.copy_(th.linspace(0.7, -0.2, K))
    m.train()      # dropout p = 0: train == eval arithmetic, but this is the training path
    sd0 = params_of(m)
    out = m(*inputs)
    loss = th.nn.functional.nll_loss(out, labels)
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in m.named_parameters()}
    case = dict(model=name, V=V, d=d, L=L, K=K, fusion=fusion, extra=extra, samples=[(list(map(int, s)), int(l)) for s, l in samples],
                params=sd0, out=out.detach().clone(), loss=float(loss.detach()), grads=grads,
                params_after_forward={k: v for k, v in params_of(m).items() if 'embedding' in k})
    # restatement cross-check (oracle vs reference, same weights)
    p = oracle_params(sd0)
    ob = OC.build_batch([s for s, _ in samples], [l for _, l in samples], kind, K)
    assert_same_flat(flat, ob, f'{name} K={K}')
    if name == 'MSGIFSR':
        o = OM.msgifsr_forward(p, ob, num_layers=L, fusion=fusion, extra=extra)
    else:
        o = OM.srgnn_forward(p, ob, num_layers=L, niser=(name == 'NISER'))
    ol = OM.nll(o, ob['labels'])
    ol.backward()
    err = (o - out).abs().max().item()
    gerr = 0.0
    for n, g in grads.items():
        og = p[n].grad
        if g is None:
            assert og is None or og.abs().max() == 0, f'{name}: oracle has grad for {n}, reference has none'
            continue
        gerr = max(gerr, ((og - g).abs().max() / max(g.abs().max().item(), 1e-6)).item())   # numerically-zero grads: absolute
    print(f'  {name:8s} K={K} L={L} fusion={fusion} extra={extra} inflate={inflate}: loss {float(loss.detach()):.6f}  |out-oracle| {err:.2e}  max rel grad err {gerr:.2e}')
    assert err < 2e-5 and gerr < 2e-4, 'oracle restatement disagrees with the reference'
    if name == 'MSGIFSR':
        assert th.equal(p['embeddings.weight'].detach(), m.embeddings.weight.detach()) or \
            (p['embeddings.weight'].detach() - m.embeddings.weight.detach()).abs().max() < 1e-6
    return case


def ggnn_case(R, samples, V, d, seed=7):
    inputs, labels, flat = ref_collate(R, samples, 'session', 1)
    th.manual_seed(seed)
    m = R.SRGNN(V, d, 1, 0.0)
    layer = m.layers[0]
    x = th.randn(inputs[0].num_nodes(), d, requires_grad=True)
    rnd = th.randn(inputs[0].num_nodes(), d)
    out = layer(inputs[0], x)
    (out * rnd).sum().backward()
    grads = {n: p.grad.detach().clone() for n, p in layer.named_parameters()}
    p = oracle_params({'layers.0.' + k: v for k, v in layer.state_dict().items()})
    ox = x.detach().clone().requires_grad_(True)
    oo = OM.ggnn_layer(p, 'layers.0.', flat, ox)
    (oo * rnd).sum().backward()
    err = (oo - out).abs().max().item()
    gerr = max(((p['layers.0.' + n].grad - g).abs().max() / g.abs().max()).item() for n, g in grads.items())
    gerr = max(gerr, ((ox.grad - x.grad).abs().max() / x.grad.abs().max()).item())
    print(f'  GGNN layer d={d}: |out-oracle| {err:.2e}  max rel grad err {gerr:.2e}')
    assert err < 1e-5 and gerr < 1e-4
    return dict(samples=[(list(map(int, s)), int(l)) for s, l in samples], d=d,
                params={'layers.0.' + k: v.detach().clone() for k, v in layer.state_dict().items()},
                x=x.detach().clone(), rnd=rnd, out=out.detach().clone(), dx=x.grad.detach().clone(),
                grads={'layers.0.' + n: g for n, g in grads.items()})


def train_case(R, name, samples, test_samples, V, d, K, bs, steps, seed=123, extra=False):
    """A few iterations of the reference's own TrainRunner loop body (Adam lr 1e-3, L2 1e-4 with
    fix_weight_decay) and its evaluate(); dropout 0 so that the trajectory is deterministic."""
    kind = 'session' if name in ('SRGNN', 'NISER') else 'ccs'
    th.manual_seed(seed)
    m = R.MSGIFSR(V, 'golden', d, 1, dropout=0.0, order=K, extra=extra, fusion=False) if name == 'MSGIFSR' \
        else getattr(R, name)(V, d, 1, 0.0)
    sd0 = params_of(m)
    runner = R.train.TrainRunner('golden', m, [], [], th.device('cpu'), lr=1e-3, weight_decay=1e-4, patience=3)
    p = oracle_params(sd0)
    opt = OM.make_adam({k: v for k, v in p.items() if v.is_floating_point()}, 1e-3, 1e-4)
    losses = []
    m.train()
    for it in range(steps):
        chunk = samples[it * bs:(it + 1) * bs]
        inputs, labels, flat = ref_collate(R, chunk, kind, K)
        runner.optimizer.zero_grad()
        scores = m(*inputs)
        loss = th.nn.functional.nll_loss(scores, labels)
        loss.backward()
        runner.optimizer.step()
        losses.append(float(loss.detach()))
        opt.zero_grad()
        o = OM.msgifsr_forward(p, flat, extra=extra) if name == 'MSGIFSR' else OM.srgnn_forward(p, flat, niser=(name == 'NISER'))
        ol = OM.nll(o, flat['labels'])
        ol.backward()
        opt.step()
        assert abs(float(ol.detach()) - losses[-1]) < 1e-4 * max(1.0, abs(losses[-1])), (it, float(ol.detach()), losses[-1])
    tl = [ref_collate(R, test_samples[i:i + bs], kind, K)[:2] for i in range(0, len(test_samples), bs)]
    mrr, hit = R.train.evaluate(m, tl, th.device('cpu'))
    print(f'  train {name} K={K} extra={extra}: losses {losses[0]:.5f} -> {losses[-1]:.5f}   MRR@20 {mrr:.5f} HR@20 {hit:.5f}')
    emb = 'embeddings.weight' if name == 'MSGIFSR' else 'embedding.weight'
    return dict(model=name, V=V, d=d, K=K, bs=bs, steps=steps, extra=extra, params=sd0, losses=losses,
                samples=[(list(map(int, s)), int(l)) for s, l in samples[:steps * bs]],
                test_samples=[(list(map(int, s)), int(l)) for s, l in test_samples],
                final_embedding=m.state_dict()[emb].detach().clone(), mrr=float(mrr), hit=float(hit),
                # every tensor of the reference's state_dict after the last step: parameters the forward never reaches
                # (grad None) must still sit at their initial values - torch.optim.Adam skips them
                final_state={k: v.detach().clone() for k, v in m.state_dict().items() if k != emb and v.is_floating_point()})


def extra_cases(R, small, test_small):
    """REnorm head (`--extra`, msgifsr.py:281-305): kept in its own files so that the other fixtures stay byte-stable."""
    print('models, REnorm head (V=%d, d=%d):' % (V_SMALL, D_SMALL))
    b0, b1 = small[:48], small[200:264]
    models = {
        'msgifsr_k1_extra': model_case(R, 'MSGIFSR', b0, V_SMALL, D_SMALL, 1, K=1, extra=True),
        'msgifsr_k1_extra_inflate_L2': model_case(R, 'MSGIFSR', b1, V_SMALL, D_SMALL, 2, K=1, extra=True, inflate=True),
        'msgifsr_k2_extra': model_case(R, 'MSGIFSR', b0, V_SMALL, D_SMALL, 1, K=2, extra=True),
        'msgifsr_k2_extra_fusion': model_case(R, 'MSGIFSR', b1, V_SMALL, D_SMALL, 1, K=2, extra=True, fusion=True, inflate=True),
        'msgifsr_k1_extra_d32': model_case(R, 'MSGIFSR', small[300:332], V_SMALL, 32, 1, K=1, extra=True),
    }
    th.save(models, GOLD / 'models_extra_golden.pt')
    th.save({'msgifsr_k1_extra': train_case(R, 'MSGIFSR', small, test_small, V_SMALL, D_SMALL, 1, 32, 6, extra=True)},
            GOLD / 'train_extra_golden.pt')


def edge_cases(R):
    """Degenerate batches the reference still runs: only single-click sessions (a ccs graph without any edge, every k-gram
    type reduced to its dummy node, SRGNN's self-loops), and a two-session batch (B == 1 breaks the reference's own
    `squeeze()`, msgifsr.py:308)."""
    print('models, edge-case batches (V=%d, d=%d):' % (V_SMALL, D_SMALL))
    singles = [([5], 7), ([9], 3), ([5], 2), ([100], 5), ([7], 7), ([3], 1), ([250], 9), ([11], 4)]
    two = [([4, 8, 4], 2), ([6], 1)]
    models = {
        'msgifsr_k1_single_clicks': model_case(R, 'MSGIFSR', singles, V_SMALL, D_SMALL, 1, K=1),
        'msgifsr_k2_single_clicks': model_case(R, 'MSGIFSR', singles, V_SMALL, D_SMALL, 1, K=2),
        'msgifsr_k1_extra_single_clicks': model_case(R, 'MSGIFSR', singles, V_SMALL, D_SMALL, 1, K=1, extra=True),
        'msgifsr_k1_two_sessions': model_case(R, 'MSGIFSR', two, V_SMALL, D_SMALL, 1, K=1),
        'msgifsr_k3_two_sessions': model_case(R, 'MSGIFSR', two, V_SMALL, D_SMALL, 1, K=3),
        'srgnn_single_clicks': model_case(R, 'SRGNN', singles, V_SMALL, D_SMALL, 1),
        'niser_two_sessions': model_case(R, 'NISER', two, V_SMALL, D_SMALL, 1),
    }
    th.save(models, GOLD / 'models_edge_golden.pt')


def test_split(sessions):
    t = OC.augmented_samples(sessions[N_SESS:N_SESS + 60])
    return [(s, l) for s, l in t if max(s + [l]) < V_SMALL]


def main():
    th.set_num_threads(1)
    R = ref_import.load()
    GOLD.mkdir(parents=True, exist_ok=True)
    sessions = ref_import.read_sessions(ref_import.REFERENCE_ROOT / 'datasets' / 'sample' / 'train.txt')
    if '--only-edge' in sys.argv:
        edge_cases(R)
        return
    if '--only-extra' in sys.argv:
        extra_cases(R, OC.augmented_samples(sessions[:N_SESS]), test_split(sessions))
        return
    ds = R.AugmentedDataset(np.array(sessions, dtype=object))
    all_samples = [(list(map(int, ds[i][0])), int(ds[i][1])) for i in range(len(ds))]
    mine = OC.augmented_samples(sessions)
    assert mine == all_samples, 'augmentation restatement differs from reference AugmentedDataset'
    small = OC.augmented_samples(sessions[:N_SESS])
    assert max(max(s) for s in sessions[:N_SESS]) < V_SMALL

    # ---- collate golden ---------------------------------------------------------------------
    print('collate:')
    hand = [[3, 1, 3, 6, 2, 5, 1, 2, 4, 1, 2], [250, 250, 250, 250, 3, 1, 2, 4, 1], [7], [5, 9], [4, 4], [1, 2, 1, 2, 1, 2, 1],
            [9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 9, 8, 7], [2, 2, 2, 2, 2]]
    cases = []
    for kind, K in (('session', 1), ('ccs', 1), ('ccs', 2), ('ccs', 3), ('ccs', 4)):
        for seq in hand:
            _, _, flat = ref_collate(R, [(seq, 0)], kind, K)
            assert_same_flat(flat, OC.build_batch([seq], [0], kind, K), f'{kind} {K} {seq}')
            cases.append(dict(seqs=[seq], **_jsonable(flat)))
        for lo, hi in ((0, 7), (40, 72), (100, 103)):
            chunk = all_samples[lo:hi]
            _, _, flat = ref_collate(R, chunk, kind, K)
            assert_same_flat(flat, OC.build_batch([s for s, _ in chunk], [l for _, l in chunk], kind, K), f'{kind} {K} batch')
            cases.append(dict(seqs=[s for s, _ in chunk], **_jsonable(flat)))
    # the whole sample set, restatement vs reference (not stored: checked here, re-checked natively in tests
    # against the restatement)
    for kind, K in (('session', 1), ('ccs', 1), ('ccs', 2), ('ccs', 3)):
        for lo in range(0, 4096, 512):
            chunk = all_samples[lo:lo + 512]
            _, _, flat = ref_collate(R, chunk, kind, K)
            assert_same_flat(flat, OC.build_batch([s for s, _ in chunk], [l for _, l in chunk], kind, K), f'{kind} {K} sample@{lo}')
    (GOLD / 'collate_golden.json').write_text(json.dumps(dict(
        source='reference src/utils/data/collate.py (unmodified) over oracle/dgl_shim', cases=cases)))
    print(f'  {len(cases)} cases written; restatement == reference on 8 x 512 sample prefixes for 4 graph kinds')

    # ---- model golden -----------------------------------------------------------------------
    print('models (V=%d, d=%d):' % (V_SMALL, D_SMALL))
    b0, b1 = small[:48], small[200:264]
    models = {
        'srgnn': model_case(R, 'SRGNN', b0, V_SMALL, D_SMALL, 1),
        'niser': model_case(R, 'NISER', b0, V_SMALL, D_SMALL, 1),
        'srgnn_b1_L2': model_case(R, 'SRGNN', b1, V_SMALL, D_SMALL, 2),
        'msgifsr_k1': model_case(R, 'MSGIFSR', b0, V_SMALL, D_SMALL, 1, K=1),
        'msgifsr_k1_inflate_L2': model_case(R, 'MSGIFSR', b1, V_SMALL, D_SMALL, 2, K=1, inflate=True),
        'msgifsr_k2': model_case(R, 'MSGIFSR', b0, V_SMALL, D_SMALL, 1, K=2),
        'msgifsr_k3': model_case(R, 'MSGIFSR', b1, V_SMALL, D_SMALL, 1, K=3, inflate=True),
        'msgifsr_k2_fusion': model_case(R, 'MSGIFSR', b0, V_SMALL, D_SMALL, 1, K=2, fusion=True, inflate=True),
        'msgifsr_k1_d32': model_case(R, 'MSGIFSR', small[300:332], V_SMALL, 32, 1, K=1),
    }
    th.save(models, GOLD / 'models_golden.pt')
    extra_cases(R, small, test_split(sessions))
    edge_cases(R)
    th.save(dict(ggnn_d16=ggnn_case(R, b0, V_SMALL, D_SMALL), ggnn_d32=ggnn_case(R, b1, V_SMALL, 32)), GOLD / 'ggnn_golden.pt')

    print('training trajectories:')
    test_small = test_split(sessions)
    trains = {
        'srgnn': train_case(R, 'SRGNN', small, test_small, V_SMALL, D_SMALL, 1, 32, 6),
        'niser': train_case(R, 'NISER', small, test_small, V_SMALL, D_SMALL, 1, 32, 6),
        'msgifsr_k1': train_case(R, 'MSGIFSR', small, test_small, V_SMALL, D_SMALL, 1, 32, 6),
        'msgifsr_k2': train_case(R, 'MSGIFSR', small, test_small, V_SMALL, D_SMALL, 2, 32, 4),
    }
    th.save(trains, GOLD / 'train_golden.pt')

    # ---- the survey's known answers on the full sample catalog (V=3429, d=8) ------------------
    kat = {}
    first32 = all_samples[:32]
    for name, K, fusion in (('SRGNN', 1, False), ('NISER', 1, False), ('MSGIFSR', 1, False), ('MSGIFSR', 2, False),
                            ('MSGIFSR', 3, False), ('MSGIFSR', 2, True)):
        c = model_case(R, name, first32, 3429, 8, 1, K=K, fusion=fusion)
        emb = 'embeddings.weight' if name == 'MSGIFSR' else 'embedding.weight'
        kat[f'{name}_K{K}_fusion{int(fusion)}'] = dict(
            loss=c['loss'], out0=c['out'][0, :3].tolist(), out_sum=float(c['out'].double().sum()),
            dE_norm=float(c['grads'][emb].norm()), top5=c['out'][0].topk(5)[1].tolist())
    (GOLD / 'known_answers_v3429_d8.json').write_text(json.dumps(kat, indent=1))
    print('done ->', GOLD)


if __name__ == '__main__':
    main()
