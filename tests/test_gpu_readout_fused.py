"""Fused per-session readout kernels (csrc/readout_fused.cu) against an fp64 restatement of AttnReadout + fc_sr +
normalise (srgnn.py:76-91,141-143; niser.py:147-148; msgifsr.py:124-155,269-273) and its autograd."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ops(pkg):
    from sessionrec_pytorch_b200 import ops as o
    return o


def _case(B, d, seed, maxlen=9):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, maxlen + 1, (B,), generator=g)
    lens[0] = 19                                        # one long session
    seg = torch.zeros(B + 1, dtype=torch.int64)
    seg[1:] = torch.cumsum(lens, 0)
    N = int(seg[-1])
    last = torch.stack([seg[b] + torch.randint(0, int(lens[b]), (1,), generator=g)[0] for b in range(B)])
    n2s = torch.repeat_interleave(torch.arange(B), lens)
    F = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    w = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) / d ** 0.5      # noqa: E731
    return dict(B=B, N=N, d=d, seg=seg, last=last, n2s=n2s, F=F, Wu=w(d, d), bu=w(d), Wv=w(d, d), we=w(d), Wsr=w(d, 2 * d))


def _chain(c, mode, grad=False):
    """fp64 forward; with grad=True the leaves F / we and the intermediates u, v, s keep their gradients."""
    F = c['F'].double().requires_grad_(grad)
    we = c['we'].double().requires_grad_(grad)
    u = F @ c['Wu'].double().t() + c['bu'].double()
    v = F[c['last']] @ c['Wv'].double().t()
    e = (torch.sigmoid(u + v[c['n2s']]) * we).sum(-1)
    gs, m, ssum = [], [], []
    for b in range(c['B']):
        lo, hi = int(c['seg'][b]), int(c['seg'][b + 1])
        mb = e[lo:hi].max()
        w = torch.exp(e[lo:hi] - mb)
        m.append(mb.detach())
        ssum.append(w.sum().detach())
        gs.append((w[:, None] * F[lo:hi]).sum(0) / w.sum())
    sr_in = torch.cat([F[c['last']], torch.stack(gs)], 1)
    s = sr_in @ c['Wsr'].double().t()
    n = s.norm(dim=-1)
    shat = s if mode == 0 else (s / n.clamp(min=1e-12)[:, None] if mode == 2 else s / (n + 1e-12)[:, None])
    if grad:
        for t in (u, v, s):
            t.retain_grad()
    return dict(F=F, we=we, u=u, v=v, e=e, m=torch.stack(m), ssum=torch.stack(ssum), sr_in=sr_in, s=s, shat=shat, n=n)


def _close(name, got, ref, tol):
    assert torch.isfinite(got).all(), name
    err = float((got.detach().cpu().double() - ref.detach()).abs().max())
    sc = max(float(ref.detach().abs().max()), 1e-12)
    assert err <= tol * sc, f'{name}: max|d| {err:.3e} vs max|ref| {sc:.3e} (rel {err / sc:.2e})'


def _device_inputs(ops, c):
    dev = lambda t, dt=None: (t if dt is None else t.to(dt)).to(DEV).contiguous()       # noqa: E731
    W = dev(c['Wsr'])
    WsrT = torch.empty(2 * c['d'], c['d'], device=DEV)
    ops.transpose(W, c['d'], 2 * c['d'], WsrT)
    torch.cuda.synchronize()
    assert torch.equal(WsrT, W.t().contiguous())
    return dict(F=dev(c['F']), we=dev(c['we']), Wsr=W, WsrT=WsrT, seg=dev(c['seg'], torch.int32), last=dev(c['last'], torch.int32))


CASES = [(512, 96, 2), (37, 64, 0), (300, 112, 3), (200, 128, 2), (5, 32, 2), (1000, 16, 2), (64, 256, 3)]


@pytest.mark.parametrize('B,d,mode', CASES)
def test_readout_tail_fwd(ops, B, d, mode):
    c = _case(B, d, seed=B + d)
    r = _chain(c, mode)
    x = _device_inputs(ops, c)
    N = c['N']
    u, v = r['u'].float().to(DEV).contiguous(), r['v'].float().to(DEV).contiguous()
    out = {k: torch.full(sh, 7.0, device=DEV) for k, sh in dict(e=(N,), ms=(B, 2), sr_in=(B, 2 * d), s=(B, d), shat=(B, d),
                                                                 rn=(B,)).items()}
    sbh = torch.zeros(B, d, dtype=torch.int16, device=DEV)
    sbl = torch.zeros(B, d, dtype=torch.int16, device=DEV)
    ops.readout_tail_fwd(x['F'], u, v, x['we'], x['WsrT'], x['seg'], x['last'], B, d, mode, out['e'], out['ms'], out['sr_in'], out['s'],
                         out['shat'], out['rn'], sbh, sbl)
    torch.cuda.synchronize()
    for k, ref in (('e', r['e']), ('sr_in', r['sr_in']), ('s', r['s']), ('shat', r['shat'])):
        _close(k, out[k], ref, 2e-5)
    _close('norm', out['rn'], r['n'], 2e-5)
    lse = out['ms'][:, 0].cpu().double() + torch.log(out['ms'][:, 1].cpu().double())
    _close('lse', lse, r['m'] + torch.log(r['ssum']), 2e-5)
    rec = sbh.view(torch.bfloat16).float() + sbl.view(torch.bfloat16).float()
    _close('bf16 hi + lo', rec, out['shat'].cpu().double(), 2.0 ** -15)


@pytest.mark.parametrize('B,d,mode', CASES)
def test_readout_head_bwd(ops, B, d, mode):
    """d shat -> (ds, du, dv, dwe, dF = alpha dg + dl) against fp64 autograd; the projection terms du Wu + dv Wv that the
    caller's GEMMs add are added in fp64 here."""
    c = _case(B, d, seed=7 * B + d)
    r = _chain(c, mode, grad=True)
    R = torch.randn(B, d, generator=torch.Generator().manual_seed(B)) / B
    (r['shat'] * R.double()).sum().backward()
    x = _device_inputs(ops, c)
    N = c['N']
    f32 = lambda t: t.detach().float().to(DEV).contiguous()       # noqa: E731
    ub, vb = f32(r['u']), f32(r['v'])
    ms = torch.stack([r['m'], r['ssum']], 1).float().to(DEV).contiguous()
    ds = torch.full((B, d), 7.0, device=DEV)
    dF = torch.full((N, d), float('nan'), device=DEV)
    dwe = torch.zeros(d, device=DEV)
    ops.readout_head_bwd(x['F'], x['we'], x['Wsr'], x['seg'], x['last'], B, d, mode, f32(r['s']), f32(r['shat']), f32(r['n']),
                         f32(r['sr_in']), f32(r['e']), ms, R.to(DEV), ub, vb, ds, dF, dwe)
    torch.cuda.synchronize()
    _close('ds', ds, r['s'].grad, 1e-4)
    _close('du', ub, r['u'].grad, 1e-4)
    _close('dv', vb, r['v'].grad, 1e-4)
    _close('dwe', dwe, r['we'].grad, 1e-4)
    full = dF.cpu().double() + r['u'].grad @ c['Wu'].double()
    full[c['last']] += r['v'].grad @ c['Wv'].double()
    _close('dF', full, r['F'].grad, 1e-4)
