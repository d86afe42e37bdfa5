"""REnorm head kernels (`--extra`, msgifsr.py:281-305: csrc/readout_ce.cu renorm_head_* / gate_*) against an fp64
restatement with the reference's own masked soft-maxes, autograd for the backward.  Tolerance 1e-4 relative (north star);
the session's own items come as ragged ascending id lists, including a 1-item session, a 70-item one (more than two
warps' worth) and one that owns the last catalog column."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
RTOL = 1e-4


@pytest.fixture(scope='module')
def ops(pkg):
    from sessionrec_pytorch_b200 import ops as o
    return o


def _sessions(B, V, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, 12, (B,), generator=g).tolist()
    lens[0] = 1
    if B > 2:
        lens[2] = min(70, V - 1)
    items = [torch.randperm(V, generator=g)[:n].sort()[0] for n in lens]
    items[-1][-1] = V - 1
    items[-1] = items[-1].unique()
    seg = torch.zeros(B + 1, dtype=torch.int32)
    seg[1:] = torch.tensor([len(i) for i in items]).cumsum(0)
    return items, torch.cat(items).int(), seg


def _ref_logp(Z, lphi, items):
    B, V = Z.shape
    mask = torch.zeros(B, V, dtype=torch.bool)
    for b, it in enumerate(items):
        mask[b, it] = True
    s_in = torch.softmax(Z.masked_fill(~mask, float('-inf')), -1)
    s_ex = torch.softmax(Z.masked_fill(mask, float('-inf')), -1)
    return torch.log(lphi[:, 0:1].exp() * s_in + lphi[:, 1:2].exp() * s_ex), mask


@pytest.mark.parametrize('B,V,split', [(5, 97, False), (64, 5000, True), (33, 1030, False), (130, 43097, True)])
@pytest.mark.parametrize('mode', ['loss', 'grad'])
def test_renorm_head_fwd_bwd(ops, B, V, split, mode):
    g = torch.Generator().manual_seed(B + V)
    items, iid, seg = _sessions(B, V, B * 7 + V)
    ldz = (V + 3) // 4 * 4
    Z0 = 12.0 * (2 * torch.rand(B, V, generator=g) - 1)
    lphi0 = torch.log_softmax(torch.randn(B, 2, generator=g), -1)
    labels = torch.randint(0, V, (B,), generator=g)
    labels[0] = int(items[0][0])                          # a label inside the session's own set
    labels[-1] = V - 1
    Zr = Z0.double().requires_grad_(True)
    lr_ = lphi0.double().requires_grad_(True)
    ref, _ = _ref_logp(Zr, lr_, items)

    Z = torch.zeros(B, ldz, device=DEV)
    Z[:, :V] = Z0.to(DEV)
    lphi = lphi0.to(DEV).contiguous()
    iid_d, seg_d = iid.to(DEV), seg.to(DEV)
    zin = torch.empty(iid.numel(), device=DEV)
    ops.renorm_head_fwd(Z, ldz, B, V, iid_d, seg_d, lphi, zin)
    torch.cuda.synchronize()
    got = Z[:, :V].cpu().double()
    assert torch.isfinite(got).all()
    err = float((got - ref.detach()).abs().max())
    assert err <= RTOL * float(ref.detach().abs().max()), f'log score: {err:.2e}'
    assert float((got.exp().sum(-1) - 1).abs().max()) < 1e-4           # phi_0 + phi_1 = 1

    scale, gs = 12.0, 0.75
    tmp = torch.empty(2 * iid.numel(), device=DEV)
    dlphi = torch.empty(B, 2, device=DEV)
    Zlo = torch.zeros_like(Z) if split else None
    if mode == 'loss':
        (gs * torch.nn.functional.nll_loss(ref, labels)).backward()
        ops.renorm_head_bwd(Z, ldz, None, 0, labels.int().to(DEV), torch.tensor([gs], device=DEV), scale, 1.0, B, V, iid_d, seg_d,
                            lphi, tmp, Z, ldz, Zlo, dlphi)                       # in place over the log-probs
        dZ = Z
    else:
        G0 = torch.randn(B, V, generator=g)
        (ref * G0.double()).sum().backward()
        G = torch.zeros(B, ldz, device=DEV)
        G[:, :V] = G0.to(DEV)
        ops.renorm_head_bwd(Z, ldz, G, ldz, None, None, scale, 1.0, B, V, iid_d, seg_d, lphi, tmp, G, ldz, Zlo, dlphi)   # over G
        dZ = G
    torch.cuda.synchronize()
    got_dz = dZ[:, :V].cpu().double() + (Zlo[:, :V].cpu().double() if split else 0)
    ref_dz = scale * Zr.grad
    err = float((got_dz - ref_dz).abs().max())
    assert err <= RTOL * float(ref_dz.abs().max()), f'dZ: {err:.2e} of {float(ref_dz.abs().max()):.2e}'
    if split:                                            # hi part is a TF32 number: low 13 mantissa bits clear
        assert int((dZ[:, :V].contiguous().view(torch.int32) & 0x1FFF).abs().max()) == 0
    err = float((dlphi.cpu().double() - lr_.grad).abs().max())
    assert err <= RTOL * max(float(lr_.grad.abs().max()), 1e-6), f'dlphi: {err:.2e}'


@pytest.mark.parametrize('B,d', [(1, 16), (77, 96), (512, 256)])
def test_gate_fwd_bwd(ops, B, d):
    g = torch.Generator().manual_seed(B + d)
    H0 = torch.randn(B, d, generator=g)
    W2 = torch.randn(2, d, generator=g) / d ** 0.5
    Hr = H0.double().requires_grad_(True)
    Wr = W2.double().requires_grad_(True)
    ref = torch.log_softmax(torch.relu(Hr) @ Wr.t(), -1)
    dl = torch.randn(B, 2, generator=g)
    (ref * dl.double()).sum().backward()
    H, lphi = H0.to(DEV).contiguous(), torch.empty(B, 2, device=DEV)
    ops.gate_fwd(H, W2.to(DEV), B, d, lphi)
    da, dH = torch.empty(B, 2, device=DEV), torch.empty(B, d, device=DEV)
    ops.gate_bwd(H, W2.to(DEV), lphi, dl.to(DEV), B, d, da, dH)
    torch.cuda.synchronize()
    assert torch.equal(H.cpu(), torch.relu(H0))
    assert float((lphi.cpu().double() - ref.detach()).abs().max()) <= RTOL * float(ref.detach().abs().max())
    assert float((dH.cpu().double() - Hr.grad).abs().max()) <= RTOL * float(Hr.grad.abs().max())
    dW = da.cpu().double().t() @ torch.relu(H0).double()                # what mm_tn(da, relu(h)) accumulates
    assert float((dW - Wr.grad).abs().max()) <= RTOL * float(Wr.grad.abs().max())
