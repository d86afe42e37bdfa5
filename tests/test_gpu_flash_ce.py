"""Fused scoring + cross-entropy head (csrc/flash_ce.cu, tcgen05 bf16 x 3) against an fp64 restatement of
`log(softmax(scale * s @ E.t()))` + `nll_loss` (srgnn.py:146-147, msgifsr.py:308-309, utils/train.py:99) and its
autograd.  Tolerance: the north-star bar, 1e-4 relative (logits live in [-scale, scale] for the normalised models)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
RTOL = 1e-4


@pytest.fixture(scope='module')
def ops(pkg):
    from sessionrec_pytorch_b200 import ops as o
    return o


def _r(*shape, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g).float()


def _split(ops, X):
    rows, cols = X.shape
    hi = torch.zeros(rows, cols, dtype=torch.int16, device=DEV)
    lo = torch.zeros(rows, cols, dtype=torch.int16, device=DEV)
    ops.split_bf16(X.to(DEV).contiguous(), cols, rows, cols, hi, lo, cols)
    torch.cuda.synchronize()
    rec = hi.view(torch.bfloat16).float() + lo.view(torch.bfloat16).float()
    rel = float(((rec.cpu() - X).abs() / X.abs().clamp(min=1e-30)).max())
    assert rel < 2.0 ** -15, f'bf16 hi + lo must reconstruct x to ~2^-17 (got {rel:.2e})'
    return hi, lo


def _ref(S, E, labels, scale, gout=1.0):
    S64 = S.double().requires_grad_(True)
    E64 = E.double().requires_grad_(True)
    Z = scale * (S64 @ E64.t())
    lse = torch.logsumexp(Z, -1)
    nll = lse - Z.gather(1, labels.unsqueeze(1)).squeeze(1)
    (gout * nll.mean()).backward()
    return lse.detach(), nll.detach(), S64.grad, E64.grad


def _errmap(name, got, ref, row_blk, col_blk=32):
    """Coarse map of max|got - ref| / max|ref| per (row block, column block): tells a layout bug from a rounding one."""
    g, r = got.cpu().double(), ref.double()
    sc = float(r.abs().max())
    rows = []
    for r0 in range(0, g.shape[0], row_blk):
        rows.append(' '.join(f'{float((g[r0:r0 + row_blk, c0:c0 + col_blk] - r[r0:r0 + row_blk, c0:c0 + col_blk]).abs().max()) / sc:8.1e}'
                             for c0 in range(0, g.shape[1], col_blk)))
    print(f'{name}: relative error map, {row_blk}-row x {col_blk}-col blocks (first 12 row blocks)\n  ' + '\n  '.join(rows[:12]))


CASES = [(128, 128, 64, 12.0, True), (16, 200, 32, 12.0, True), (100, 1000, 64, 12.0, True), (300, 5000, 128, 12.0, True),
         (512, 43097, 96, 12.0, True), (2048, 3000, 16, 1.0, False), (130, 129, 112, 1.0, False), (640, 17000, 128, 12.0, True),
         # d = 256: the wide kernels (catalog tile streamed in K chunks, 64-row backward tiles, dE produced transposed)
         (128, 128, 256, 12.0, True), (100, 1000, 256, 12.0, True), (130, 129, 256, 1.0, False), (300, 5000, 256, 12.0, True),
         (512, 3703, 256, 12.0, True), (2048, 17000, 256, 1.0, False)]


@pytest.mark.parametrize('B,V,d,scale,normed', CASES)
def test_flash_ce_fwd_bwd(ops, B, V, d, scale, normed):
    S, E = _r(B, d), _r(V, d, seed=1)
    if normed:
        S = torch.nn.functional.normalize(S, dim=-1)
        E = torch.nn.functional.normalize(E, dim=-1)
    else:                                   # SRGNN-like: un-normalised, small entries
        S, E = 0.3 * S, 0.3 * E
    labels = torch.randint(0, V, (B,), generator=torch.Generator().manual_seed(B + V))
    labels[0] = V - 1
    labels[-1] = 0
    Sh, Sl = _split(ops, S)
    Eh, El = _split(ops, E)
    lab = labels.int().to(DEV)
    lse = torch.full((B,), 7.0, device=DEV)
    nll = torch.full((B,), 7.0, device=DEV)
    part = torch.empty(ops.flash_ce_part_floats(B, V), device=DEV)
    ops.flash_ce_fwd(B, V, d, Sh, Sl, d, Eh, El, d, scale, lab, lse, nll, part)
    torch.cuda.synchronize()
    gout = 0.75
    rl, rn, rdS, rdE = _ref(S, E, labels, scale, gout)
    el = float((lse.cpu().double() - rl).abs().max())
    en = float((nll.cpu().double() - rn).abs().max())
    print(f'fwd B={B} V={V} d={d}: max|d lse| = {el:.2e}, max|d nll| = {en:.2e} (|lse| ~ {float(rl.abs().max()):.2f})')
    assert el <= RTOL * max(1.0, float(rl.abs().max()))
    assert en <= RTOL * max(1.0, float(rn.abs().max()))

    parts = ops.flash_ce_bwd_parts(B)
    assert parts == (B + 127) // 128
    dS = torch.full((B, d), 3.0, device=DEV)                       # overwritten
    dEp = torch.full((parts, V, d), float('nan'), device=DEV)      # every element must be written
    g = torch.tensor([gout], device=DEV)
    ops.flash_ce_bwd(B, V, d, Sh, Sl, d, Eh, El, d, scale, lab, lse, g, dS, dEp)
    dE = torch.zeros(V, d, device=DEV)
    ops.sum_parts(dEp, V * d, parts, V * d, dE, accumulate=False)
    torch.cuda.synchronize()
    assert torch.isfinite(dEp).all(), 'a dE partial was left unwritten'
    for name, got, ref in (('dS', dS, rdS), ('dE', dE, rdE)):
        err = float((got.cpu().double() - ref).abs().max())
        sc = float(ref.abs().max())
        print(f'bwd {name}: max|d| = {err:.3e}, max|ref| = {sc:.3e}, rel = {err / sc:.2e}')
        if not err <= RTOL * sc:
            _errmap(name, got, ref, 32 if name == 'dS' else max(32, (V + 11) // 12 // 32 * 32))
        assert err <= RTOL * sc, f'{name}: rel err {err / sc:.2e}'


def test_flash_ce_label_outside_shard(ops):
    """Catalog sharding: a label that lies outside [0, V) contributes nll = 0 and no onehot term."""
    B, V, d = 64, 500, 32
    S = torch.nn.functional.normalize(_r(B, d), dim=-1)
    E = torch.nn.functional.normalize(_r(V, d, seed=1), dim=-1)
    Sh, Sl = _split(ops, S)
    Eh, El = _split(ops, E)
    lab = torch.full((B,), -1, dtype=torch.int32, device=DEV)
    lse, nll = torch.empty(B, device=DEV), torch.full((B,), 5.0, device=DEV)
    part = torch.empty(ops.flash_ce_part_floats(B, V), device=DEV)
    ops.flash_ce_fwd(B, V, d, Sh, Sl, d, Eh, El, d, 12.0, lab, lse, nll, part)
    torch.cuda.synchronize()
    assert float(nll.abs().max()) == 0.0
    ref = torch.logsumexp(12.0 * (S.double() @ E.double().t()), -1)
    assert float((lse.cpu().double() - ref).abs().max()) <= RTOL * float(ref.abs().max())


def test_flash_ce_rejects_unsupported_dim(ops, pkg):
    B, V, d = 8, 64, 8
    z = torch.zeros(B, d, dtype=torch.int16, device=DEV)
    e = torch.zeros(V, d, dtype=torch.int16, device=DEV)
    with pytest.raises(pkg._lib.SessRecError):
        ops.flash_ce_fwd(B, V, d, z, z, d, e, e, d, 1.0, torch.zeros(B, dtype=torch.int32, device=DEV),
                         torch.empty(B, device=DEV), None, torch.empty(ops.flash_ce_part_floats(B, V), device=DEV))


@pytest.mark.parametrize('B,V,d,K', [(512, 43097, 96, 20), (300, 5000, 256, 20), (100, 1000, 64, 32), (16, 40, 32, 20), (130, 129, 128, 1)])
def test_flash_ce_topk_matches_fp64_ranking(ops, B, V, d, K):
    """Fused evaluation head: the K best items per row straight from the scoring kernel (no (B, V) matrix) against an fp64
    ranking of the same logits; positions where two consecutive reference scores are closer than the path's tolerance may swap."""
    S = torch.nn.functional.normalize(_r(B, d), dim=-1)
    E = torch.nn.functional.normalize(_r(V, d, seed=1), dim=-1)
    E[5] = E[3]                                             # an exact tie: the smaller id must come first
    Sh, Sl = _split(ops, S)
    Eh, El = _split(ops, E)
    idx = torch.full((B, K), -7, dtype=torch.int32, device=DEV)
    val = torch.empty(B, K, device=DEV)
    ops.flash_ce_topk(B, V, d, Sh, Sl, d, Eh, El, d, 12.0, K, idx, val)
    torch.cuda.synchronize()
    Z = 12.0 * (S.double() @ E.double().t())
    rv, ri = Z.topk(min(K + 1, V))
    got = idx.cpu().long()
    assert int(got.min()) >= 0 and int(got.max()) < V
    gv = Z.gather(1, got)                                   # reference scores of the returned ids
    assert float((val.cpu().double() - gv).abs().max()) <= 1e-4 * 12.0, 'returned values are not the logits of the returned ids'
    assert float((gv - rv[:, :K]).abs().max()) <= 2e-4 * 12.0, 'a returned item is not among the K best within tolerance'
    gaps = (rv[:, :-1] - rv[:, 1:]).min(-1)[0] if rv.shape[1] > 1 else torch.ones(B, dtype=torch.float64)
    safe = gaps > 1e-3
    assert int(safe.sum()) >= B // 4
    assert torch.equal(got[safe], ri[safe][:, :K]), 'ids differ on rows without near-ties'
    for b in range(B):
        assert len(set(got[b].tolist())) == K, 'duplicate ids in a row'
        row = got[b].tolist()
        if 3 in row and 5 in row:
            assert row.index(3) < row.index(5), 'exact tie: the smaller id comes first'
