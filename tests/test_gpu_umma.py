"""tcgen05 3xTF32 catalog GEMM (csrc/umma_gemm.cu) against fp64 on the CPU, all three operand forms, ragged sizes."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ops(pkg):
    from sessionrec_pytorch_b200 import ops as o
    return o


def _r(*shape, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g).float()


def _split(ops, X, ld):
    rows, cols = X.shape
    hi = torch.zeros(rows, ld, device=DEV)
    lo = torch.zeros(rows, ld, device=DEV)
    Xd = torch.zeros(rows, ld, device=DEV)
    Xd[:, :cols] = X.to(DEV)
    ops.split_tf32(Xd, ld, rows, cols, hi, lo, ld)
    torch.cuda.synchronize()
    assert torch.equal((hi + lo)[:, :cols].cpu(), X), 'hi + lo must reconstruct x exactly'
    return hi, lo


def _check(name, got, ref):
    err = float((got.double().cpu() - ref).abs().max())
    scale = float(ref.abs().max())
    print(f'{name}: max|d| = {err:.3e}, max|ref| = {scale:.3e}, rel = {err / scale:.2e}')
    assert err <= 1e-5 * scale, f'{name}: rel err {err / scale:.2e}'


@pytest.mark.parametrize('M,N,K', [(128, 128, 32), (128, 128, 96), (512, 1001, 96), (200, 300, 40), (64, 2000, 256), (512, 96, 64)])
def test_umma_nt(ops, M, N, K):
    A, B = _r(M, K), _r(N, K, seed=1)
    ldk = (K + 3) // 4 * 4
    Ah, Al = _split(ops, A, ldk)
    Bh, Bl = _split(ops, B, ldk)
    ldc = (N + 3) // 4 * 4
    C = torch.full((M, ldc), 7.0, device=DEV)
    ops.umma_gemm(0, M, N, K, Ah, Al, ldk, Bh, Bl, ldk, C, ldc, alpha=12.0)
    torch.cuda.synchronize()
    _check(f'nt {M}x{N}x{K}', C[:, :N], 12.0 * (A.double() @ B.double().t()))
    assert bool((C[:, N:] == 7.0).all()), 'padding columns must stay untouched'


@pytest.mark.parametrize('M,N,K,S', [(128, 96, 64, 1), (512, 96, 1001, 5), (200, 64, 4099, 7), (512, 256, 700, 3)])
def test_umma_nn(ops, M, N, K, S):
    """dS = dZ @ E: A[M, K] K-major, B[K, N] MN-major, split-K with atomic accumulation."""
    A, B = _r(M, K), _r(K, N, seed=2)
    lda = (K + 3) // 4 * 4
    Ah, Al = _split(ops, A, lda)
    Bh, Bl = _split(ops, B, N)
    C = torch.ones(M, N, device=DEV)
    ops.umma_gemm(1, M, N, K, Ah, Al, lda, Bh, Bl, N, C, N, alpha=0.5, accumulate=True, split_k=S)
    torch.cuda.synchronize()
    _check(f'nn {M}x{N}x{K}', C, 1.0 + 0.5 * (A.double() @ B.double()))


@pytest.mark.parametrize('M,N,K', [(128, 96, 32), (1001, 96, 512), (4099, 64, 200), (300, 256, 70)])
def test_umma_tn(ops, M, N, K):
    """dE = dZ^T @ s: A[K, M] and B[K, N] both MN-major."""
    A, B = _r(K, M), _r(K, N, seed=3)
    lda = (M + 3) // 4 * 4
    Ah, Al = _split(ops, A, lda)
    Bh, Bl = _split(ops, B, N)
    C = torch.zeros(M, N, device=DEV)
    ops.umma_gemm(2, M, N, K, Ah, Al, lda, Bh, Bl, N, C, N)
    torch.cuda.synchronize()
    _check(f'tn {M}x{N}x{K}', C, A.double().t() @ B.double())


@pytest.mark.parametrize('M,N,K', [(128, 256, 32), (512, 4099, 96), (200, 1001, 40), (512, 43097, 96), (2048, 3000, 256)])
def test_umma_score_fwd_fused_lse(ops, M, N, K):
    """Persistent forward kernel: Z, row log-sum-exp and label NLL in one pass."""
    A = torch.nn.functional.normalize(_r(M, K), dim=-1)
    B = torch.nn.functional.normalize(_r(N, K, seed=1), dim=-1)
    labels = torch.randint(0, N, (M,), generator=torch.Generator().manual_seed(M + N))
    ldk = (K + 3) // 4 * 4
    Ah, Al = _split(ops, A, ldk)
    Bh, Bl = _split(ops, B, ldk)
    ldz = (N + 3) // 4 * 4
    Z = torch.full((M, ldz), 7.0, device=DEV)
    lse, nll = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    part = torch.empty(4 * ((N + 255) // 256) * M + M, device=DEV)
    ops.umma_score_fwd(M, N, K, Ah, Al, ldk, Bh, Bl, ldk, Z, ldz, 12.0, labels.int().to(DEV), lse, nll, part)
    torch.cuda.synchronize()
    ref = 12.0 * (A.double() @ B.double().t())
    _check(f'score Z {M}x{N}x{K}', Z[:, :N], ref)
    rl = torch.logsumexp(ref, -1)
    assert float((lse.cpu().double() - rl).abs().max()) < 2e-5, float((lse.cpu().double() - rl).abs().max())
    rn = rl - ref.gather(1, labels.unsqueeze(1)).squeeze(1)
    assert float((nll.cpu().double() - rn).abs().max()) < 3e-5
    assert bool((Z[:, N:] == 7.0).all())


@pytest.mark.parametrize('form,M,N,K,bias,acc', [
    (0, 9000, 768, 512, True, False),      # GGNN gate projection at the cfg2 shape: gi = hn W_ih^T + b_ih
    (0, 333, 100, 72, True, False),        # ragged everything
    (0, 2048, 256, 512, False, True),      # accumulate into an existing C
    (1, 9000, 256, 256, False, True),      # dF += du W_u
    (1, 2048, 512, 256, False, False),     # d sr_in = ds W_sr: 512 output columns = two N tiles
    (1, 500, 40, 300, False, False),
    (2, 256, 256, 9000, False, True),      # dW_u += du^T F, split-K picked by the library
    (2, 256, 512, 2048, False, True),      # dW_sr += ds^T sr_in: two N tiles
    (2, 100, 36, 777, False, True),
])
def test_tc_gemm_plain_fp32_operands(ops, form, M, N, K, bias, acc):
    """srk_tc_gemm: operand split + 3xTF32 tensor-core GEMM from plain (pitched) fp32 operands, optional bias, any N."""
    g = torch.Generator().manual_seed(form * 7 + M + N + K)
    sa = (K, M) if form == 2 else (M, K)
    sb = (N, K) if form == 0 else (K, N)
    pa, pb = (sa[1] + 3) // 4 * 4 + 4, (sb[1] + 3) // 4 * 4 + 8          # pitched: views into wider buffers
    Af, Bf = torch.randn(sa[0], pa, generator=g), torch.randn(sb[0], pb, generator=g)
    A, Bm = Af[:, :sa[1]], Bf[:, :sb[1]]
    bv = torch.randn(N, generator=g) if bias else None
    C0 = torch.randn(M, N, generator=g)
    Cd = C0.to(DEV).contiguous()
    Ad, Bd = Af.to(DEV), Bf.to(DEV)
    ops.tc_gemm(form, M, N, K, Ad, pa, Bd, pb, Cd, N, bias=None if bv is None else bv.to(DEV), alpha=0.75, accumulate=acc, split_k=0)
    torch.cuda.synchronize()
    a, b = A.double(), Bm.double()
    ref = 0.75 * ((a.t() if form == 2 else a) @ (b.t() if form == 0 else b))
    if bv is not None:
        ref = ref + bv.double()
    if acc:
        ref = ref + C0.double()
    _check(f'tc_gemm form {form} {M}x{N}x{K}', Cd, ref)
