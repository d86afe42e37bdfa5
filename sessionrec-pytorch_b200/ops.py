"""Thin tensor-level wrappers over the C ABI (include/sessrec_b200.h).  Every function only enqueues kernels on
the current CUDA stream.  No torch math lives here: torch is used for memory and streams only."""
import ctypes

import torch

from . import _lib
from ._lib import Dropout, GatInst, ptr

_count = [0]
# SESSREC_HOST_PROFILE=1: seconds spent inside the native one-call training steps (the C call alone), next to
# bench.py's host_enqueue_ms_per_step (the whole train_step call) - tells how much of the enqueue time is Python
import os as _os
import time as _time
HOST_PROFILE = _os.environ.get('SESSREC_HOST_PROFILE', '0') == '1'
native_call_s = [0.0, 0]


def timed_native(fn):
    if not HOST_PROFILE:
        return fn()
    t = _time.perf_counter()
    r = fn()
    native_call_s[0] += _time.perf_counter() - t
    native_call_s[1] += 1
    return r


def launches():
    return _count[0]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _call(name, *args):
    _count[0] += 1
    return _lib.lib().call(name, *args, _stream())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SessRecError('sessrec_b200 kernels need CUDA tensors; there is no CPU fallback')


def drop_cfg(p, site, seed):
    return Dropout(float(p), int(site), int(seed) & 0xFFFFFFFFFFFFFFFF)


def _dref(dc):
    return None if dc is None else ctypes.cast(ctypes.pointer(dc), ctypes.c_void_p)


# ---- GEMM forms -------------------------------------------------------------------------------------------

def gemm(M, N, K, A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, a_idx=None, b_idx=None, c_idx=None, bias=None, alpha=1.0,
         accumulate=False, split_k=0):
    _need_cuda(A, B, C)
    _call('srk_gemm', M, N, K, ptr(A), sa_m, sa_k, ptr(B), sb_k, sb_n, ptr(C), ldc, ptr(a_idx), ptr(b_idx), ptr(c_idx),
          ptr(bias), float(alpha), int(bool(accumulate)), int(split_k if accumulate else 1))


# Encoder GEMMs go to the tcgen05 tensor cores (srk_tc_gemm: 3xTF32, fp32-faithful) once they are big enough to pay for
# the operand-split launch; below that the fp32 CUDA-core kernel wins on latency.  SESSREC_TC_GEMM=0 disables the routing.
import os as _os

TC_GEMM = _os.environ.get('SESSREC_TC_GEMM', '1') != '0' and _os.environ.get('SESSREC_NO_UMMA', '0') != '1'
TC_MIN_MACS = int(_os.environ.get('SESSREC_TC_MIN_MACS', str(1 << 24)))


def _tc_ok(M, N, K, *mats):
    if not TC_GEMM or M * N * K < TC_MIN_MACS or K < 32 or N < 16:
        return False
    for t, ld in mats:
        if t.data_ptr() % 16 or ld % 4:
            return False
    return True


def tc_gemm(form, M, N, K, A, lda, Bm, ldb, C, ldc, bias=None, alpha=1.0, accumulate=False, split_k=0):
    """C (+)= alpha * op(A) op(B) (+ bias) on the tensor cores from plain fp32 operands (form 0: A[M,K] B[N,K]; 1: A[M,K] B[K,N];
    2: A[K,M] B[K,N])."""
    _need_cuda(A, Bm, C)
    n = int(_lib.lib().functions['srk_tc_gemm_scratch_floats'](form, M, N, K))
    scratch = torch.empty(n, dtype=torch.float32, device=C.device)
    _call('srk_tc_gemm', form, M, N, K, ptr(A), lda, ptr(Bm), ldb, ptr(C), ldc, ptr(bias), float(alpha),
          int(bool(accumulate)), int(split_k), ptr(scratch))


def linear_nt(X, W, C, M=None, K=None, lda=None, ldc=None, a_idx=None, c_idx=None, bias=None, alpha=1.0,
              accumulate=False):
    """C[M, N] (+)= alpha * X[M, K] @ W[N, K]^T (+ bias)."""
    N, Kw = W.shape
    K = Kw if K is None else K
    M = X.shape[0] if M is None else M
    lda_ = X.stride(0) if lda is None else lda
    if a_idx is None and c_idx is None and _tc_ok(M, N, K, (X, lda_), (W, W.stride(0))):
        return tc_gemm(0, M, N, K, X, lda_, W, W.stride(0), C, C.stride(0) if ldc is None else ldc, bias=bias, alpha=alpha,
                       accumulate=accumulate, split_k=1)
    gemm(M, N, K, X, X.stride(0) if lda is None else lda, 1, W, 1, W.stride(0), C, C.stride(0) if ldc is None else ldc,
         a_idx=a_idx, c_idx=c_idx, bias=bias, alpha=alpha, accumulate=accumulate)


def mm_nn(A, Bm, C, M=None, lda=None, ldc=None, c_idx=None, alpha=1.0, accumulate=False):
    """C[M, N] (+)= alpha * A[M, K] @ Bm[K, N]."""
    K, N = Bm.shape
    M = A.shape[0] if M is None else M
    lda_ = A.stride(0) if lda is None else lda
    if c_idx is None and _tc_ok(M, N, K, (A, lda_), (Bm, Bm.stride(0))):
        return tc_gemm(1, M, N, K, A, lda_, Bm, Bm.stride(0), C, C.stride(0) if ldc is None else ldc, alpha=alpha,
                       accumulate=accumulate, split_k=1)
    gemm(M, N, K, A, lda_, 1, Bm, Bm.stride(0), 1, C, C.stride(0) if ldc is None else ldc,
         c_idx=c_idx, alpha=alpha, accumulate=accumulate)


def mm_tn(A, Bm, C, K=None, lda=None, ldb=None, ldc=None, b_idx=None, alpha=1.0, accumulate=True, M=None, N=None):
    """C[M, N] (+)= alpha * A[K, M]^T @ Bm[K, N]  (weight-gradient form; split-K picked automatically)."""
    K = A.shape[0] if K is None else K
    M = A.shape[1] if M is None else M
    N = Bm.shape[1] if N is None else N
    lda_, ldb_ = A.stride(0) if lda is None else lda, Bm.stride(0) if ldb is None else ldb
    if b_idx is None and accumulate and _tc_ok(M, N, K, (A, lda_), (Bm, ldb_)):
        return tc_gemm(2, M, N, K, A, lda_, Bm, ldb_, C, C.stride(0) if ldc is None else ldc, alpha=alpha, accumulate=True,
                       split_k=0)
    gemm(M, N, K, A, 1, lda_, Bm, ldb_, 1, C,
         C.stride(0) if ldc is None else ldc, b_idx=b_idx, alpha=alpha, accumulate=accumulate, split_k=0)


def umma_gemm(form, M, N, K, Ahi, Alo, lda, Bhi, Blo, ldb, C, ldc, alpha=1.0, accumulate=False, split_k=1):
    _need_cuda(Ahi, Alo, Bhi, Blo, C)
    _call('srk_umma_gemm', form, M, N, K, ptr(Ahi), ptr(Alo), lda, ptr(Bhi), ptr(Blo), ldb, ptr(C), ldc, float(alpha),
          int(bool(accumulate)), int(split_k))


def comm_allreduce(buf, op=0):
    """In-place all-reduce on this library's own NCCL communicator (parallel.init_comm): op 0 sum, 1 max, 2 average."""
    _need_cuda(buf)
    _call('srk_comm_allreduce', ptr(buf), buf.numel(), int(op))


def umma_score_fwd(M, N, K, Ahi, Alo, lda, Bhi, Blo, ldb, Z, ldz, alpha, labels, lse, nll, part):
    _call('srk_umma_score_fwd', M, N, K, ptr(Ahi), ptr(Alo), lda, ptr(Bhi), ptr(Blo), ldb, ptr(Z), ldz, float(alpha),
          ptr(labels), ptr(lse), ptr(nll), ptr(part))


def split_tf32(X, ldx, rows, cols, hi, lo, ldo):
    _call('srk_split_tf32', ptr(X), ldx, rows, cols, ptr(hi), ptr(lo), ldo)


# ---- fused scoring + cross-entropy head (csrc/flash_ce.cu) ---------------------------------------------------

def flash_ce_supported(d):
    return bool(_lib.lib().functions['srk_flash_ce_supported'](int(d)))


def split_bf16(X, ldx, rows, cols, hi, lo, ldo):
    _call('srk_split_bf16', ptr(X), ldx, rows, cols, ptr(hi), ptr(lo), ldo)


def flash_ce_part_floats(B, V):
    return int(_lib.lib().functions['srk_flash_ce_part_floats'](B, V))


def flash_ce_bwd_parts(B):
    return int(_lib.lib().functions['srk_flash_ce_bwd_parts'](B))


def flash_ce_fwd(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, labels, lse, nll, part):
    _need_cuda(Shi, Slo, Ehi, Elo, labels, lse, part)
    _call('srk_flash_ce_fwd', B, V, d, ptr(Shi), ptr(Slo), lds, ptr(Ehi), ptr(Elo), lde, float(scale), ptr(labels), ptr(lse),
          ptr(nll), ptr(part))


def flash_ce_bwd(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, labels, lse, gout, dS, dEpart):
    _need_cuda(Shi, Slo, Ehi, Elo, labels, lse, dS, dEpart)
    _call('srk_flash_ce_bwd', B, V, d, ptr(Shi), ptr(Slo), lds, ptr(Ehi), ptr(Elo), lde, float(scale), ptr(labels), ptr(lse),
          ptr(gout), ptr(dS), ptr(dEpart))


def flash_ce_topk(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, k, out_idx, out_val=None):
    """ids [B, k] int32 (best first) of the k largest logits per row, fused into the scoring kernel (no (B, V) matrix)."""
    _need_cuda(Shi, Slo, Ehi, Elo, out_idx)
    n = int(_lib.lib().functions['srk_flash_ce_topk_scratch_floats'](B, V, k))
    scratch = torch.empty(n, dtype=torch.float32, device=out_idx.device)
    _call('srk_flash_ce_topk', B, V, d, ptr(Shi), ptr(Slo), lds, ptr(Ehi), ptr(Elo), lde, float(scale), int(k), ptr(out_idx),
          ptr(out_val), ptr(scratch))


def sum_parts(parts, stride, nparts, n, out, accumulate=False):
    _call('srk_sum_parts', ptr(parts), stride, nparts, n, ptr(out), int(bool(accumulate)))


# ---- embedding ---------------------------------------------------------------------------------------------

def embed_gather_fwd(E, iid, P, d, mode, dc, X, rnorm, x_first=None):
    _need_cuda(E, iid, X)
    _call('srk_embed_gather_fwd', ptr(E), ptr(iid), P, d, mode, _dref(dc), ptr(X), ptr(rnorm), ptr(x_first))


def embed_scatter_ws_floats(P, d):
    return int(_lib.lib().functions['srk_embed_scatter_ws_floats'](int(P), int(d)))


def embed_scatter_bwd(E, t, d, mode, dc, rnorm, dX, dX_first, dE, deterministic=True, ws=None):
    """dE[item] += gradient of every occurrence of the item.  deterministic: runs cut by a chunk boundary are combined in a
    fixed order through a small workspace (no atomics); False keeps the one-launch variant with atomicAdd on those rows."""
    if deterministic and ws is None:
        ws = torch.empty(max(embed_scatter_ws_floats(t['P'], d), 1), dtype=torch.float32, device=dE.device)
    if not deterministic:
        ws = None
    _call('srk_embed_scatter_bwd_ws', ptr(E), ptr(t['iid']), ptr(t['perm']), ptr(t['uoff']), ptr(t['uid']), t['U'], t['P'], d, mode,
          _dref(dc), ptr(rnorm), ptr(dX), ptr(dX_first), ptr(dE), ptr(ws))


def catalog_prep_fwd(E, mode, max_norm, Ehat, enorm, Ehi=None, Elo=None, Bhi=None, Blo=None):
    V, d = E.shape
    _call('srk_catalog_prep_fwd', ptr(E), V, d, mode, float(max_norm), ptr(Ehat), ptr(enorm), ptr(Ehi), ptr(Elo),
          ptr(Bhi), ptr(Blo))


def catalog_prep_bwd(E, Ehat, enorm, dEhat, mode, dE, nparts=1):
    """dE += rownorm-backward of dEhat; dEhat may be `nparts` stacked [V, d] partial sums (flash CE backward)."""
    V, d = E.shape
    _call('srk_catalog_prep_bwd', ptr(E), ptr(Ehat), ptr(enorm), ptr(dEhat), int(nparts), V, d, mode, ptr(dE))


def renorm_rows(E, uid, U, max_norm=1.0):
    _call('srk_renorm_rows', ptr(E), ptr(uid), U, E.shape[1], float(max_norm))


def rownorm_fwd(X, ldx, R, d, mode, Y, ldy, rnorm):
    _call('srk_rownorm_fwd', ptr(X), ldx, R, d, mode, ptr(Y), ldy, ptr(rnorm))


def rownorm_bwd(X, ldx, Y, ldy, rnorm, dY, lddy, R, d, mode, dX, lddx, accumulate=False):
    _call('srk_rownorm_bwd', ptr(X), ldx, ptr(Y), ldy, ptr(rnorm), ptr(dY), lddy, R, d, mode, ptr(dX), lddx,
          int(bool(accumulate)))


def expander_combine_fwd(X, h, N, k, d, out, rnorm):
    _call('srk_expander_combine_fwd', ptr(X), ptr(h), N, k, d, ptr(out), ptr(rnorm))


def expander_combine_bwd(out, rnorm, dout, N, k, d, dh, dX):
    _call('srk_expander_combine_bwd', ptr(out), ptr(rnorm), ptr(dout), N, k, d, ptr(dh), ptr(dX))


# ---- elementwise -------------------------------------------------------------------------------------------

def dropout_apply(X, Y, n, dc, accumulate=False):
    _call('srk_dropout_apply', ptr(X), ptr(Y), n, _dref(dc), int(bool(accumulate)))


def fill(X, value=0.0):
    _call('srk_fill', ptr(X), X.numel(), float(value))


def gather_rows(X, idx, R, d, Y, ldy):
    _call('srk_gather_rows', ptr(X), ptr(idx), R, d, ptr(Y), ldy)


def scatter_add_rows(X, ldx, idx, R, d, Y):
    _call('srk_scatter_add_rows', ptr(X), ldx, ptr(idx), R, d, ptr(Y))


def colsum(X, ldx, R, d, out, accumulate=True):
    _call('srk_colsum', ptr(X), ldx, R, d, ptr(out), int(bool(accumulate)))


def mean(x, n, out):
    _call('srk_mean', ptr(x), n, ptr(out))


# ---- readout / CE ------------------------------------------------------------------------------------------

def readout_fwd(F, u, v, we, seg, last, B, d, with_last, e, ms, sr_in):
    _call('srk_readout_fwd', ptr(F), ptr(u), ptr(v), ptr(we), ptr(seg), ptr(last), B, d, int(with_last), ptr(e), ptr(ms),
          ptr(sr_in))


def transpose(X, rows, cols, Y):
    _call('srk_transpose', ptr(X), rows, cols, ptr(Y))


def readout_tail_fwd(F, u, v, we, WsrT, seg, last, B, d, norm_mode, e, ms, sr_in, s, shat, rn_s, sbh=None, sbl=None):
    """WsrT = Wsr^T ([2d, d], ops.transpose)."""
    _need_cuda(F, u, v, we, WsrT, e, ms, sr_in, s, shat)
    _call('srk_readout_tail_fwd', ptr(F), ptr(u), ptr(v), ptr(we), ptr(WsrT), ptr(seg), ptr(last), B, d, int(norm_mode), ptr(e),
          ptr(ms), ptr(sr_in), ptr(s), ptr(shat), ptr(rn_s), ptr(sbh), ptr(sbl))


def readout_head_bwd(F, we, Wsr, seg, last, B, d, norm_mode, s, shat, rn_s, sr_in, e, ms, dshat, u, v, ds, dF, dwe):
    _need_cuda(F, we, Wsr, s, shat, sr_in, e, ms, dshat, u, v, ds, dF, dwe)
    _call('srk_readout_head_bwd', ptr(F), ptr(we), ptr(Wsr), ptr(seg), ptr(last), B, d, int(norm_mode), ptr(s), ptr(shat),
          ptr(rn_s), ptr(sr_in), ptr(e), ptr(ms), ptr(dshat), ptr(u), ptr(v), ptr(ds), ptr(dF), ptr(dwe))


def readout_bwd(F, u, v, we, seg, last, e, ms, sr_in, dsr_in, B, d, with_last, dF, dwe):
    _call('srk_readout_bwd', ptr(F), ptr(u), ptr(v), ptr(we), ptr(seg), ptr(last), ptr(e), ptr(ms), ptr(sr_in),
          ptr(dsr_in), B, d, int(with_last), ptr(dF), ptr(dwe))


def ce_rows_fwd(Z, ldz, labels, B, V, write_logp, lse, nll):
    _call('srk_ce_rows_fwd', ptr(Z), ldz, ptr(labels), B, V, int(write_logp), ptr(lse), ptr(nll))


def ce_rows_bwd(Z, ldz, labels, lse, gscale, scale, B, V, z_is_logp, Zlo=None):
    _call('srk_ce_rows_bwd', ptr(Z), ldz, ptr(labels), ptr(lse), ptr(gscale), float(scale), B, V, int(z_is_logp), ptr(Zlo))


def ce_rows_bwd_cols(Z, ldz, labels, lse, gscale, scale, B, col0, ncols, Zlo=None):
    _call('srk_ce_rows_bwd_cols', ptr(Z), ldz, ptr(labels), ptr(lse), ptr(gscale), float(scale), B, col0, ncols, ptr(Zlo))


def logp_bwd(LP, ldlp, G, ldg, scale, B, V, DZ, lddz, DZlo=None):
    _call('srk_logp_bwd', ptr(LP), ldlp, ptr(G), ldg, float(scale), B, V, ptr(DZ), lddz, ptr(DZlo))


def renorm_head_fwd(Z, ldz, B, V, iid, seg, lphi, zin):
    """REnorm head (msgifsr.py:281-305): scaled logits Z -> log(phi_0 softmax_in + phi_1 softmax_ex), in place."""
    _need_cuda(Z, iid, seg, lphi, zin)
    _call('srk_renorm_head_fwd', ptr(Z), ldz, B, V, ptr(iid), ptr(seg), ptr(lphi), ptr(zin))


def renorm_head_bwd(LP, ldlp, G, ldg, labels, gscale, scale, dl_scale, B, V, iid, seg, lphi, tmp, DZ, lddz, DZlo, dlphi):
    _need_cuda(LP, iid, seg, lphi, tmp, DZ, dlphi)
    _call('srk_renorm_head_bwd', ptr(LP), ldlp, ptr(G), ldg, ptr(labels), ptr(gscale), float(scale), float(dl_scale), B, V,
          ptr(iid), ptr(seg), ptr(lphi), ptr(tmp), ptr(DZ), lddz, ptr(DZlo), ptr(dlphi))


def gate_fwd(H, W2, B, d, lphi):
    _call('srk_gate_fwd', ptr(H), ptr(W2), B, d, ptr(lphi))


def gate_bwd(Hr, W2, lphi, dlphi, B, d, da, dH):
    _call('srk_gate_bwd', ptr(Hr), ptr(W2), ptr(lphi), ptr(dlphi), B, d, ptr(da), ptr(dH))


def topk_rows(Z, ldz, B, V, k, out_idx, out_val=None):
    _call('srk_topk_rows', ptr(Z), ldz, B, V, k, ptr(out_idx), ptr(out_val))


def mix_logp_fwd(Zall, head_stride, ldz, lse, alpha, K, B, V, out, ldo):
    _call('srk_mix_logp_fwd', ptr(Zall), head_stride, ldz, ptr(lse), ptr(alpha), K, B, V, ptr(out), ldo)


def mix_loss_fwd(nll, alpha, K, B, out):
    _call('srk_mix_loss_fwd', ptr(nll), ptr(alpha), K, B, ptr(out))


def mix_bwd(Zall, Zlo, head_stride, ldz, lse, alpha, K, B, V, G, ldg, labels, gscale, scale, rsum):
    _call('srk_mix_bwd', ptr(Zall), ptr(Zlo), head_stride, ldz, ptr(lse), ptr(alpha), K, B, V, ptr(G), ldg, ptr(labels),
          ptr(gscale), float(scale), ptr(rsum))


def mix_alpha_bwd(rsum, alpha, K, B, dalpha):
    _call('srk_mix_alpha_bwd', ptr(rsum), ptr(alpha), K, B, ptr(dalpha))


# ---- GGNN ----------------------------------------------------------------------------------------------------

def ggnn_aggregate_fwd(X, N, d, rel, NN, wsum):
    _call('srk_ggnn_aggregate_fwd', ptr(X), N, d, ptr(rel['in_ptr']), ptr(rel['in_src']), ptr(rel['in_eid']),
          ptr(rel['out_ptr']), ptr(rel['out_dst']), ptr(rel['out_eid']), ptr(rel['w']), ptr(NN), ptr(wsum))


def ggnn_aggregate_bwd(dNN, N, d, rel, wsum, dX, accumulate):
    _call('srk_ggnn_aggregate_bwd', ptr(dNN), N, d, ptr(rel['in_ptr']), ptr(rel['in_src']), ptr(rel['in_eid']),
          ptr(rel['out_ptr']), ptr(rel['out_dst']), ptr(rel['out_eid']), ptr(rel['w']), ptr(wsum), ptr(dX),
          int(bool(accumulate)))


def gru_pointwise_fwd(gi, gh, h, N, d, hnew):
    _call('srk_gru_pointwise_fwd', ptr(gi), ptr(gh), ptr(h), N, d, ptr(hnew))


def gru_pointwise_bwd(gi, gh, h, dhnew, N, d, dh, accumulate):
    _call('srk_gru_pointwise_bwd', ptr(gi), ptr(gh), ptr(h), ptr(dhnew), N, d, ptr(dh), int(bool(accumulate)))


# ---- GAT -----------------------------------------------------------------------------------------------------

def gat_prep(W, al, ar, d, Waug, wr):
    _call('srk_gat_prep', ptr(W), ptr(al), ptr(ar), d, ptr(Waug), ptr(wr))


def gat_prep_bwd(W, al, ar, dWaug, dwr, d, dW, dal, dar):
    _call('srk_gat_prep_bwd', ptr(W), ptr(al), ptr(ar), ptr(dWaug), ptr(dwr), d, ptr(dW), ptr(dal), ptr(dar))


def gat_inst_array(insts):
    arr = (GatInst * max(len(insts), 1))()
    for i, g in enumerate(insts):
        arr[i] = g
    return arr


def gat_aggregate_fwd(arr, n_inst, N, d, segmean, node2seg, dc, normalize, H, rnorm, amax):
    _call('srk_gat_aggregate_fwd', ctypes.cast(arr, ctypes.c_void_p), n_inst, N, d, ptr(segmean), ptr(node2seg),
          _dref(dc), int(normalize), ptr(H), ptr(rnorm), ptr(amax))


def gat_aggregate_bwd_dst(arr, n_inst, N, d, dc, normalize, H, rnorm, amax, dH, dHpre):
    _call('srk_gat_aggregate_bwd_dst', ctypes.cast(arr, ctypes.c_void_p), n_inst, N, d, _dref(dc), int(normalize), ptr(H),
          ptr(rnorm), ptr(amax), ptr(dH), ptr(dHpre))


def gat_aggregate_bwd_src(inst, d, dc, dHpre, amax):
    _call('srk_gat_aggregate_bwd_src', ctypes.cast(ctypes.pointer(inst), ctypes.c_void_p), d, _dref(dc), ptr(dHpre),
          ptr(amax))


def gat_bias_bwd(dHpre, amax, N, d, dbias):
    _call('srk_gat_bias_bwd', ptr(dHpre), ptr(amax), N, d, ptr(dbias))


def segmean_fwd(X, seg, B, d, out):
    _call('srk_segmean_fwd', ptr(X), ptr(seg), B, d, ptr(out))


def segmean_bwd(dHpre, seg, B, d, dX, accumulate):
    _call('srk_segmean_bwd', ptr(dHpre), ptr(seg), B, d, ptr(dX), int(bool(accumulate)))


# ---- optimizer -------------------------------------------------------------------------------------------------

def adam_step(param, grad, m, v, seg_off, seg_decay, n_seg, lr, b1, b2, eps, step, grad_scale=1.0):
    _call('srk_adam_step', ptr(param), ptr(grad), ptr(m), ptr(v), param.numel(), ptr(seg_off), ptr(seg_decay), n_seg,
          float(lr), float(b1), float(b2), float(eps), int(step), float(grad_scale))


def kernel_launches():
    """Kernels launched by libsessrec_b200.so so far (counted inside the library)."""
    return int(_lib.lib().functions['srk_launch_count']())
