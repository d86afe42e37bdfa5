"""Shared machinery of the drop-in modules: flat parameters, the autograd bridge (one Function per model call;
the whole forward/backward below it is our kernels), the catalog scoring + cross-entropy head, fused training
step with the reference's Adam/L2 semantics."""
import os

import torch
from torch import nn

from . import _lib, ops
from .flat import FlatParams


class _Bridge(torch.autograd.Function):
    """forward(model, batch, mode, *params): params are passed only so that autograd routes gradients to them."""

    @staticmethod
    def forward(ctx, model, batch, mode, *params):
        out, tape = model._fwd(batch, mode, need_grad=any(ctx.needs_input_grad))
        ctx.model, ctx.tape, ctx.mode = model, tape, mode
        return out

    @staticmethod
    def backward(ctx, gout):
        model, tape = ctx.model, ctx.tape
        gflat = torch.zeros_like(model._flat.data)
        model._bwd(tape, gout.contiguous(), gflat)
        ctx.tape = None
        return (None, None, None, *model._flat.views(gflat))


class SessRecModule(nn.Module):
    """Base of SRGNN / NISER / MSGIFSR.  Sub-classes implement _encode_fwd / _encode_bwd (everything up to the
    session representation) and declare the catalog head via _head_cfg()."""

    def __init__(self):
        super().__init__()
        self._flat = None
        self._seed = 0x5EED
        self._step = 0
        self._fixed_seed = None
        self._opt = None
        self.use_tensor_cores = os.environ.get('SESSREC_NO_UMMA', '0') != '1'

    # ---- parameters -------------------------------------------------------------------------------------
    def _ensure_flat(self):
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise _lib.SessRecError(f'{type(self).__name__}: parameters live on {p0.device}; this path only runs on '
                                    'CUDA (sm_100a) and has no CPU fallback - call model.to("cuda") first')
        if self._flat is None or not self._flat.valid():
            self._flat = FlatParams(self)
            self._opt = None
        return self._flat

    def set_dropout_seed(self, seed):
        """Pin the counter-based dropout seed (tests inject the same masks into the oracle)."""
        self._fixed_seed = seed

    def _next_seed(self):
        if self._fixed_seed is not None:
            return self._fixed_seed
        self._step += 1
        return (self._seed * 0x9E3779B97F4A7C15 + self._step * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF

    def _p(self):
        return self.dropout_p if self.training else 0.0

    # ---- public API --------------------------------------------------------------------------------------
    def forward(self, mg, sg=None):
        """(B, V) fp32 log-probabilities with autograd - the reference contract (`utils/train.py:97-99`)."""
        self._ensure_flat()
        return _Bridge.apply(self, mg, 'logp', *self._flat.params)

    def loss(self, mg, labels=None):
        """Fused path: mean NLL of the batch's own labels (never rewrites Z into log-probs)."""
        self._ensure_flat()
        return _Bridge.apply(self, mg, 'loss', *self._flat.params)

    # ---- catalog scoring + CE head -----------------------------------------------------------------------
    def _head_fwd(self, shat, ld_s, Ehat, scale, batch, mode, tape, Ehi=None, Elo=None):
        """Z = scale * shat Ehat^T on the tcgen05 tensor cores (3xTF32, csrc/umma_gemm.cu) whenever the embedding
        dim fits one UMMA N tile (d <= 256); otherwise on the fp32 CUDA-core GEMM."""
        B, (V, d) = batch.B, Ehat.shape
        dev = Ehat.device
        ldz = (V + 3) // 4 * 4                      # 16-byte aligned rows: TMA / vector loads in the backward GEMMs
        Z = torch.empty(B, ldz, dtype=torch.float32, device=dev)
        umma = self.use_tensor_cores and d <= 256
        if umma:
            if Ehi is None:
                Ehi, Elo = torch.empty_like(Ehat), torch.empty_like(Ehat)
                ops.split_tf32(Ehat, d, V, d, Ehi, Elo, d)
            sh = torch.empty(B, d, dtype=torch.float32, device=dev)
            sl = torch.empty(B, d, dtype=torch.float32, device=dev)
            ops.split_tf32(shat, ld_s, B, d, sh, sl, d)
            ops.umma_gemm(0, B, V, d, sh, sl, d, Ehi, Elo, d, Z, ldz, alpha=scale)
            tape.update(Ehi=Ehi, Elo=Elo, sh=sh, sl=sl)
        else:
            ops.gemm(B, V, d, shat, ld_s, 1, Ehat, 1, d, Z, ldz, alpha=scale)
        lse = torch.empty(B, dtype=torch.float32, device=dev)
        tape.update(Z=Z, ldz=ldz, lse=lse, scale=scale, Ehat=Ehat, shat=shat, ld_s=ld_s, umma=umma)
        if mode == 'loss':
            nll = torch.empty(B, dtype=torch.float32, device=dev)
            ops.ce_rows_fwd(Z, ldz, batch.labels, B, V, False, lse, nll)
            out = torch.empty((), dtype=torch.float32, device=dev)
            ops.mean(nll, B, out)
            return out
        ops.ce_rows_fwd(Z, ldz, None, B, V, True, lse, None)
        return Z[:, :V]

    def _head_bwd(self, tape, batch, mode, gout, dEhat, overwrite=False):
        """Returns d shat [B, d]; adds the catalog gradient into dEhat [V, d] (overwrite=True: dEhat is a scratch
        buffer that may be stored to directly)."""
        Z, ldz, Ehat, shat = tape['Z'], tape['ldz'], tape['Ehat'], tape['shat']
        B, (V, d) = batch.B, Ehat.shape
        umma = tape['umma']
        Zlo = torch.empty_like(Z) if umma else None
        if mode == 'loss':
            ops.ce_rows_bwd(Z, ldz, batch.labels, tape['lse'], gout.reshape(1), tape['scale'], B, V, False, Zlo)
            dZ = Z
        else:
            dZ = torch.empty_like(Z)
            ops.logp_bwd(Z, ldz, gout, gout.stride(0), tape['scale'], B, V, dZ, ldz, Zlo)
        dshat = torch.zeros(B, d, dtype=torch.float32, device=Z.device)
        if umma:
            nkb = (V + 31) // 32
            split = max(1, min(nkb, 296 // ((B + 127) // 128)))
            ops.umma_gemm(1, B, d, V, dZ, Zlo, ldz, tape['Ehi'], tape['Elo'], d, dshat, d, accumulate=True, split_k=split)
            ops.umma_gemm(2, V, d, B, dZ, Zlo, ldz, tape['sh'], tape['sl'], d, dEhat, d, accumulate=not overwrite)
            return dshat
        ops.gemm(B, d, V, dZ, ldz, 1, Ehat, d, 1, dshat, d, accumulate=True, split_k=0)          # dZ @ Ehat
        ops.gemm(V, d, B, dZ, 1, ldz, shat, tape['ld_s'], 1, dEhat, d, accumulate=True, split_k=0)  # dZ^T @ shat
        return dshat

    # ---- fused training step (body of `TrainRunner.train`, utils/train.py:95-101) ---------------------------
    def configure_optimizer(self, lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8):
        fp = self._ensure_flat()
        seg_off, seg_decay = fp.decay_segments(weight_decay)
        self._opt = dict(lr=lr, betas=betas, eps=eps, step=0, seg_off=seg_off, seg_decay=seg_decay,
                         m=torch.zeros_like(fp.data), v=torch.zeros_like(fp.data), n_seg=len(fp.names))
        return self._opt

    def train_step(self, batch, group=None):
        """zero_grad + forward + nll_loss + backward + Adam step, all on the current stream; returns the loss
        as a 0-d device tensor (no host sync).  With a torch.distributed process group (data parallel: every rank
        owns a slice of the global batch) the flat gradient buffer is summed with ONE NCCL all-reduce and the
        1/world_size mean is folded into the Adam kernel."""
        fp = self._ensure_flat()
        if self._opt is None:
            self.configure_optimizer()
        o = self._opt
        with torch.no_grad():
            loss, tape = self._fwd(batch, 'loss', need_grad=True)
            ops.fill(fp.grad, 0.0)
            self._bwd(tape, self._one(), fp.grad)
            scale = 1.0
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(fp.grad, group=group)
                scale = 1.0 / dist.get_world_size(group)
            o['step'] += 1
            ops.adam_step(fp.data, fp.grad, o['m'], o['v'], o['seg_off'], o['seg_decay'], o['n_seg'], o['lr'],
                          o['betas'][0], o['betas'][1], o['eps'], o['step'], scale)
        return loss

    def _one(self):
        if getattr(self, '_one_t', None) is None or self._one_t.device != self._flat.data.device:
            self._one_t = torch.ones(1, dtype=torch.float32, device=self._flat.data.device)
        return self._one_t
