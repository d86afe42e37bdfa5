"""Shared machinery of the drop-in modules: flat parameters, the autograd bridge (one Function per model call;
the whole forward/backward below it is our kernels), the catalog scoring + cross-entropy head, fused training
step with the reference's Adam/L2 semantics."""
import os

import torch
from torch import nn

from . import _lib, ops
from ._lib import NORM_NONE
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
        self._shard = None
        self.fused_lse = os.environ.get('SESSREC_NO_FUSED_LSE', '0') != '1'
        self.head_chunks = int(os.environ.get('SESSREC_HEAD_CHUNKS', '1'))   # > 1 measured slower on B200 (r1g sweep)
        # fused scoring + CE head (csrc/flash_ce.cu): loss() / train_step() never materialise the (B, V) logits
        self.flash_ce = os.environ.get('SESSREC_NO_FLASH_CE', '0') != '1'

    # ---- parameters -------------------------------------------------------------------------------------
    def _ensure_flat(self):
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise _lib.SessRecError(f'{type(self).__name__}: parameters live on {p0.device}; this path only runs on '
                                    'CUDA (sm_100a) and has no CPU fallback - call model.to("cuda") first')
        if self._flat is None or not self._flat.valid(self):
            # the parameters were moved / re-assigned (model.to(), load_state_dict(assign=True), p.data = ...): re-flatten and
            # carry the optimizer over by parameter name - hyper-parameters, step count and both moments survive
            old = self.optimizer_state_dict() if self._opt is not None else None
            self._flat = FlatParams(self)
            self._opt = None
            self._native = None
            if old is not None:
                self.load_optimizer_state_dict(old)
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

    @torch.no_grad()
    def topk(self, mg, k=20):
        """Ids [B, k] (int64, best first) of the k highest-scoring items per session: the `logits.topk(k)[1]` of the
        reference's evaluate() (`utils/train.py:49`) without materialising (B, V) log-probabilities."""
        self._ensure_flat()
        if k <= 32 and self._shard is None and self._use_flash(self.embedding_dim, 'topk'):
            # fused: the scoring kernel's soft-max threads keep the k best logits of their rows while the tiles go by
            self._topk_k = int(k)
            idx, _ = self._fwd(mg, 'topk', need_grad=False)
            return idx.long()
        Zv, _ = self._fwd(mg, 'logits', need_grad=False)
        B, V = Zv.shape
        idx = torch.empty(B, k, dtype=torch.int32, device=Zv.device)
        ops.topk_rows(Zv, Zv.stride(0), B, V, k, idx)
        return idx.long()

    # ---- catalog scoring + CE head -----------------------------------------------------------------------
    def shard_catalog(self, group):
        """Catalog-row sharding (BASELINE config 5): this rank scores only rows [lo, hi) of the item table; per session
        the soft-max statistics + label logit cross NVLink in one small all-reduce, dS in a second one, and the head's
        share of the table gradient in a third.  The session encoder stays replicated."""
        self._shard = group

    def _rows(self, V):
        if self._shard is None:
            return 0, V
        import torch.distributed as dist
        from .parallel import shard_slice
        return shard_slice(V, dist.get_rank(self._shard), dist.get_world_size(self._shard))

    def _single_head(self):
        """False when the model scores several session representations against the catalog (order-fusion head)."""
        return True

    def _use_flash(self, d, mode):
        return (self.use_tensor_cores and self.flash_ce and mode in ('loss', 'topk') and self._single_head()
                and ops.flash_ce_supported(d))

    def _catalog_fwd(self, E, norm_mode, max_norm, tape):
        """Catalog pre-pass: in-place max_norm renorm (MSGIFSR) + row normalisation of the rows this rank scores, with
        the TF32 hi/lo split for the tcgen05 GEMM emitted by the same kernel."""
        V, d = E.shape
        lo, hi = self._rows(V)
        dev = E.device
        flash = self._use_flash(d, tape.get('mode'))
        umma = self.use_tensor_cores and d <= 256 and not flash
        if self._shard is not None and max_norm > 0:
            ops.renorm_rows(E, None, V, max_norm)          # every replica renorms every row (the reference does too)
            max_norm = 0.0
        El = E[lo:hi]
        Ehi = Elo = enorm = Bhi = Blo = None
        if flash:                                          # bf16 hi/lo pair of the scored rows (bit patterns)
            Bhi = torch.empty(hi - lo, d, dtype=torch.int16, device=dev)
            Blo = torch.empty(hi - lo, d, dtype=torch.int16, device=dev)
        if norm_mode == NORM_NONE:
            Ehat = El
            if umma:
                Ehi, Elo = torch.empty_like(El), torch.empty_like(El)
                ops.split_tf32(El, d, hi - lo, d, Ehi, Elo, d)
            if flash:
                ops.split_bf16(El, d, hi - lo, d, Bhi, Blo, d)
        else:
            Ehat = torch.empty_like(El)
            enorm = torch.empty(hi - lo, dtype=torch.float32, device=dev)
            if umma:
                Ehi, Elo = torch.empty_like(El), torch.empty_like(El)
            ops.catalog_prep_fwd(El, norm_mode, max_norm, Ehat, enorm, Ehi, Elo, Bhi, Blo)
        tape['cat'] = dict(lo=lo, hi=hi, Ehat=Ehat, enorm=enorm, Ehi=Ehi, Elo=Elo, Bhi=Bhi, Blo=Blo, norm_mode=norm_mode,
                           umma=umma, flash=flash)

    def _head_fwd(self, shat, ld_s, scale, batch, mode, tape):
        """Z = scale * shat Ehat^T on the tcgen05 tensor cores (3xTF32, csrc/umma_gemm.cu) whenever the embedding
        dim fits one UMMA N tile (d <= 256), else on the fp32 CUDA-core GEMM; then log-sum-exp / NLL / log-probs."""
        cat = tape['cat']
        Ehat, umma = cat['Ehat'], cat['umma']
        B, (V, d) = batch.B, Ehat.shape                 # V = rows scored by this rank
        dev = Ehat.device
        if cat['flash']:
            # fused head: soft-max statistics and the label logit straight from the tensor-core tiles
            Shi = torch.empty(B, d, dtype=torch.int16, device=dev)
            Slo = torch.empty(B, d, dtype=torch.int16, device=dev)
            ops.split_bf16(shat, ld_s, B, d, Shi, Slo, d)
            if mode == 'topk':
                idx = torch.empty(B, self._topk_k, dtype=torch.int32, device=dev)
                ops.flash_ce_topk(B, V, d, Shi, Slo, d, cat['Bhi'], cat['Blo'], d, scale, self._topk_k, idx)
                return idx
            lse = torch.empty(B, dtype=torch.float32, device=dev)
            nll = torch.empty(B, dtype=torch.float32, device=dev)
            part = torch.empty(ops.flash_ce_part_floats(B, V), dtype=torch.float32, device=dev)
            tape.update(Shi=Shi, Slo=Slo, scale=scale, shat=shat, ld_s=ld_s)
            if self._shard is not None:
                return self._flash_fwd_sharded(B, V, d, scale, batch, lse, nll, part, tape)
            ops.flash_ce_fwd(B, V, d, Shi, Slo, d, cat['Bhi'], cat['Blo'], d, scale, batch.labels, lse, nll, part)
            out = torch.empty((), dtype=torch.float32, device=dev)
            ops.mean(nll, B, out)
            tape.update(lse=lse, labels=batch.labels)
            return out
        ldz = (V + 3) // 4 * 4                      # 16-byte aligned rows: TMA / vector loads in the backward GEMMs
        Z = torch.empty(B, ldz, dtype=torch.float32, device=dev)
        lse = torch.empty(B, dtype=torch.float32, device=dev)
        fused_lse = umma and self.fused_lse and self._shard is None and mode in ('loss', 'logits')
        if umma:
            sh = torch.empty(B, d, dtype=torch.float32, device=dev)
            sl = torch.empty(B, d, dtype=torch.float32, device=dev)
            ops.split_tf32(shat, ld_s, B, d, sh, sl, d)
            tape.update(sh=sh, sl=sl)
            if fused_lse:
                # persistent tcgen05 kernel: Z and its row log-sum-exp (+ label logit) in one pass over the catalog
                part = torch.empty(4 * ((V + 255) // 256) * B + B, dtype=torch.float32, device=dev)
                nll = torch.empty(B, dtype=torch.float32, device=dev) if mode == 'loss' else None
                ops.umma_score_fwd(B, V, d, sh, sl, d, cat['Ehi'], cat['Elo'], d, Z, ldz, scale,
                                   batch.labels if mode == 'loss' else None, lse, nll, part)
            else:
                ops.umma_gemm(0, B, V, d, sh, sl, d, cat['Ehi'], cat['Elo'], d, Z, ldz, alpha=scale)
        else:
            ops.gemm(B, V, d, shat, ld_s, 1, Ehat, 1, d, Z, ldz, alpha=scale)
        tape.update(Z=Z, ldz=ldz, lse=lse, scale=scale, shat=shat, ld_s=ld_s)
        if fused_lse:
            if mode == 'logits':
                return Z[:, :V]
            out = torch.empty((), dtype=torch.float32, device=dev)
            ops.mean(nll, B, out)
            tape['labels'] = batch.labels
            return out
        if mode == 'logits':
            if self._shard is not None:
                raise _lib.SessRecError('catalog-sharded mode returns the loss only (each rank holds a slice of the logits)')
            return Z[:, :V]
        if self._shard is not None:
            return self._head_fwd_sharded(Z, ldz, lse, batch, mode, tape)
        if mode == 'loss':
            nll = torch.empty(B, dtype=torch.float32, device=dev)
            ops.ce_rows_fwd(Z, ldz, batch.labels, B, V, False, lse, nll)
            out = torch.empty((), dtype=torch.float32, device=dev)
            ops.mean(nll, B, out)
            tape['labels'] = batch.labels
            return out
        ops.ce_rows_fwd(Z, ldz, None, B, V, True, lse, None)
        return Z[:, :V]

    def _flash_fwd_sharded(self, B, V, d, scale, batch, lse_local, nll, part, tape):
        """Catalog-sharded fused head: this rank's rows give a local log-sum-exp and, where it owns the label, the label
        logit.  Cosine heads (|z| <= scale) exchange them in ONE [2, B] SUM all-reduce with a constant shift; unbounded
        logits (SRGNN) need a MAX all-reduce first."""
        import torch.distributed as dist
        cat = tape['cat']
        lo, hi = cat['lo'], cat['hi']
        lab = batch.labels
        own = (lab >= lo) & (lab < hi)
        ll = torch.where(own, lab - lo, torch.full_like(lab, -1))          # label column inside this shard, else -1
        ops.flash_ce_fwd(B, V, d, tape['Shi'], tape['Slo'], d, cat['Bhi'], cat['Blo'], d, scale, ll, lse_local, nll, part)
        zlab = torch.where(own, lse_local - nll, torch.zeros_like(nll))    # nll = lse_local - label logit where owned
        if cat['norm_mode'] != NORM_NONE:
            pack = torch.stack([torch.exp(lse_local - scale), zlab])
            dist.all_reduce(pack, group=self._shard)
            lse = scale + torch.log(pack[0])
        else:
            m = lse_local.clone()
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self._shard)
            pack = torch.stack([torch.exp(lse_local - m), zlab])
            dist.all_reduce(pack, group=self._shard)
            lse = m + torch.log(pack[0])
        tape.update(lse=lse, labels=ll)
        return (lse - pack[1]).mean()

    def _head_fwd_sharded(self, Z, ldz, lse_local, batch, mode, tape):
        import torch.distributed as dist
        if mode != 'loss':
            raise _lib.SessRecError('catalog-sharded mode supports loss() / train_step() only')
        cat, B = tape['cat'], batch.B
        lo, hi = cat['lo'], cat['hi']
        lab = batch.labels
        own = (lab >= lo) & (lab < hi)
        ll = torch.where(own, lab - lo, torch.full_like(lab, -1))          # label column inside this shard, else -1
        nll = torch.empty(B, dtype=torch.float32, device=Z.device)
        ops.ce_rows_fwd(Z, ldz, ll, B, hi - lo, False, lse_local, nll)      # local log-sum-exp; nll = 0 where not owned
        m = lse_local.clone()
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self._shard)
        pack = torch.stack([torch.exp(lse_local - m), torch.where(own, lse_local - nll, torch.zeros_like(nll))])
        dist.all_reduce(pack, group=self._shard)                           # [2, B]: sum of exp, label logit
        lse = m + torch.log(pack[0])
        tape['lse'], tape['labels'] = lse, ll
        return (lse - pack[1]).mean()

    def _head_bwd(self, tape, batch, mode, gout, gE, E):
        """Backward of the head: returns d shat [B, d] and adds the catalog's share of the table gradient into gE."""
        cat = tape['cat']
        if cat['flash']:
            Ehat = cat['Ehat']
            B, (V, d) = batch.B, Ehat.shape
            parts = ops.flash_ce_bwd_parts(B)
            dEpart = torch.empty(parts, V, d, dtype=torch.float32, device=Ehat.device)
            dshat = torch.empty(B, d, dtype=torch.float32, device=Ehat.device)
            ops.flash_ce_bwd(B, V, d, tape['Shi'], tape['Slo'], d, cat['Bhi'], cat['Blo'], d, tape['scale'], tape['labels'],
                             tape['lse'], gout.reshape(1), dshat, dEpart)
            if self._shard is not None:
                import torch.distributed as dist
                lo, hi = cat['lo'], cat['hi']
                dist.all_reduce(dshat, group=self._shard)     # every rank needs the full d shat for the replicated encoder
                ghead = torch.zeros_like(gE)
                if cat['norm_mode'] == NORM_NONE:
                    ops.sum_parts(dEpart, V * d, parts, V * d, ghead[lo:hi], accumulate=False)
                else:
                    ops.catalog_prep_bwd(E[lo:hi], Ehat, cat['enorm'], dEpart, cat['norm_mode'], ghead[lo:hi], nparts=parts)
                dist.all_reduce(ghead, group=self._shard)
                ops.dropout_apply(ghead, gE, ghead.numel(), None, accumulate=True)
            elif cat['norm_mode'] == NORM_NONE:          # SRGNN: the table itself is scored
                ops.sum_parts(dEpart, V * d, parts, V * d, gE, accumulate=True)
            else:
                ops.catalog_prep_bwd(E, Ehat, cat['enorm'], dEpart, cat['norm_mode'], gE, nparts=parts)
            return dshat
        Z, ldz, Ehat, shat, umma = tape['Z'], tape['ldz'], cat['Ehat'], tape['shat'], cat['umma']
        lo, hi = cat['lo'], cat['hi']
        B, (V, d) = batch.B, Ehat.shape
        dev = Z.device
        Zlo = torch.empty_like(Z) if umma else None
        chunked = mode == 'loss' and umma and self.head_chunks > 1
        if mode == 'loss':
            if not chunked:
                ops.ce_rows_bwd(Z, ldz, tape['labels'], tape['lse'], gout.reshape(1), tape['scale'], B, V, False, Zlo)
            dZ = Z
        else:
            dZ = torch.empty_like(Z)
            ops.logp_bwd(Z, ldz, gout, gout.stride(0), tape['scale'], B, V, dZ, ldz, Zlo)
        direct = cat['norm_mode'] == NORM_NONE and self._shard is None       # SRGNN: dE accumulates straight into gE
        if direct:
            dEhat = gE
        elif umma:
            dEhat = torch.empty(V, d, dtype=torch.float32, device=dev)
        else:
            dEhat = torch.zeros(V, d, dtype=torch.float32, device=dev)
        dshat = torch.zeros(B, d, dtype=torch.float32, device=dev)
        if umma:
            # chunked over catalog columns (fused-loss mode): each chunk's dZ hi/lo pair stays L2-resident between the
            # CE-backward pass that writes it and the two tensor-core GEMMs that read it
            Vc = ((V + self.head_chunks - 1) // self.head_chunks + 255) // 256 * 256 if chunked else V
            for c0 in range(0, V, Vc):
                nc = min(Vc, V - c0)
                if chunked:
                    ops.ce_rows_bwd_cols(Z, ldz, tape['labels'], tape['lse'], gout.reshape(1), tape['scale'], B, c0, nc, Zlo)
                split = max(1, min((nc + 31) // 32, 148 // ((B + 127) // 128)))
                ops.umma_gemm(1, B, d, nc, dZ[:, c0:], Zlo[:, c0:], ldz, cat['Ehi'][c0:], cat['Elo'][c0:], d, dshat, d,
                              accumulate=True, split_k=split)
                ops.umma_gemm(2, nc, d, B, dZ[:, c0:], Zlo[:, c0:], ldz, tape['sh'], tape['sl'], d, dEhat[c0:], d,
                              accumulate=direct)
        else:
            ops.gemm(B, d, V, dZ, ldz, 1, Ehat, d, 1, dshat, d, accumulate=True, split_k=0)          # dZ @ Ehat
            ops.gemm(V, d, B, dZ, 1, ldz, shat, tape['ld_s'], 1, dEhat, d, accumulate=True, split_k=0)  # dZ^T @ shat
        if self._shard is not None:
            import torch.distributed as dist
            dist.all_reduce(dshat, group=self._shard)
            ghead = torch.zeros_like(gE)
            if cat['norm_mode'] == NORM_NONE:
                ops.dropout_apply(dEhat, ghead[lo:hi], dEhat.numel(), None, accumulate=True)
            else:
                ops.catalog_prep_bwd(E[lo:hi], Ehat, cat['enorm'], dEhat, cat['norm_mode'], ghead[lo:hi])
            dist.all_reduce(ghead, group=self._shard)
            ops.dropout_apply(ghead, gE, ghead.numel(), None, accumulate=True)
        elif not direct:
            ops.catalog_prep_bwd(E, Ehat, cat['enorm'], dEhat, cat['norm_mode'], gE)
        return dshat

    # ---- fused training step (body of `TrainRunner.train`, utils/train.py:95-101) ---------------------------
    def _inactive_params(self, batch=None):
        """Names of the parameters the reference's forward never reaches in this configuration (their .grad stays None, so
        torch.optim.Adam never touches them: no update, no weight decay).  batch: some are data dependent."""
        return frozenset()

    def configure_optimizer(self, lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8):
        """torch.optim.Adam(fix_weight_decay(model), lr, weight_decay) of `TrainRunner.__init__` (`utils/train.py:70-74`)
        for the fused train_step: one flat buffer per moment, per-parameter decay / activity as segments."""
        fp = self._ensure_flat()
        seg_off, _ = fp.decay_segments(weight_decay)
        self._opt = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, step=0, seg_off=seg_off, seg_decay={},
                         m=torch.zeros_like(fp.data), v=torch.zeros_like(fp.data), n_seg=len(fp.names))
        return self._opt

    def set_lr(self, lr):
        """`scheduler.step()` of the reference loop (`utils/train.py:111`): the learning rate of the following steps."""
        self._ensure_flat()
        if self._opt is None:
            self.configure_optimizer(lr=lr)
        self._opt['lr'] = float(lr)

    def _segments(self, batch=None, owned_rows=None):
        """(seg_off, seg_decay, n_seg) of the Adam kernels for this batch: per-parameter weight decay, negative = inactive
        (parameters the reference's forward does not reach for this batch; under catalog sharding the table rows of the
        other owners).  Cached on the device by the batch's set of edge-less relations - the only data-dependent input -
        so a step pays one dict lookup."""
        o = self._opt
        key = (batch.empty_relations() if batch is not None else None, owned_rows)
        t = o['seg_decay'].get(key)
        if t is None:
            off, dec = self._flat.decay_segments(o['weight_decay'], self._inactive_params(batch), owned_rows)
            t = o['seg_decay'][key] = (off, dec, int(dec.numel()))
        return t

    def optimizer_state_dict(self):
        """Hyper-parameters, step count and the Adam moments keyed by parameter name (checkpoint / resume; the reference
        itself never saves anything)."""
        o, fp = self._opt, self._flat
        if o is None:
            raise _lib.SessRecError('optimizer_state_dict: call configure_optimizer() first')
        return dict(lr=o['lr'], betas=o['betas'], eps=o['eps'], weight_decay=o['weight_decay'], step=o['step'],
                    exp_avg={n: v.detach().clone() for n, v in zip(fp.names, fp.views(o['m']))},
                    exp_avg_sq={n: v.detach().clone() for n, v in zip(fp.names, fp.views(o['v']))})

    def load_optimizer_state_dict(self, sd):
        o = self.configure_optimizer(lr=sd['lr'], weight_decay=sd['weight_decay'], betas=sd['betas'], eps=sd['eps'])
        fp = self._flat
        missing = [n for n in fp.names if n not in sd['exp_avg'] or n not in sd['exp_avg_sq']]
        if missing:
            raise _lib.SessRecError(f'load_optimizer_state_dict: no moments for {missing[:3]} ...')
        for n, m, v in zip(fp.names, fp.views(o['m']), fp.views(o['v'])):
            m.copy_(sd['exp_avg'][n])
            v.copy_(sd['exp_avg_sq'][n])
        o['step'] = int(sd['step'])
        return o

    def train_step(self, batch, group=None, global_batch=None):
        """zero_grad + forward + nll_loss + backward + Adam step, all on the current stream; returns the loss
        as a 0-d device tensor (no host sync).  With a torch.distributed process group (data parallel: every rank
        owns a slice of the global batch) the flat gradient buffer is summed with ONE NCCL all-reduce.  Every rank's
        backward is seeded with B_local / B_global, so the sum is the gradient of the GLOBAL-batch mean loss - the
        reference's single-device step - also when the shards are uneven (last batch of an epoch); a rank whose shard
        is empty (batch None or B == 0) contributes zeros and still joins the collective.  global_batch: B_global when
        the caller knows it (equal shards: world * B); otherwise it is all-reduced (one extra 4-byte collective)."""
        fp = self._ensure_flat()
        if self._opt is None:
            self.configure_optimizer()
        o = self._opt
        with torch.no_grad():
            seed = self._dp_weight(batch, group, global_batch)
            if batch is None or batch.B == 0:
                loss = torch.zeros((), dtype=torch.float32, device=fp.data.device)
                ops.fill(fp.grad, 0.0)
            else:
                loss, tape = self._fwd(batch, 'loss', need_grad=True)
                ops.fill(fp.grad, 0.0)
                self._bwd(tape, seed, fp.grad)
            if group is not None:
                from . import parallel
                if parallel.comm_ready() and getattr(self, 'dp_allreduce_inside', False):
                    ops.comm_allreduce(fp.grad)            # same communicator as the native steps of the other ranks
                else:
                    import torch.distributed as dist
                    dist.all_reduce(fp.grad, group=group)
            o['step'] += 1
            seg_off, seg_decay, n_seg = self._segments(batch)
            ops.adam_step(fp.data, fp.grad, o['m'], o['v'], seg_off, seg_decay, n_seg, o['lr'],
                          o['betas'][0], o['betas'][1], o['eps'], o['step'], 1.0)
        return loss

    def _dp_weight(self, batch, group, global_batch):
        """Device scalar the backward is seeded with: 1 on a single device, B_local / B_global under data parallelism."""
        if group is None:
            return self._one()
        import torch.distributed as dist
        dev = self._flat.data.device
        b_local = 0 if batch is None else batch.B
        if global_batch is None:
            cnt = torch.tensor([float(b_local)], dtype=torch.float32, device=dev)
            tot = cnt.clone()
            dist.all_reduce(tot, group=group)
            return cnt / tot
        cache = self.__dict__.setdefault('_dp_w', {})
        key = (b_local, int(global_batch), str(dev))
        if key not in cache:
            cache[key] = torch.tensor([b_local / float(global_batch)], dtype=torch.float32, device=dev)
        return cache[key]

    def _one(self):
        if getattr(self, '_one_t', None) is None or self._one_t.device != self._flat.data.device:
            self._one_t = torch.ones(1, dtype=torch.float32, device=self._flat.data.device)
        return self._one_t
