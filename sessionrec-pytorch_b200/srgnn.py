"""Drop-in SRGNN / NISER modules on the sm_100a kernels.

Same constructor signatures, parameter names / shapes / registration order (hence `state_dict` keys and the
`reset_parameters` RNG consumption) as the reference's `src/models/srgnn.py:95-129` and `src/models/niser.py:93-128`;
`forward(batch) -> (B, V) log-probabilities`.  The sub-modules below only own parameters: all arithmetic runs in
the C-ABI kernels (ops.py)."""
import math

import torch
from torch import nn

from . import ops
from ._lib import NORM_EPS, NORM_NISER, NORM_NONE, SITE_EMBED, SITE_GGNN, SITE_READOUT
from .base import SessRecModule


class SRGNNLayer(nn.Module):
    """Parameter holder of the gated graph layer (`srgnn.py:11-19`).  The reference evaluates these layers and then
    discards their output (`srgnn.py:135-142`); so do we (see SRGNN.compute_dead_layers)."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.gru = nn.GRUCell(2 * input_dim, output_dim)
        self.W1 = nn.Linear(input_dim, output_dim, bias=False)
        self.W2 = nn.Linear(input_dim, output_dim, bias=False)


class AttnReadout(nn.Module):
    """Parameter holder of `srgnn.py:53-74` (batch_norm=None, fc_out=None in both models)."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.fc_u = nn.Linear(input_dim, hidden_dim, bias=False)
        self.fc_v = nn.Linear(input_dim, hidden_dim, bias=True)
        self.fc_e = nn.Linear(hidden_dim, 1, bias=False)


def ggnn_layer_fwd(layer, batch, x, p, seed, site, need_grad=True):
    """`SRGNNLayer.forward` (`srgnn.py:31-51`): returns (h_new, tape)."""
    t, rel = batch.types[1], batch.rels[0]
    N, d, dev = t['N'], x.shape[1], x.device
    ft = x
    dc = None
    if p > 0:
        dc = ops.drop_cfg(p, site, seed)
        ft = torch.empty_like(x)
        ops.dropout_apply(x, ft, x.numel(), dc)
    NN = torch.empty(N, 2 * d, dtype=torch.float32, device=dev)
    wsum = torch.empty(N, 2, dtype=torch.float32, device=dev)
    ops.ggnn_aggregate_fwd(ft, N, d, rel, NN, wsum)
    hn = torch.empty(N, 2 * d, dtype=torch.float32, device=dev)
    ops.linear_nt(NN, layer.W1.weight, hn, lda=2 * d, ldc=2 * d)
    ops.linear_nt(NN[:, d:], layer.W2.weight, hn[:, d:], lda=2 * d, ldc=2 * d)
    gi = torch.empty(N, 3 * d, dtype=torch.float32, device=dev)
    gh = torch.empty(N, 3 * d, dtype=torch.float32, device=dev)
    ops.linear_nt(hn, layer.gru.weight_ih, gi, bias=layer.gru.bias_ih)
    ops.linear_nt(x, layer.gru.weight_hh, gh, bias=layer.gru.bias_hh)
    hnew = torch.empty(N, d, dtype=torch.float32, device=dev)
    ops.gru_pointwise_fwd(gi, gh, x, N, d, hnew)
    tape = dict(x=x, NN=NN, wsum=wsum, hn=hn, gi=gi, gh=gh, dc=dc) if need_grad else None
    return hnew, tape


def ggnn_layer_bwd(layer, batch, tape, dhnew, g, dx, accumulate):
    """Backward of the layer.  g: dict name -> gradient view for W1.weight, W2.weight, gru.*; dx (+)= d loss / d x."""
    t, rel = batch.types[1], batch.rels[0]
    x, NN, hn, gi, gh = tape['x'], tape['NN'], tape['hn'], tape['gi'], tape['gh']
    N, d, dev = t['N'], x.shape[1], x.device
    ops.gru_pointwise_bwd(gi, gh, x, dhnew, N, d, dx, accumulate)          # gi, gh now hold their gradients
    ops.mm_tn(gi, hn, g['gru.weight_ih'])
    ops.colsum(gi, 3 * d, N, 3 * d, g['gru.bias_ih'])
    ops.mm_tn(gh, x, g['gru.weight_hh'])
    ops.colsum(gh, 3 * d, N, 3 * d, g['gru.bias_hh'])
    ops.mm_nn(gh, layer.gru.weight_hh, dx, accumulate=True)
    dhn = torch.empty(N, 2 * d, dtype=torch.float32, device=dev)
    ops.mm_nn(gi, layer.gru.weight_ih, dhn)
    ops.mm_tn(dhn, NN, g['W1.weight'], M=d, N=d, lda=2 * d, ldb=2 * d)
    ops.mm_tn(dhn[:, d:], NN[:, d:], g['W2.weight'], M=d, N=d, lda=2 * d, ldb=2 * d)
    dNN = torch.empty(N, 2 * d, dtype=torch.float32, device=dev)
    ops.mm_nn(dhn, layer.W1.weight, dNN, lda=2 * d, ldc=2 * d)
    ops.mm_nn(dhn[:, d:], layer.W2.weight, dNN[:, d:], lda=2 * d, ldc=2 * d)
    if tape['dc'] is None:
        ops.ggnn_aggregate_bwd(dNN, N, d, rel, tape['wsum'], dx, True)
    else:
        dft = torch.empty(N, d, dtype=torch.float32, device=dev)
        ops.ggnn_aggregate_bwd(dNN, N, d, rel, tape['wsum'], dft, False)
        ops.dropout_apply(dft, dx, dft.numel(), tape['dc'], accumulate=True)


class SRGNN(SessRecModule):
    def __init__(self, num_items, embedding_dim, num_layers, feat_drop=0.0):
        super().__init__()
        self._build(num_items, embedding_dim, num_layers, feat_drop)
        self.niser, self.scale = False, 1.0
        self.reset_parameters()

    def _build(self, num_items, embedding_dim, num_layers, feat_drop):
        self.embedding = nn.Embedding(num_items, embedding_dim)
        self.register_buffer('indices', torch.arange(num_items, dtype=torch.long))
        self.embedding_dim, self.num_layers, self.num_items = embedding_dim, num_layers, num_items
        self.layers = nn.ModuleList([SRGNNLayer(embedding_dim, embedding_dim) for _ in range(num_layers)])
        self.readout = AttnReadout(embedding_dim, embedding_dim)
        self.fc_sr = nn.Linear(2 * embedding_dim, embedding_dim, bias=False)
        self.dropout_p = float(feat_drop)
        # The reference runs the GGNN layers and drops their result; keep the same amount of device work by default.
        self.compute_dead_layers = True
        self.native_step = True
        # data parallel: True = the gradient all-reduce is enqueued by the native step itself on the library's own communicator
        # (needs parallel.init_comm).  Off by default: equal to the torch.distributed path at 2 GPUs (0.465 vs 0.469 ms/step), but
        # at 8 GPUs back-to-back steps stall on it (1.6 vs 0.54 ms/step, profiles/r2o_*_8gpu.json) - not understood yet
        self.dp_allreduce_inside = False

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.embedding_dim)
        for weight in self.parameters():
            weight.data.uniform_(-stdv, stdv)

    def _inactive_params(self, batch=None):
        # the GGNN layers are evaluated and discarded (srgnn.py:135-142): their parameters never get a gradient
        return frozenset(n for n, _ in self.named_parameters() if n.startswith('layers.'))

    # ---- native fused step (csrc/step_srgnn.cu) ------------------------------------------------------------------
    def _slot_offsets(self):
        import numpy as np
        fp = self._flat
        names = ['embedding.weight']
        for l in range(self.num_layers):
            names += [f'layers.{l}.{n}' for n in ('gru.weight_ih', 'gru.weight_hh', 'gru.bias_ih', 'gru.bias_hh', 'W1.weight',
                                                  'W2.weight')]
        names += ['readout.fc_u.weight', 'readout.fc_v.weight', 'readout.fc_v.bias', 'readout.fc_e.weight', 'fc_sr.weight']
        return np.ascontiguousarray([fp.offsets[fp.index[n]] for n in names], dtype=np.int64)

    def train_step(self, batch, group=None, global_batch=None):
        """One TrainRunner iteration (`utils/train.py:95-101`) in ONE C call (srk_srgnn_train_step): zero_grad, forward,
        nll_loss, backward, Adam.  The catalog-sharded head is composed from the staged kernels instead."""
        fp = self._ensure_flat()
        if batch is None or batch.B == 0 or not self.native_step or self._shard is not None or batch.kind != 'session':
            return super().train_step(batch, group, global_batch)
        import ctypes
        from . import parallel
        from ._lib import lib, ptr
        if self._opt is None:
            self.configure_optimizer()
        o = self._opt
        st = getattr(self, '_native', None)
        if st is None or st['flat'] is not fp:
            st = self._native = dict(flat=fp, slots=self._slot_offsets(), ws=None, ws_bytes=0)
        L = lib()
        V, d = self.num_items, self.embedding_dim
        need = L.call('srk_srgnn_workspace_bytes', batch.B, batch.N1, batch.M1, V, d, self.num_layers)
        if need > st['ws_bytes']:
            st['ws_bytes'] = int(need * 1.2)
            st['ws'] = torch.empty(st['ws_bytes'], dtype=torch.uint8, device=fp.data.device)
        p, seed = self._p(), self._next_seed()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        gseed = self._dp_weight(batch, group, global_batch)
        seg_off, seg_decay, n_seg = self._segments(batch)
        o['step'] += 1
        loss = torch.empty((), dtype=torch.float32, device=fp.data.device)
        flags = int(self.use_tensor_cores) | (2 if self.fused_lse else 0) | (4 if self.flash_ce else 0)

        def call(phase):
            ops._count[0] += 1
            L.call('srk_srgnn_train_step', ptr(batch.buf), ctypes.c_void_p(batch.hdr.ctypes.data), ptr(fp.data), ptr(fp.grad),
                   ctypes.c_void_p(st['slots'].ctypes.data), V, d, self.num_layers, int(self.niser),
                   float(self.scale) if self.scale else 1.0, int(self.compute_dead_layers), float(p), ctypes.c_uint64(seed),
                   flags, ptr(st['ws']), st['ws_bytes'], ptr(gseed), ptr(loss), 1, ptr(o['m']), ptr(o['v']), fp.data.numel(),
                   ptr(seg_off), ptr(seg_decay), n_seg, float(o['lr']), float(o['betas'][0]), float(o['betas'][1]),
                   float(o['eps']), int(o['step']), 1.0, phase, stream)
        if group is None:
            call(0)
        elif parallel.comm_ready() and self.dp_allreduce_inside:
            call(3)             # the gradient all-reduce is enqueued by the step itself (csrc/comm.cu)
        else:
            import torch.distributed as dist
            call(1)
            dist.all_reduce(fp.grad, group=group)
            call(2)
        return loss

    # ---- forward / backward over the kernels -----------------------------------------------------------------
    def _fwd(self, batch, mode, need_grad=True):
        t = batch.types[1]
        N, B, d, dev = t['N'], batch.B, self.embedding_dim, self.embedding.weight.device
        E = self.embedding.weight.data
        p, seed = self._p(), self._next_seed()
        tape = dict(batch=batch, p=p, seed=seed, mode=mode)
        emb_mode = NORM_NISER if self.niser else NORM_NONE
        dc_e = ops.drop_cfg(p, SITE_EMBED + 1, seed) if p > 0 else None
        X = torch.empty(N, d, dtype=torch.float32, device=dev)
        rn = torch.empty(N, dtype=torch.float32, device=dev) if self.niser else None
        dead = self.compute_dead_layers and self.num_layers > 0
        x_first = torch.empty(N, d, dtype=torch.float32, device=dev) if (self.niser and dead) else None
        ops.embed_gather_fwd(E, t['iid'], N, d, emb_mode, dc_e, X, rn, x_first)
        if dead:
            out = x_first if self.niser else X
            for l, layer in enumerate(self.layers):
                out, _ = ggnn_layer_fwd(layer, batch, out, p, seed, SITE_GGNN + l, need_grad=False)
        # readout (with its own feat_drop on top of the embedding dropout, srgnn.py:79)
        F, dc_r = X, None
        if p > 0:
            dc_r = ops.drop_cfg(p, SITE_READOUT, seed)
            F = torch.empty_like(X)
            ops.dropout_apply(X, F, X.numel(), dc_r)
        u = torch.empty(N, d, dtype=torch.float32, device=dev)
        v = torch.empty(B, d, dtype=torch.float32, device=dev)
        ops.linear_nt(F, self.readout.fc_u.weight, u)
        ops.linear_nt(F, self.readout.fc_v.weight, v, M=B, a_idx=t['last'], bias=self.readout.fc_v.bias)
        e = torch.empty(N, dtype=torch.float32, device=dev)
        ms = torch.empty(B, 2, dtype=torch.float32, device=dev)
        sr_in = torch.empty(B, 2 * d, dtype=torch.float32, device=dev)
        ops.readout_fwd(F, u, v, self.readout.fc_e.weight, t['seg'], t['last'], B, d, p == 0, e, ms, sr_in)
        if p > 0:
            ops.gather_rows(X, t['last'], B, d, sr_in, 2 * d)          # sr_l uses the once-dropped rows
        s = torch.empty(B, d, dtype=torch.float32, device=dev)
        ops.linear_nt(sr_in, self.fc_sr.weight, s)
        if self.niser:
            shat = torch.empty_like(s)
            rn_s = torch.empty(B, dtype=torch.float32, device=dev)
            ops.rownorm_fwd(s, d, B, d, NORM_EPS, shat, d, rn_s)
            self._catalog_fwd(E, NORM_EPS, 0.0, tape)
            tape.update(s=s, rn_s=rn_s)
        else:
            shat = s
            self._catalog_fwd(E, NORM_NONE, 0.0, tape)
        tape.update(X=X, rn=rn, F=F, u=u, v=v, e=e, ms=ms, sr_in=sr_in, dc_e=dc_e, dc_r=dc_r, emb_mode=emb_mode)
        out = self._head_fwd(shat, d, float(self.scale) if self.scale else 1.0, batch, mode, tape)
        return out, (tape if need_grad else None)

    def _bwd(self, tape, gout, gflat):
        fp, batch = self._flat, tape['batch']
        t = batch.types[1]
        N, B, d, V = t['N'], batch.B, self.embedding_dim, self.num_items
        E = self.embedding.weight.data
        dev = E.device
        g = lambda name: fp.view(gflat, name)          # noqa: E731
        gE = g('embedding.weight')
        if self.niser:
            dshat = self._head_bwd(tape, batch, tape['mode'], gout, gE, E)
            ds = torch.empty(B, d, dtype=torch.float32, device=dev)
            ops.rownorm_bwd(tape['s'], d, tape['shat'], d, tape['rn_s'], dshat, d, B, d, NORM_EPS, ds, d)
        else:
            ds = self._head_bwd(tape, batch, tape['mode'], gout, gE, E)
        sr_in, F, u, v = tape['sr_in'], tape['F'], tape['u'], tape['v']
        dsr_in = torch.empty(B, 2 * d, dtype=torch.float32, device=dev)
        ops.mm_nn(ds, self.fc_sr.weight, dsr_in)
        ops.mm_tn(ds, sr_in, g('fc_sr.weight'))
        p = tape['p']
        dF = torch.empty(N, d, dtype=torch.float32, device=dev)
        ops.readout_bwd(F, u, v, self.readout.fc_e.weight, t['seg'], t['last'], tape['e'], tape['ms'], sr_in, dsr_in, B,
                        d, p == 0, dF, g('readout.fc_e.weight'))
        ops.mm_nn(u, self.readout.fc_u.weight, dF, accumulate=True)                          # u holds du
        ops.mm_tn(u, F, g('readout.fc_u.weight'))
        ops.mm_nn(v, self.readout.fc_v.weight, dF, c_idx=t['last'], accumulate=True)         # v holds dv
        ops.mm_tn(v, F, g('readout.fc_v.weight'), b_idx=t['last'])
        ops.colsum(v, d, B, d, g('readout.fc_v.bias'))
        if p > 0:
            dX = torch.empty_like(dF)
            ops.dropout_apply(dF, dX, dF.numel(), tape['dc_r'])
            ops.scatter_add_rows(dsr_in, 2 * d, t['last'], B, d, dX)
        else:
            dX = dF
        ops.embed_scatter_bwd(E, t, d, tape['emb_mode'], tape['dc_e'], tape['rn'], dX, None, gE)


class NISER(SRGNN):
    def __init__(self, num_items, embedding_dim, num_layers, feat_drop=0.0, norm=True, scale=12):
        SessRecModule.__init__(self)
        self._build(num_items, embedding_dim, num_layers, feat_drop)
        self.norm, self.scale = norm, scale
        self.niser = bool(norm)
        self.reset_parameters()
