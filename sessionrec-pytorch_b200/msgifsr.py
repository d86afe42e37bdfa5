"""Drop-in MSGIFSR module on the sm_100a kernels.

Constructor signature, parameter names / shapes / registration order (`state_dict` keys) follow the reference's
`src/models/msgifsr.py:159-227` (+ `gnn_models/gatconv.py:135-176` for the GAT modules and DGL's
`HeteroGraphConv.mods` ModuleDict); `forward(batch) -> (B, V) log-probabilities`.  Sub-modules only own
parameters; the arithmetic is in the C-ABI kernels."""
import math

import torch
from torch import nn

from . import ops
from ._lib import (HEADS, NORM_L2, SITE_EMBED, SITE_GAT_ATTN, SITE_GAT_DST, SITE_GAT_SRC, GatInst, SessRecError)
from .base import SessRecModule

SCALE = 12.0     # msgifsr.py:309


class GATConv(nn.Module):
    """Parameter holder of `gatconv.py:135-176` (num_heads=8, residual=Identity, bias)."""

    def __init__(self, in_feats, out_feats, num_heads=HEADS):
        super().__init__()
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.bias = nn.Parameter(torch.empty(num_heads * out_feats))


class HeteroGraphConv(nn.Module):
    def __init__(self, mods):
        super().__init__()
        self.mods = nn.ModuleDict(mods)


class MSHGNN(nn.Module):
    """Parameter holder of `msgifsr.py:47-68` (lint/linq/link and the PReLU activation are never used by the
    reference's forward but are part of its state_dict)."""

    def __init__(self, input_dim, output_dim, order):
        super().__init__()
        self.activation = nn.PReLU(output_dim)
        def mods():
            m = {f'intra{i + 1}': GATConv(input_dim, output_dim) for i in range(order)}
            m['inter'] = GATConv(input_dim, output_dim)
            return m
        self.conv1 = HeteroGraphConv(mods())
        self.conv2 = HeteroGraphConv(mods())
        self.lint = nn.Linear(output_dim, 1, bias=False)
        self.linq = nn.Linear(output_dim, output_dim)
        self.link = nn.Linear(output_dim, output_dim, bias=False)


class SemanticExpander(nn.Module):
    def __init__(self, input_dim, order):
        super().__init__()
        self.GRUs = nn.ModuleList([nn.GRU(input_dim, input_dim, 1, True, True) for _ in range(order)])


class AttnReadout(nn.Module):
    def __init__(self, input_dim, hidden_dim, order):
        super().__init__()
        self.fc_u = nn.ModuleList([nn.Linear(input_dim, hidden_dim, bias=True) for _ in range(order)])
        self.fc_v = nn.ModuleList([nn.Linear(input_dim, hidden_dim, bias=False) for _ in range(order)])
        self.fc_e = nn.ModuleList([nn.Linear(hidden_dim, 1, bias=False) for _ in range(order)])


def gat_slot(layer, conv, etype, st, dt, K):
    """Dense id of one (layer, conv, relation instance): dropout-site numbering shared with oracle/models.py."""
    if etype == 'inter':
        r = K + (dt - 2 if st == 1 else (K - 1) + st - 2)
    else:
        r = st - 1
    return (layer * 2 + conv) * (3 * K) + r


class MSGIFSR(SessRecModule):
    def __init__(self, num_items, datasets, embedding_dim, num_layers, dropout=0.0, reducer='mean', order=3, norm=True,
                 extra=True, fusion=True, device=torch.device('cpu')):
        super().__init__()
        if reducer != 'mean':
            raise SessRecError('only reducer="mean" (the reference default) is built')
        self.embeddings = nn.Embedding(num_items, embedding_dim, max_norm=1)
        self.num_items = num_items
        self.register_buffer('indices', torch.arange(num_items, dtype=torch.long))
        self.embedding_dim, self.num_layers, self.order = embedding_dim, num_layers, order
        self.reducer, self.norm = reducer, norm
        self.layers = nn.ModuleList()          # registered before the expander, like the reference (msgifsr.py:169)
        self.alpha = nn.Parameter(torch.empty(order))
        self.beta = nn.Parameter(torch.empty(1))
        self.expander = SemanticExpander(embedding_dim, order)
        for _ in range(num_layers):
            self.layers.append(MSHGNN(embedding_dim, embedding_dim, order))
        self.readout = AttnReadout(embedding_dim, embedding_dim, order)
        self.fc_sr = nn.ModuleList([nn.Linear(2 * embedding_dim, embedding_dim, bias=False) for _ in range(order)])
        self.sc_sr = nn.ModuleList([nn.Sequential(nn.Linear(embedding_dim, embedding_dim, bias=True), nn.ReLU(),
                                                  nn.Linear(embedding_dim, 2, bias=False), nn.Softmax(dim=-1))
                                    for _ in range(order)])
        self.dropout_p = float(dropout)
        self.reset_parameters()
        self.alpha.data = torch.zeros(order)
        self.alpha.data[0] = 1.0
        self.beta.data = torch.tensor(1.0)
        self.fusion, self.extra = fusion, extra
        self.native_step = True
        # data parallel: True = the gradient all-reduce is enqueued by the native step itself on the library's own communicator
        # (needs parallel.init_comm).  Off by default: equal to the torch.distributed path at 2 GPUs (0.465 vs 0.469 ms/step), but
        # at 8 GPUs back-to-back steps stall on it (1.6 vs 0.54 ms/step, profiles/r2o_*_8gpu.json) - not understood yet
        self.dp_allreduce_inside = False

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.embedding_dim)
        for weight in self.parameters():
            weight.data.uniform_(-stdv, stdv)

    # ---- one MSHGNN layer ------------------------------------------------------------------------------------
    def _instances(self, l, batch):
        """[(conv, rel, module_name, reversed)] for every relation instance with >= 1 edge, per destination type."""
        out = {k: [] for k in range(1, self.order + 1)}
        for conv in (0, 1):
            for rel in batch.rels:
                if rel['M'] == 0:
                    continue
                et = 'inter' if rel['name'].startswith('inter') else rel['name']
                st, dt = (rel['dt'], rel['st']) if conv else (rel['st'], rel['dt'])
                out[dt].append(dict(conv=conv, rel=rel, et=et, st=st, dt=dt,
                                    slot=gat_slot(l, conv, et, st, dt, self.order)))
        return out

    def _layer_fwd(self, l, batch, feats, p, seed, normalize, need_grad):
        d, dev = self.embedding_dim, self.embeddings.weight.device
        layer = self.layers[l]
        ltape = dict(feats=feats, types={})
        H = {}
        ldz = HEADS * d + HEADS
        by_dst = self._instances(l, batch)
        dc_attn = ops.drop_cfg(p, SITE_GAT_ATTN, seed) if p > 0 else None
        for k in range(1, self.order + 1):
            t = batch.types[k]
            N = t['N']
            insts, recs = [], []
            for it in by_dst[k]:
                mod = getattr(layer, f'conv{it["conv"] + 1}').mods[it['et']]
                rel, st = it['rel'], it['st']
                Ns = batch.types[st]['N']
                Waug = torch.empty(ldz, d, dtype=torch.float32, device=dev)
                wr = torch.empty(HEADS, d, dtype=torch.float32, device=dev)
                ops.gat_prep(mod.fc.weight, mod.attn_l, mod.attn_r, d, Waug, wr)
                xs, xd, dcs, dcd = feats[st], feats[k], None, None
                if p > 0:
                    dcs = ops.drop_cfg(p, SITE_GAT_SRC + 4 * it['slot'], seed)
                    dcd = ops.drop_cfg(p, SITE_GAT_DST + 4 * it['slot'], seed)
                    xs, xd = torch.empty_like(feats[st]), torch.empty_like(feats[k])
                    ops.dropout_apply(feats[st], xs, xs.numel(), dcs)
                    ops.dropout_apply(feats[k], xd, xd.numel(), dcd)
                Zel = torch.empty(Ns, ldz, dtype=torch.float32, device=dev)
                er = torch.empty(N, HEADS, dtype=torch.float32, device=dev)
                ops.linear_nt(xs, Waug, Zel)
                ops.linear_nt(xd, wr, er)
                att = torch.empty(max(rel['M'], 1), HEADS, dtype=torch.float32, device=dev)
                gi = GatInst()
                if it['conv'] == 0:
                    cs = (rel['in_ptr'], rel['in_src'], rel['in_eid'], rel['out_ptr'], rel['out_dst'], rel['out_eid'])
                else:          # reversed graph: roles of the two CSRs swap
                    cs = (rel['out_ptr'], rel['out_dst'], rel['out_eid'], rel['in_ptr'], rel['in_src'], rel['in_eid'])
                (gi.in_ptr, gi.in_src, gi.in_eid, gi.out_ptr, gi.out_dst, gi.out_eid) = (c.data_ptr() for c in cs)
                gi.Zel, gi.er, gi.bias, gi.xdst, gi.att = (Zel.data_ptr(), er.data_ptr(), mod.bias.data_ptr(),
                                                           xd.data_ptr(), att.data_ptr())
                gi.n_src, gi.n_dst, gi.n_edges = Ns, N, rel['M']
                gi.attn_site = SITE_GAT_ATTN + 4 * it['slot']
                insts.append(gi)
                recs.append(dict(it=it, mod=mod, Waug=Waug, wr=wr, xs=xs, xd=xd, dcs=dcs, dcd=dcd, Zel=Zel, er=er, att=att,
                                 cs=cs))
            segmean = torch.empty(batch.B, d, dtype=torch.float32, device=dev)
            ops.segmean_fwd(feats[k], t['seg'], batch.B, d, segmean)
            Hk = torch.empty(N, d, dtype=torch.float32, device=dev)
            rn = torch.empty(N, dtype=torch.float32, device=dev)
            amax = torch.empty(N, d, dtype=torch.uint8, device=dev)
            arr = ops.gat_inst_array(insts)
            ops.gat_aggregate_fwd(arr, len(insts), N, d, segmean, t['node2seg'], dc_attn, normalize, Hk, rn, amax)
            H[k] = Hk
            ltape['types'][k] = dict(insts=insts, recs=recs, arr=arr, H=Hk, rn=rn, amax=amax, normalize=normalize,
                                     dc_attn=dc_attn)
        return H, (ltape if need_grad else None)

    def _layer_bwd(self, l, batch, ltape, dH, g):
        """dH: {k: grad of the layer output}; returns {k: grad of the layer input}."""
        d, dev = self.embedding_dim, self.embeddings.weight.device
        ldz = HEADS * d + HEADS
        feats = ltape['feats']
        dfeat = {k: torch.empty_like(feats[k]) for k in feats}
        pending = []
        for k in range(1, self.order + 1):
            t, tt = batch.types[k], ltape['types'][k]
            N = t['N']
            for gi, rec in zip(tt['insts'], tt['recs']):
                M, Ns = rec['it']['rel']['M'], gi.n_src
                rec['dedge'] = torch.empty(max(M, 1), HEADS, dtype=torch.float32, device=dev)
                rec['der'] = torch.empty(N, HEADS, dtype=torch.float32, device=dev)
                rec['dZel'] = torch.empty(Ns, ldz, dtype=torch.float32, device=dev)
                gi.dedge, gi.der, gi.dZel = rec['dedge'].data_ptr(), rec['der'].data_ptr(), rec['dZel'].data_ptr()
            arr = ops.gat_inst_array(tt['insts'])
            dHpre = torch.empty(N, d, dtype=torch.float32, device=dev)
            ops.gat_aggregate_bwd_dst(arr, len(tt['insts']), N, d, tt['dc_attn'], tt['normalize'], tt['H'], tt['rn'],
                                      tt['amax'], dH[k], dHpre)
            ops.segmean_bwd(dHpre, t['seg'], batch.B, d, dfeat[k], False)        # first writer of dfeat[k]
            tt['dHpre'] = dHpre
            pending.append(k)
        for k in pending:
            t, tt = batch.types[k], ltape['types'][k]
            N, dHpre = t['N'], tt['dHpre']
            for gi, rec in zip(tt['insts'], tt['recs']):
                it, mod = rec['it'], rec['mod']
                name = f'layers.{l}.conv{it["conv"] + 1}.mods.{it["et"]}.'
                st, Ns = it['st'], gi.n_src
                ops.gat_bias_bwd(dHpre, tt['amax'], N, d, g(name + 'bias'))
                ops.gat_aggregate_bwd_src(gi, d, tt['dc_attn'], dHpre, tt['amax'])
                dWaug = torch.zeros(ldz, d, dtype=torch.float32, device=dev)
                dwr = torch.zeros(HEADS, d, dtype=torch.float32, device=dev)
                ops.mm_tn(rec['dZel'], rec['xs'], dWaug)
                ops.mm_tn(rec['der'], rec['xd'], dwr)
                ops.gat_prep_bwd(mod.fc.weight, mod.attn_l, mod.attn_r, dWaug, dwr, d, g(name + 'fc.weight'),
                                 g(name + 'attn_l'), g(name + 'attn_r'))
                if rec['dcs'] is None:
                    ops.mm_nn(rec['dZel'], rec['Waug'], dfeat[st], accumulate=True)
                    ops.mm_nn(rec['der'], rec['wr'], dfeat[k], accumulate=True)
                    ops.dropout_apply(dHpre, dfeat[k], dHpre.numel(), None, accumulate=True)     # residual
                else:
                    tmp = torch.empty(Ns, d, dtype=torch.float32, device=dev)
                    ops.mm_nn(rec['dZel'], rec['Waug'], tmp)
                    ops.dropout_apply(tmp, dfeat[st], tmp.numel(), rec['dcs'], accumulate=True)
                    tmp2 = dHpre.clone()
                    ops.mm_nn(rec['der'], rec['wr'], tmp2, accumulate=True)
                    ops.dropout_apply(tmp2, dfeat[k], tmp2.numel(), rec['dcd'], accumulate=True)
        return dfeat

    def _single_head(self):
        return not (self.order > 1 and self.fusion)

    def _use_flash(self, d, mode):
        # the REnorm head needs the session's own logits apart from the rest: it runs on the materialised scores
        return not self.extra and super()._use_flash(d, mode)

    def _inactive_params(self, batch=None):
        """Parameters whose .grad stays None in the reference: lint / linq / link, the PReLU, beta;
        alpha unless the order-fusion head runs; sc_sr unless the REnorm head runs (which only ever uses module 0,
        msgifsr.py:283); the GRU of the last order (GRUs[k - 2] serves order k); the `inter` convolutions at order 1; and,
        batch dependent, every convolution whose relation has no edge in this batch (HeteroGraphConv skips it)."""
        K = self.order
        rel_edges = None
        if batch is not None:
            rel_edges = {}
            for name, m in batch.rel_edge_counts().items():
                et = 'inter' if name.startswith('inter') else name
                rel_edges[et] = rel_edges.get(et, 0) + m
        out = []
        for n, _ in self.named_parameters():
            dead = False
            if n == 'beta' or '.lint.' in n or '.linq.' in n or '.link.' in n or '.activation.' in n:
                dead = True
            elif n == 'alpha':
                dead = not (K > 1 and self.fusion)
            elif n.startswith('sc_sr.'):
                dead = not (self.extra and n.startswith('sc_sr.0.'))
            elif n.startswith('expander.GRUs.'):
                dead = int(n.split('.')[2]) > K - 2
            elif '.mods.' in n:
                et = n.split('.mods.')[1].split('.')[0]
                dead = (et == 'inter' and K == 1) or (rel_edges is not None and rel_edges.get(et, 0) == 0)
            if dead:
                out.append(n)
        return frozenset(out)

    # ---- native fused step (csrc/step.cu) ------------------------------------------------------------------------
    def _native_ok(self, batch):
        from . import parallel
        # catalog sharding inside the native step needs this library's own communicator (parallel.init_comm); without it the
        # sharded head is composed from the staged kernels with torch.distributed collectives
        return (self.order == 1 and batch.K == 1 and not self.extra and self.norm and self.num_layers >= 1
                and (self._shard is None or parallel.comm_ready()))

    def _native_k_ok(self, batch):
        """The general native step (csrc/step_k.cu): any order, k-gram node types, without --extra / --fusion heads."""
        return (self.order > 1 and batch.K == self.order and not self.extra and not self.fusion and self.norm
                and self.num_layers >= 1 and self._shard is None and self.use_tensor_cores and self.flash_ce
                and ops.flash_ce_supported(self.embedding_dim))

    def _slot_names_k(self):
        """Parameter names in the slot order srk_msgifsr_k_train_step expects (include/sessrec_b200.h)."""
        K = self.order
        names = ['embeddings.weight']
        for l in range(self.num_layers):
            for c in (1, 2):
                for et in [f'intra{k}' for k in range(1, K + 1)] + ['inter']:
                    names += [f'layers.{l}.conv{c}.mods.{et}.{n}' for n in ('attn_l', 'attn_r', 'bias', 'fc.weight')]
        for k in range(2, K + 1):
            names += [f'expander.GRUs.{k - 2}.{n}' for n in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0')]
        names += ['readout.fc_u.0.weight', 'readout.fc_u.0.bias', 'readout.fc_v.0.weight', 'readout.fc_e.0.weight',
                  'fc_sr.0.weight']
        return names

    def _slot_offsets_k(self):
        import numpy as np
        fp = self._flat
        return np.ascontiguousarray([fp.offsets[fp.index[n]] for n in self._slot_names_k()], dtype=np.int64)

    def _train_step_k(self, batch, group, global_batch):
        """One TrainRunner iteration of an order-K model in ONE C call (srk_msgifsr_k_train_step)."""
        import ctypes
        from ._lib import lib, ptr
        fp = self._flat
        if self._opt is None:
            self.configure_optimizer()
        o = self._opt
        st = getattr(self, '_native_k', None)
        if st is None or st['flat'] is not fp:
            st = self._native_k = dict(flat=fp, slots=self._slot_offsets_k(), ws=None, ws_bytes=0)
        L = lib()
        hdr = ctypes.c_void_p(batch.hdr.ctypes.data)
        need = L.call('srk_msgifsr_k_workspace_bytes', hdr, self.num_items, self.embedding_dim, self.num_layers)
        if need > st['ws_bytes']:
            st['ws_bytes'] = int(need * 1.2)
            st['ws'] = torch.empty(st['ws_bytes'], dtype=torch.uint8, device=fp.data.device)
        p, seed = self._p(), self._next_seed()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        world = 1
        if group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(group)
        gseed = self._dp_weight(batch, group, global_batch)
        seg_off, seg_decay, n_seg = self._segments(batch, None)
        o['step'] += 1
        loss = torch.empty((), dtype=torch.float32, device=fp.data.device)

        def call(phase):
            ops._count[0] += 1
            L.call('srk_msgifsr_k_train_step', ptr(batch.buf), hdr, ptr(fp.data), ptr(fp.grad),
                   ctypes.c_void_p(st['slots'].ctypes.data), int(st['slots'].size), self.num_items, self.embedding_dim,
                   self.num_layers, float(p), ctypes.c_uint64(seed), 5, ptr(st['ws']), st['ws_bytes'], ptr(gseed), ptr(loss), 1,
                   ptr(o['m']), ptr(o['v']), fp.data.numel(), ptr(seg_off), ptr(seg_decay), n_seg, float(o['lr']),
                   float(o['betas'][0]), float(o['betas'][1]), float(o['eps']), int(o['step']), 1.0, phase, stream)
        if world == 1:
            call(0)
        else:
            call(1)
            dist.all_reduce(fp.grad, group=group)
            call(2)
        return loss

    def _slot_offsets(self):
        import numpy as np
        fp = self._flat
        names = ['embeddings.weight']
        for l in range(self.num_layers):
            for c in (1, 2):
                names += [f'layers.{l}.conv{c}.mods.intra1.{n}' for n in ('attn_l', 'attn_r', 'bias', 'fc.weight')]
        names += ['readout.fc_u.0.weight', 'readout.fc_u.0.bias', 'readout.fc_v.0.weight', 'readout.fc_e.0.weight',
                  'fc_sr.0.weight']
        return np.ascontiguousarray([fp.offsets[fp.index[n]] for n in names], dtype=np.int64)

    def train_step(self, batch, group=None, global_batch=None):
        """One TrainRunner iteration in ONE C call (srk_msgifsr_train_step): zero_grad, forward, nll_loss, backward,
        Adam.  Falls back to the staged Python composition for configurations the native step does not cover."""
        fp = self._ensure_flat()
        if batch is not None and batch.B > 0 and self.native_step and self._native_k_ok(batch):
            return self._train_step_k(batch, group, global_batch)
        if batch is None or batch.B == 0 or not self._native_ok(batch) or not self.native_step:
            return super().train_step(batch, group, global_batch)
        import ctypes
        from ._lib import lib, ptr
        if self._opt is None:
            self.configure_optimizer()
        o = self._opt
        st = getattr(self, '_native', None)
        if st is None or st['flat'] is not fp:
            st = self._native = dict(flat=fp, slots=self._slot_offsets(), ws=None, ws_bytes=0,
                                     loss=torch.zeros((), dtype=torch.float32, device=fp.data.device))
        L = lib()
        need = L.call('srk_msgifsr_workspace_bytes', batch.B, batch.N1, batch.M1, self.num_items, self.embedding_dim,
                      self.num_layers)
        if need > st['ws_bytes']:
            st['ws_bytes'] = int(need * 1.2)
            st['ws'] = torch.empty(st['ws_bytes'], dtype=torch.uint8, device=fp.data.device)
        p, seed = self._p(), self._next_seed()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        world = 1
        if group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(group)
        from . import parallel
        gseed = self._dp_weight(batch, group, global_batch)       # B_local / B_global: the all-reduced sum is the global mean
        shard = self._shard is not None
        owned = None
        if shard:
            if group is not None:
                raise SessRecError('catalog sharding and data parallelism are separate modes: pass group=None')
            lo, hi = self._rows(self.num_items)
            owned = ('embeddings.weight', lo, hi)
        seg_off, seg_decay, n_seg = self._segments(batch, owned)
        o['step'] += 1
        loss = torch.empty((), dtype=torch.float32, device=fp.data.device)

        def call(phase):
            ops._count[0] += 1
            L.call('srk_msgifsr_train_step', ptr(batch.buf), ctypes.c_void_p(batch.hdr.ctypes.data), ptr(fp.data),
                   ptr(fp.grad), ctypes.c_void_p(st['slots'].ctypes.data), self.num_items, self.embedding_dim,
                   self.num_layers, float(p), ctypes.c_uint64(seed),
                   int(self.use_tensor_cores) | (2 if self.fused_lse else 0) | (4 if self.flash_ce else 0) | (8 if shard else 0),
                   ptr(st['ws']), st['ws_bytes'], ptr(gseed), ptr(loss), 1, ptr(o['m']), ptr(o['v']), fp.data.numel(),
                   ptr(seg_off), ptr(seg_decay), n_seg, float(o['lr']), float(o['betas'][0]),
                   float(o['betas'][1]), float(o['eps']), int(o['step']), 1.0, phase, int(self.head_chunks), stream)
        if world == 1:
            ops.timed_native(lambda: call(0))
        elif parallel.comm_ready() and self.dp_allreduce_inside:
            call(3)             # the gradient all-reduce is enqueued by the step itself (csrc/comm.cu), no return to Python
        else:
            call(1)
            dist.all_reduce(fp.grad, group=group)
            call(2)
        return loss

    # ---- SemanticExpander for the k-gram node types (msgifsr.py:32-45) --------------------------------------------
    def _expander_fwd(self, k, batch, E, p, seed, need_grad):
        t = batch.types[k]
        N, d, dev = t['N'], self.embedding_dim, E.device
        gru = self.expander.GRUs[k - 2]
        dc = ops.drop_cfg(p, SITE_EMBED + k, seed) if p > 0 else None
        Xk = torch.empty(N * k, d, dtype=torch.float32, device=dev)
        ops.embed_gather_fwd(E, t['iid'], N * k, d, 0, dc, Xk, None)
        Xv = Xk.view(N, k * d)
        hs = [torch.zeros(N, d, dtype=torch.float32, device=dev)]
        gis, ghs = [], []
        for step in range(k):
            gi = torch.empty(N, 3 * d, dtype=torch.float32, device=dev)
            gh = torch.empty(N, 3 * d, dtype=torch.float32, device=dev)
            ops.linear_nt(Xv[:, step * d:(step + 1) * d], gru.weight_ih_l0, gi, bias=gru.bias_ih_l0)
            ops.linear_nt(hs[-1], gru.weight_hh_l0, gh, bias=gru.bias_hh_l0)
            hn = torch.empty(N, d, dtype=torch.float32, device=dev)
            ops.gru_pointwise_fwd(gi, gh, hs[-1], N, d, hn)
            hs.append(hn)
            gis.append(gi)
            ghs.append(gh)
        out = torch.empty(N, d, dtype=torch.float32, device=dev)
        rn = torch.empty(N, dtype=torch.float32, device=dev)
        ops.expander_combine_fwd(Xk, hs[-1], N, k, d, out, rn)
        return out, (dict(Xk=Xk, hs=hs, gis=gis, ghs=ghs, out=out, rn=rn, dc=dc) if need_grad else None)

    def _expander_bwd(self, k, batch, E, et, dout, g, gE):
        t = batch.types[k]
        N, d, dev = t['N'], self.embedding_dim, E.device
        gru = self.expander.GRUs[k - 2]
        name = f'expander.GRUs.{k - 2}.'
        dXk = torch.empty(N * k, d, dtype=torch.float32, device=dev)
        dh = torch.empty(N, d, dtype=torch.float32, device=dev)
        ops.expander_combine_bwd(et['out'], et['rn'], dout, N, k, d, dh, dXk)
        Xv, dXv = et['Xk'].view(N, k * d), dXk.view(N, k * d)
        for step in reversed(range(k)):
            gi, gh, hprev = et['gis'][step], et['ghs'][step], et['hs'][step]
            dprev = torch.empty(N, d, dtype=torch.float32, device=dev)
            ops.gru_pointwise_bwd(gi, gh, hprev, dh, N, d, dprev, False)           # gi, gh <- their gradients
            xs = Xv[:, step * d:(step + 1) * d]
            ops.mm_tn(gi, xs, g(name + 'weight_ih_l0'))
            ops.colsum(gi, 3 * d, N, 3 * d, g(name + 'bias_ih_l0'))
            ops.mm_nn(gi, gru.weight_ih_l0, dXv[:, step * d:(step + 1) * d], accumulate=True)
            ops.mm_tn(gh, hprev, g(name + 'weight_hh_l0'))
            ops.colsum(gh, 3 * d, N, 3 * d, g(name + 'bias_hh_l0'))
            ops.mm_nn(gh, gru.weight_hh_l0, dprev, accumulate=True)
            dh = dprev
        ops.embed_scatter_bwd(E, t, d, 0, et['dc'], None, dXk, None, gE)

    # ---- whole model -------------------------------------------------------------------------------------------
    def _fwd(self, batch, mode, need_grad=True):
        if batch.K != self.order:
            raise SessRecError(f'batch built for order {batch.K}, model has order {self.order}')
        if self.extra and self._shard is not None:
            raise SessRecError('catalog sharding with the REnorm head (extra=True) is not built')
        if not self.norm:
            raise SessRecError('MSGIFSR norm=False is not built (the reference argparse can only produce True)')
        if self.num_layers == 0:
            raise SessRecError('MSGIFSR needs num_layers >= 1')
        K, d, B, V, dev = self.order, self.embedding_dim, batch.B, self.num_items, self.embeddings.weight.device
        E = self.embeddings.weight.data
        p, seed = self._p(), self._next_seed()
        tape = dict(batch=batch, p=p, seed=seed, mode=mode)
        # nn.Embedding(max_norm=1): touched rows are renormed at the gather, all rows at the scoring head
        # (msgifsr.py:247,276).  Rows are independent, so one pre-pass over the catalog does both.
        self._catalog_fwd(E, NORM_L2, 1.0, tape)
        t1 = batch.types[1]
        N = t1['N']
        dc_e = ops.drop_cfg(p, SITE_EMBED + 1, seed) if p > 0 else None
        X = torch.empty(N, d, dtype=torch.float32, device=dev)
        rnX = torch.empty(N, dtype=torch.float32, device=dev)
        ops.embed_gather_fwd(E, t1['iid'], N, d, NORM_L2, dc_e, X, rnX)
        h, etapes = {1: X}, {}
        for k in range(2, K + 1):
            h[k], etapes[k] = self._expander_fwd(k, batch, E, p, seed, need_grad)
        ltapes = []
        for l in range(self.num_layers):
            h, lt = self._layer_fwd(l, batch, h, p, seed, normalize=(l == self.num_layers - 1), need_grad=need_grad)
            ltapes.append(lt)
        # readout over the per-session concatenation of all orders' rows (msgifsr.py:127-146).  Without fusion only
        # order 1's score is returned (msgifsr.py:316-317): the other heads are dead code and get exact-zero grads.
        if K == 1:
            rows, seg, last_row = h[1], t1['seg'], t1['last']
        else:
            rows = torch.zeros(batch.R, d, dtype=torch.float32, device=dev)
            for k in range(1, K + 1):
                ops.scatter_add_rows(h[k], d, batch.types[k]['row_of'], batch.types[k]['N'], d, rows)
            seg, last_row = batch.row_seg, t1['last_row']
        R = rows.shape[0]
        heads = list(range(K)) if (K > 1 and self.fusion) else [0]
        hd = []
        for i in heads:
            ti = batch.types[i + 1]
            u = torch.empty(R, d, dtype=torch.float32, device=dev)
            v = torch.empty(B, d, dtype=torch.float32, device=dev)
            ops.linear_nt(rows, self.readout.fc_u[i].weight, u, bias=self.readout.fc_u[i].bias)
            ops.linear_nt(h[i + 1], self.readout.fc_v[i].weight, v, M=B, a_idx=ti['last'])
            e = torch.empty(R, dtype=torch.float32, device=dev)
            ms = torch.empty(B, 2, dtype=torch.float32, device=dev)
            sr_in = torch.empty(B, 2 * d, dtype=torch.float32, device=dev)
            last_row = ti['last'] if K == 1 else ti['last_row']
            ops.readout_fwd(rows, u, v, self.readout.fc_e[i].weight, seg, last_row, B, d, True, e, ms, sr_in)
            s = torch.empty(B, d, dtype=torch.float32, device=dev)
            ops.linear_nt(sr_in, self.fc_sr[i].weight, s)
            shat = torch.empty_like(s)
            rn_s = torch.empty(B, dtype=torch.float32, device=dev)
            ops.rownorm_fwd(s, d, B, d, NORM_L2, shat, d, rn_s)
            hd.append(dict(i=i, u=u, v=v, e=e, ms=ms, sr_in=sr_in, s=s, shat=shat, rn_s=rn_s, last_row=last_row))
        tape.update(X=X, rnX=rnX, dc_e=dc_e, etapes=etapes, ltapes=ltapes, h=h, rows=rows, seg=seg, heads=hd)
        if len(heads) == 1 and self.extra:
            out = self._renorm_head_fwd(hd[0], batch, mode, tape)
        elif len(heads) == 1:
            out = self._head_fwd(hd[0]['shat'], d, SCALE, batch, mode, tape)
        else:
            out = self._fusion_head_fwd(hd, batch, mode, tape)
        return out, (tape if need_grad else None)

    # ---- REnorm head (`--extra`, msgifsr.py:281-305) -----------------------------------------------------------------
    def _scores(self, hk, cat, B, V, d, Z, ldz):
        """Z = 12 * shat Ehat^T, materialised (3xTF32 tcgen05 GEMM, or the fp32 CUDA-core GEMM without tensor cores)."""
        if cat['umma']:
            hk['sh'], hk['sl'] = torch.empty_like(hk['shat']), torch.empty_like(hk['shat'])
            ops.split_tf32(hk['shat'], d, B, d, hk['sh'], hk['sl'], d)
            ops.umma_gemm(0, B, V, d, hk['sh'], hk['sl'], d, cat['Ehi'], cat['Elo'], d, Z, ldz, alpha=SCALE)
        else:
            ops.gemm(B, V, d, hk['shat'], d, 1, cat['Ehat'], 1, d, Z, ldz, alpha=SCALE)

    def _scores_bwd(self, hk, cat, B, V, d, dZ, Zlo, ldz, dshat, dEhat):
        """dshat += dZ Ehat, dEhat += dZ^T shat (dZ as a TF32 hi/lo pair on the tensor-core path)."""
        if cat['umma']:
            split = max(1, min((V + 31) // 32, 148 // ((B + 127) // 128)))
            ops.umma_gemm(1, B, d, V, dZ, Zlo, ldz, cat['Ehi'], cat['Elo'], d, dshat, d, accumulate=True, split_k=split)
            ops.umma_gemm(2, V, d, B, dZ, Zlo, ldz, hk['sh'], hk['sl'], d, dEhat, d, accumulate=True)
        else:
            ops.gemm(B, d, V, dZ, ldz, 1, cat['Ehat'], d, 1, dshat, d, accumulate=True, split_k=0)
            ops.gemm(V, d, B, dZ, 1, ldz, hk['shat'], d, 1, dEhat, d, accumulate=True, split_k=0)

    def _renorm_fwd(self, hk, batch, Z, ldz, V):
        """Gate phi = sc_sr[0](shat) (every order uses module 0, msgifsr.py:283) and the in place rewrite of the scaled
        logits Z into log(phi_0 softmax_in + phi_1 softmax_ex)."""
        B, d, dev = batch.B, self.embedding_dim, Z.device
        t1 = batch.types[1]
        lin1, lin2 = self.sc_sr[0][0], self.sc_sr[0][2]
        hk['gate_h'] = torch.empty(B, d, dtype=torch.float32, device=dev)
        hk['lphi'] = torch.empty(B, 2, dtype=torch.float32, device=dev)
        ops.linear_nt(hk['shat'], lin1.weight, hk['gate_h'], bias=lin1.bias)
        ops.gate_fwd(hk['gate_h'], lin2.weight, B, d, hk['lphi'])
        zin = torch.empty(max(t1['N'], 1), dtype=torch.float32, device=dev)
        ops.renorm_head_fwd(Z, ldz, B, V, t1['iid'], t1['seg'], hk['lphi'], zin)

    def _renorm_bwd(self, hk, batch, LP, ldz, V, G, ldg, labels, gscale, scale, dl_scale, dZ, Zlo):
        """dZ (may alias LP or G) and the gradient at log phi of one REnorm head."""
        t1 = batch.types[1]
        dlphi = torch.empty(batch.B, 2, dtype=torch.float32, device=LP.device)
        tmp = torch.empty(2 * max(t1['N'], 1), dtype=torch.float32, device=LP.device)
        ops.renorm_head_bwd(LP, ldz, G, ldg, labels, gscale, scale, dl_scale, batch.B, V, t1['iid'], t1['seg'], hk['lphi'], tmp,
                            dZ, ldz, Zlo, dlphi)
        return dlphi

    def _gate_bwd(self, hk, dlphi, dshat, g):
        B, d, dev = dlphi.shape[0], self.embedding_dim, dlphi.device
        lin1, lin2 = self.sc_sr[0][0], self.sc_sr[0][2]
        da = torch.empty(B, 2, dtype=torch.float32, device=dev)
        dH = torch.empty(B, d, dtype=torch.float32, device=dev)
        ops.gate_bwd(hk['gate_h'], lin2.weight, hk['lphi'], dlphi, B, d, da, dH)
        ops.mm_tn(da, hk['gate_h'], g('sc_sr.0.2.weight'))
        ops.mm_tn(dH, hk['shat'], g('sc_sr.0.0.weight'))
        ops.colsum(dH, d, B, d, g('sc_sr.0.0.bias'))
        ops.mm_nn(dH, lin1.weight, dshat, accumulate=True)

    def _renorm_head_fwd(self, hk, batch, mode, tape):
        cat = tape['cat']
        B, (V, d) = batch.B, cat['Ehat'].shape
        dev = cat['Ehat'].device
        ldz = (V + 3) // 4 * 4
        Z = torch.empty(B, ldz, dtype=torch.float32, device=dev)
        self._scores(hk, cat, B, V, d, Z, ldz)
        self._renorm_fwd(hk, batch, Z, ldz, V)
        tape.update(Z=Z, ldz=ldz, renorm=True)
        if mode != 'loss':                    # 'logp' and 'logits' (top-k ranks the final score) are the same matrix here
            return Z[:, :V]
        lse = torch.empty(B, dtype=torch.float32, device=dev)        # ~ 0: the rewritten rows are log-probabilities
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        ops.ce_rows_fwd(Z, ldz, batch.labels, B, V, False, lse, nll)
        out = torch.empty((), dtype=torch.float32, device=dev)
        ops.mean(nll, B, out)
        return out

    def _renorm_head_bwd(self, tape, batch, mode, gout, gE, E, g):
        cat, hk = tape['cat'], tape['heads'][0]
        B, (V, d) = batch.B, cat['Ehat'].shape
        dev = cat['Ehat'].device
        Z, ldz = tape['Z'], tape['ldz']
        Zlo = torch.empty_like(Z) if cat['umma'] else None
        if mode == 'loss':
            dZ = Z                                # nobody else holds the log-probs: rewrite them in place
            dlphi = self._renorm_bwd(hk, batch, Z, ldz, V, None, 0, batch.labels, gout.reshape(1), SCALE, 1.0, dZ, Zlo)
        else:
            dZ = torch.empty_like(Z)              # Z is the tensor forward() returned to the caller
            dlphi = self._renorm_bwd(hk, batch, Z, ldz, V, gout, gout.stride(0), None, None, SCALE, 1.0, dZ, Zlo)
        dEhat = torch.zeros(V, d, dtype=torch.float32, device=dev)
        dshat = torch.zeros(B, d, dtype=torch.float32, device=dev)
        self._scores_bwd(hk, cat, B, V, d, dZ, Zlo, ldz, dshat, dEhat)
        ops.catalog_prep_bwd(E, cat['Ehat'], cat['enorm'], dEhat, NORM_L2, gE)
        self._gate_bwd(hk, dlphi, dshat, g)
        return dshat

    # ---- order-fusion head (msgifsr.py:311-315): log sum_k softmax(alpha)_k softmax(12 sr_k E^T) --------------------
    def _fusion_head_fwd(self, hd, batch, mode, tape):
        if self._shard is not None:
            raise SessRecError('catalog sharding with the order-fusion head is not built')
        cat = tape['cat']
        Ehat, umma = cat['Ehat'], cat['umma']
        K, B, (V, d) = len(hd), batch.B, Ehat.shape
        dev = Ehat.device
        ldz = (V + 3) // 4 * 4
        Zall = torch.empty(K, B, ldz, dtype=torch.float32, device=dev)
        lse = torch.empty(K, B, dtype=torch.float32, device=dev)
        nll = torch.empty(K, B, dtype=torch.float32, device=dev)
        for k, hk in enumerate(hd):
            if umma:
                hk['sh'], hk['sl'] = torch.empty_like(hk['shat']), torch.empty_like(hk['shat'])
                ops.split_tf32(hk['shat'], d, B, d, hk['sh'], hk['sl'], d)
                ops.umma_gemm(0, B, V, d, hk['sh'], hk['sl'], d, cat['Ehi'], cat['Elo'], d, Zall[k], ldz, alpha=SCALE)
            else:
                ops.gemm(B, V, d, hk['shat'], d, 1, Ehat, 1, d, Zall[k], ldz, alpha=SCALE)
            if self.extra:                      # REnorm inside every order's head: Zall[k] becomes log score_k (lse ~ 0)
                self._renorm_fwd(hk, batch, Zall[k], ldz, V)
            ops.ce_rows_fwd(Zall[k], ldz, batch.labels if mode == 'loss' else None, B, V, False, lse[k], nll[k])
        tape.update(Zall=Zall, ldz=ldz, lse_all=lse, fusion=True)
        if self.extra:
            tape['LPall'] = Zall.clone()        # mix_bwd rewrites Zall in place; the REnorm backward needs log score_k
        if mode == 'loss':
            out = torch.empty((), dtype=torch.float32, device=dev)
            ops.mix_loss_fwd(nll, self.alpha, K, B, out)
            return out
        out = torch.empty(B, ldz, dtype=torch.float32, device=dev)
        ops.mix_logp_fwd(Zall, B * ldz, ldz, lse, self.alpha, K, B, V, out, ldz)
        return out[:, :V]

    def _fusion_head_bwd(self, tape, batch, mode, gout, gE, E, g):
        cat, hd = tape['cat'], tape['heads']
        Ehat, umma = cat['Ehat'], cat['umma']
        K, B, (V, d) = len(hd), batch.B, Ehat.shape
        dev = Ehat.device
        Zall, ldz = tape['Zall'], tape['ldz']
        Zlo = torch.empty_like(Zall) if umma else None
        mix_lo = None if self.extra else Zlo     # REnorm: the mixture's gradient stays one fp32 matrix, split afterwards
        rsum = torch.empty(K, B, dtype=torch.float32, device=dev)
        if mode == 'loss':
            ops.mix_bwd(Zall, mix_lo, B * ldz, ldz, tape['lse_all'], self.alpha, K, B, V, None, 0, batch.labels, gout.reshape(1),
                        SCALE, rsum)
        else:
            ops.mix_bwd(Zall, mix_lo, B * ldz, ldz, tape['lse_all'], self.alpha, K, B, V, gout, gout.stride(0), None, None, SCALE,
                        rsum)
        ops.mix_alpha_bwd(rsum, self.alpha, K, B, g('alpha'))
        dEhat = torch.zeros(V, d, dtype=torch.float32, device=dev)
        out = []
        for k, hk in enumerate(hd):
            dshat = torch.zeros(B, d, dtype=torch.float32, device=dev)
            if self.extra:                      # Zall[k] holds SCALE * d loss / d log score_k
                dlphi = self._renorm_bwd(hk, batch, tape['LPall'][k], ldz, V, Zall[k], ldz, None, None, 1.0, 1.0 / SCALE,
                                         Zall[k], Zlo[k] if umma else None)
                self._gate_bwd(hk, dlphi, dshat, g)
            if umma:
                split = max(1, min((V + 31) // 32, 148 // ((B + 127) // 128)))
                ops.umma_gemm(1, B, d, V, Zall[k], Zlo[k], ldz, cat['Ehi'], cat['Elo'], d, dshat, d, accumulate=True, split_k=split)
                ops.umma_gemm(2, V, d, B, Zall[k], Zlo[k], ldz, hk['sh'], hk['sl'], d, dEhat, d, accumulate=True)
            else:
                ops.gemm(B, d, V, Zall[k], ldz, 1, Ehat, d, 1, dshat, d, accumulate=True, split_k=0)
                ops.gemm(V, d, B, Zall[k], 1, ldz, hk['shat'], d, 1, dEhat, d, accumulate=True, split_k=0)
            out.append(dshat)
        ops.catalog_prep_bwd(E, Ehat, cat['enorm'], dEhat, NORM_L2, gE)
        return out

    def _bwd(self, tape, gout, gflat):
        fp, batch = self._flat, tape['batch']
        t1 = batch.types[1]
        K, B, d, V = self.order, batch.B, self.embedding_dim, self.num_items
        E = self.embeddings.weight.data
        dev = E.device
        g = lambda name: fp.view(gflat, name)          # noqa: E731
        gE = g('embeddings.weight')
        hd, rows, h = tape['heads'], tape['rows'], tape['h']
        if tape.get('fusion'):
            dshats = self._fusion_head_bwd(tape, batch, tape['mode'], gout, gE, E, g)
        elif tape.get('renorm'):
            dshats = [self._renorm_head_bwd(tape, batch, tape['mode'], gout, gE, E, g)]
        else:
            dshats = [self._head_bwd(tape, batch, tape['mode'], gout, gE, E)]
        R = rows.shape[0]
        drows = None
        for hk, dshat in zip(hd, dshats):
            i = hk['i']
            ds = torch.empty(B, d, dtype=torch.float32, device=dev)
            ops.rownorm_bwd(hk['s'], d, hk['shat'], d, hk['rn_s'], dshat, d, B, d, NORM_L2, ds, d)
            dsr_in = torch.empty(B, 2 * d, dtype=torch.float32, device=dev)
            ops.mm_nn(ds, self.fc_sr[i].weight, dsr_in)
            ops.mm_tn(ds, hk['sr_in'], g(f'fc_sr.{i}.weight'))
            dr = torch.empty(R, d, dtype=torch.float32, device=dev)
            ops.readout_bwd(rows, hk['u'], hk['v'], self.readout.fc_e[i].weight, tape['seg'], hk['last_row'], hk['e'], hk['ms'],
                            hk['sr_in'], dsr_in, B, d, True, dr, g(f'readout.fc_e.{i}.weight'))
            ops.mm_nn(hk['u'], self.readout.fc_u[i].weight, dr, accumulate=True)            # u holds du
            ops.mm_tn(hk['u'], rows, g(f'readout.fc_u.{i}.weight'))
            ops.colsum(hk['u'], d, R, d, g(f'readout.fc_u.{i}.bias'))
            if drows is None:
                drows = dr
            else:
                ops.dropout_apply(dr, drows, dr.numel(), None, accumulate=True)
        if K == 1:
            dH = {1: drows}
        else:
            dH = {}
            for k in range(1, K + 1):
                Nk = batch.types[k]['N']
                dH[k] = torch.empty(Nk, d, dtype=torch.float32, device=dev)
                ops.gather_rows(drows, batch.types[k]['row_of'], Nk, d, dH[k], d)
        for hk in hd:
            i = hk['i']
            ti = batch.types[i + 1]
            ops.mm_nn(hk['v'], self.readout.fc_v[i].weight, dH[i + 1], c_idx=ti['last'], accumulate=True)     # v holds dv
            ops.mm_tn(hk['v'], h[i + 1], g(f'readout.fc_v.{i}.weight'), b_idx=ti['last'])
        for l in reversed(range(self.num_layers)):
            dH = self._layer_bwd(l, batch, tape['ltapes'][l], dH, g)
        for k in range(2, K + 1):
            self._expander_bwd(k, batch, E, tape['etapes'][k], dH[k], g, gE)
        ops.embed_scatter_bwd(E, t1, d, NORM_L2, tape['dc_e'], tape['rnX'], dH[1], None, gE)
