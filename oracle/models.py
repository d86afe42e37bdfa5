"""CPU fp32 restatement (plain PyTorch ops + autograd) of the reference's per-batch training path.

TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the CUDA path and the "port" CPU baseline
that `bench.py` times; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The product package never does.

Parity status: PINNED against the unmodified reference (`/root/reference/src/models/*.py` run over
`oracle/dgl_shim`, because DGL 0.7.2 itself is un-vendored and un-installable here) through the golden
vectors in `tests/golden/` (`oracle/make_golden.py`, `tests/test_oracle_models.py`).  The DGL
primitives themselves remain "parity unpinned" against a real DGL build (none reachable); their
restated semantics are listed in `oracle/dgl_shim/dgl/__init__.py`.

Everything works on the flat batch of `oracle/collate.py` and on a `params` dict keyed by the
reference's own `state_dict` names, so weights are interchangeable with the reference and the product.
"""
import math

import numpy as np
import torch as th
import torch.nn.functional as TF

HEADS = 8            # `msgifsr.py:58-62`: GATConv(..., 8, ...)
NEG_SLOPE = 0.2      # `gatconv.py:143`
SCALE = 12.0         # `msgifsr.py:288,309`, `niser.py:93`

# ---- dropout -------------------------------------------------------------------------------------
# Dropout sites (shared with the CUDA path, see include/sessrec_b200.h SRK_SITE_*): a site id and the
# element's flat index select the counter-based random number, so masks are reproducible on both sides.
SITE_EMBED = 0x100        # + k (order)           models' feat_drop on the gathered rows
SITE_READOUT = 0x200      #                       SRGNN/NISER readout feat_drop
SITE_GGNN = 0x300         # + layer               SRGNNLayer.dropout
SITE_GAT_SRC = 0x1000     # + slot * 4            GATConv.feat_drop on the source copy
SITE_GAT_DST = 0x1001     # + slot * 4            GATConv.feat_drop on the destination copy
SITE_GAT_ATTN = 0x1002    # + slot * 4            GATConv.attn_drop on the edge attention


def counter_uniform24(seed, site, n):
    """24-bit uniforms from a splitmix64 finaliser of (seed, site, index); the CUDA kernels use the
    same arithmetic (`csrc/common.cuh: srk_rand24`)."""
    M = (1 << 64) - 1
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over='ignore'):
        x = (np.uint64(seed & M) + np.uint64(0x9E3779B97F4A7C15) * ((np.uint64(site) << np.uint64(40)) + idx + np.uint64(1)))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(40)).astype(np.int64)


def keep_threshold(p):
    return int(math.floor(p * float(1 << 24)))


class Dropout:
    """Site-aware dropout.  seed=None: torch's own RNG (same stream as the reference when called in the
    same order with the same shapes).  seed=int: the counter-based masks the CUDA kernels regenerate."""

    def __init__(self, p=0.0, training=False, seed=None):
        self.p, self.training, self.seed = float(p), bool(training), seed

    def __call__(self, x, site):
        if not self.training or self.p == 0.0:
            return x
        if self.seed is None:
            return TF.dropout(x, self.p, True)
        u = counter_uniform24(self.seed, site, x.numel())
        keep = th.from_numpy((u >= keep_threshold(self.p)).astype(np.float32)).reshape(x.shape)
        return x * (keep * np.float32(1.0 / (1.0 - self.p)))


NO_DROPOUT = Dropout()

# ---- small helpers --------------------------------------------------------------------------------


def _t(a):
    return th.as_tensor(np.asarray(a), dtype=th.long)


def _seg_ids(seg):
    seg = _t(seg)
    return th.repeat_interleave(th.arange(seg.numel() - 1), seg[1:] - seg[:-1])


def _seg_sum(x, ids, B):
    return th.zeros((B,) + tuple(x.shape[1:]), dtype=x.dtype).index_add(0, ids, x)


def _seg_softmax(e, ids, B):
    """DGL `segment_softmax`: max-shift, exp, divide by the segment sum."""
    idx = ids.reshape((-1,) + (1,) * (e.dim() - 1)).expand_as(e)
    mx = th.full((B,) + tuple(e.shape[1:]), float('-inf')).scatter_reduce(0, idx, e.detach(), 'amax')
    ex = th.exp(e - mx[ids])
    return ex / _seg_sum(ex, ids, B)[ids]


def _norm(x):
    return th.norm(x, p=2, dim=-1, keepdim=True)


def _gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch GRUCell arithmetic: chunks ordered r, z, n; h' = (1 - z) n + z h."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    d = h.shape[-1]
    r = th.sigmoid(gi[:, :d] + gh[:, :d])
    z = th.sigmoid(gi[:, d:2 * d] + gh[:, d:2 * d])
    n = th.tanh(gi[:, 2 * d:] + r * gh[:, 2 * d:])
    return (1 - z) * n + z * h


def renorm_rows_(weight, rows, max_norm=1.0):
    """`nn.Embedding(max_norm=1)` side effect (`msgifsr.py:162`): rows whose L2 norm exceeds max_norm are
    rescaled IN PLACE by max_norm / (norm + 1e-7), outside autograd."""
    with th.no_grad():
        rows = th.unique(rows.reshape(-1))
        n = weight[rows].norm(dim=-1)
        scale = th.where(n > max_norm, max_norm / (n + 1e-7), th.ones_like(n))
        weight[rows] = weight[rows] * scale.unsqueeze(-1)


# ---- SRGNN / NISER --------------------------------------------------------------------------------


def ggnn_layer(params, prefix, batch, x, drop=NO_DROPOUT, site=SITE_GGNN):
    """`SRGNNLayer.forward`, reference `srgnn.py:31-51` (= `niser.py:32-49`): weighted mean over in-edges
    and over out-edges (the reversed graph), two linear maps, GRUCell with the un-dropped input as state."""
    _, _, src, dst = batch['rel']['intra1']
    src, dst = _t(src), _t(dst)
    w = th.as_tensor(np.asarray(batch['w']), dtype=x.dtype)
    N = x.shape[0]
    ft = drop(x, site)
    if src.numel() == 0:
        return x

    def wmean(frm, to):
        num = th.zeros_like(ft).index_add(0, to, ft[frm] * w.unsqueeze(-1))
        den = th.zeros(N, dtype=x.dtype).index_add(0, to, w)
        return th.where(den.unsqueeze(-1) > 0, num / den.clamp(min=1).unsqueeze(-1), th.zeros_like(num))

    n_in = wmean(src, dst) @ params[prefix + 'W1.weight'].t()
    n_out = wmean(dst, src) @ params[prefix + 'W2.weight'].t()
    return _gru_cell(th.cat([n_in, n_out], 1), x, params[prefix + 'gru.weight_ih'], params[prefix + 'gru.weight_hh'],
                     params[prefix + 'gru.bias_ih'], params[prefix + 'gru.bias_hh'])


def _readout_single(params, f, last, ids, B):
    """`AttnReadout.forward`, reference `srgnn.py:76-91` (after its feat_drop)."""
    u = f @ params['readout.fc_u.weight'].t()
    v = f[last] @ params['readout.fc_v.weight'].t() + params['readout.fc_v.bias']
    e = th.sigmoid(u + v[ids]) @ params['readout.fc_e.weight'].t()
    alpha = _seg_softmax(e, ids, B)
    return _seg_sum(f * alpha, ids, B)


def srgnn_forward(params, batch, drop=NO_DROPOUT, num_layers=1, niser=False, scale=SCALE, return_aux=False):
    """`SRGNN.forward` (`srgnn.py:131-148`) and, with niser=True, `NISER.forward` (`niser.py:130-157`).
    Returns (B, V) log-probabilities.  The GGNN layers are evaluated and discarded exactly like the
    reference (their output never reaches the score)."""
    E = params['embedding.weight']
    iid, last, B = _t(batch['iid'][1]), _t(batch['last'][1]), batch['B']
    ids = _seg_ids(batch['seg'][1])
    feat = drop(E[iid], SITE_EMBED + 1)
    if niser:
        feat = feat / (_norm(feat) + 1e-12)
    out = feat
    for l in range(num_layers):
        out = ggnn_layer(params, f'layers.{l}.', batch, out, drop, SITE_GGNN + l)
    if niser:
        feat = feat / _norm(feat)
    sr_g = _readout_single(params, drop(feat, SITE_READOUT), last, ids, B)
    sr = th.cat([feat[last], sr_g], 1) @ params['fc_sr.weight'].t()
    target = E
    if niser:
        sr = sr / (_norm(sr) + 1e-12)
        target = E / (_norm(E) + 1e-12)
    logits = sr @ target.t()
    if niser and scale:
        logits = scale * logits
    res = th.log(th.softmax(logits, -1))
    return (res, dict(sr=sr, feat=feat, ggnn=out)) if return_aux else res


# ---- MSGIFSR --------------------------------------------------------------------------------------


def gat_conv(params, prefix, h_src_in, h_dst_in, src, dst, drop, slot):
    """`GATConv.forward` as driven by `MSHGNN` (`gatconv.py:254-319`, `msgifsr.py:58-64,74-75`): shared fc
    for source and destination copies (each with its own feat_drop mask), u_add_v -> LeakyReLU ->
    edge softmax over in-edges -> attn_drop -> weighted sum, Identity residual on the destination copy,
    bias.  Returns [N_dst, 8, d]."""
    W = params[prefix + 'fc.weight']
    d = W.shape[1]
    h_src = drop(h_src_in, SITE_GAT_SRC + 4 * slot)
    h_dst = drop(h_dst_in, SITE_GAT_DST + 4 * slot)
    z_src = (h_src @ W.t()).view(-1, HEADS, d)
    z_dst = (h_dst @ W.t()).view(-1, HEADS, d)
    el = (z_src * params[prefix + 'attn_l']).sum(-1)
    er = (z_dst * params[prefix + 'attn_r']).sum(-1)
    e = TF.leaky_relu(el[src] + er[dst], NEG_SLOPE)                       # [M, 8]
    n_dst = h_dst.shape[0]
    idx = dst.unsqueeze(-1).expand_as(e)
    mx = th.full((n_dst, HEADS), float('-inf')).scatter_reduce(0, idx, e.detach(), 'amax')
    ex = th.exp(e - mx[dst])
    a = ex / th.zeros(n_dst, HEADS).index_add(0, dst, ex)[dst]
    a = drop(a.unsqueeze(-1), SITE_GAT_ATTN + 4 * slot)                   # [M, 8, 1]
    rst = th.zeros(n_dst, HEADS, d).index_add(0, dst, z_src[src] * a)
    return rst + h_dst.view(n_dst, 1, d) + params[prefix + 'bias'].view(1, HEADS, d)


def _relations(batch, reverse):
    """Canonical (src_type, etype, dst_type) triples in DGL's sorted order with their COO arrays; the
    reversed graph flips every edge and turns (s, e, t) into (t, e, s) (`msgifsr.py:75`)."""
    rels = []
    for name, (st, dt, src, dst) in batch['rel'].items():
        et = 'inter' if name.startswith('inter') else name
        if reverse:
            rels.append((f's{dt}', et, f's{st}', dt, st, _t(dst), _t(src)))
        else:
            rels.append((f's{st}', et, f's{dt}', st, dt, _t(src), _t(dst)))
    return sorted(rels, key=lambda r: r[:3])


def gat_slot(layer, conv, etype, st, dt, K):
    """Dense id of one (layer, conv, relation instance) for dropout sites; shared with the product."""
    if etype == 'inter':
        r = K + (dt - 2 if st == 1 else (K - 1) + st - 2)
    else:
        r = st - 1
    return (layer * 2 + conv) * (3 * K) + r


def mshgnn_layer(params, l, batch, feats, drop=NO_DROPOUT):
    """`MSHGNN.forward` (`msgifsr.py:70-91`): HeteroGraphConv(sum) of GATConvs on the graph and on its
    reverse, max over heads, plus the broadcast per-session mean of the layer input."""
    K, B = batch['K'], batch['B']
    hs = []
    for conv in (0, 1):
        outs = {}
        for (_, et, _, st, dt, src, dst) in _relations(batch, reverse=bool(conv)):
            if src.numel() == 0:                     # HeteroGraphConv skips relations without edges
                continue
            o = gat_conv(params, f'layers.{l}.conv{conv + 1}.mods.{et}.', feats[st], feats[dt], src, dst, drop,
                         gat_slot(l, conv, et, st, dt, K))
            outs.setdefault(dt, []).append(o)
        hs.append({k: th.stack(v, 0).sum(0) for k, v in outs.items()})
    h = {}
    for k in range(1, K + 1):
        d = feats[k].shape[-1]
        acc = hs[0].get(k, th.zeros(1, d)) + hs[1].get(k, th.zeros(1, d))
        if acc.dim() > 2:
            acc = acc.max(1)[0]
        ids = _seg_ids(batch['seg'][k])
        cnt = (_t(batch['seg'][k])[1:] - _t(batch['seg'][k])[:-1]).clamp(min=1).to(feats[k].dtype)
        mean = _seg_sum(feats[k], ids, B) / cnt.unsqueeze(-1)
        h[k] = mean[ids] + acc
    return h


def _expander(params, k, feat):
    """`SemanticExpander.forward` with reducer='mean' (`msgifsr.py:32-45`): identity for items, else
    0.5 * mean over the k items + 0.5 * final hidden state of GRUs[k-2]."""
    if k == 1:
        return feat
    p = f'expander.GRUs.{k - 2}.'
    h = th.zeros(feat.shape[0], feat.shape[2])
    for t in range(k):
        h = _gru_cell(feat[:, t], h, params[p + 'weight_ih_l0'], params[p + 'weight_hh_l0'],
                      params[p + 'bias_ih_l0'], params[p + 'bias_hh_l0'])
    return 0.5 * feat.mean(1) + 0.5 * h


def mixed_rows(batch):
    """Row order of the MSGIFSR readout (`msgifsr.py:127-138`): per session, its s1 nodes, then its s2
    nodes, ...  Returns (order_of_row[R], node_of_row[R], seg[B+1]) as int64 numpy arrays."""
    K, B = batch['K'], batch['B']
    segs = [np.asarray(batch['seg'][k], np.int64) for k in range(1, K + 1)]
    kk, nn_ = [], []
    for b in range(B):
        for k in range(K):
            n = np.arange(segs[k][b], segs[k][b + 1])
            kk.append(np.full(n.shape, k + 1, np.int64))
            nn_.append(n)
    tot = np.zeros(B + 1, np.int64)
    np.cumsum(sum((s[1:] - s[:-1]) for s in segs), out=tot[1:])
    return np.concatenate(kk), np.concatenate(nn_), tot


def msgifsr_forward(params, batch, drop=NO_DROPOUT, num_layers=1, fusion=False, norm=True, return_aux=False, extra=False):
    """`MSGIFSR.forward` (`msgifsr.py:241-323`), with or without the REnorm head (`extra`, `:281-305`).  Mutates
    params['embeddings.weight'] in place exactly like `nn.Embedding(max_norm=1)` does (touched rows at the
    gather, all rows at the scoring head)."""
    E = params['embeddings.weight']
    K, B = batch['K'], batch['B']
    feats = {}
    for k in range(1, K + 1):
        iid = _t(batch['iid'][k])
        renorm_rows_(E, iid)
        f = _expander(params, k, drop(E[iid], SITE_EMBED + k))
        f = f.masked_fill(f != f, 0)
        feats[k] = TF.normalize(f, dim=-1) if norm else f
    h = feats
    for l in range(num_layers):
        h = mshgnn_layer(params, l, batch, h, drop)
    if norm:
        h = {k: TF.normalize(v, dim=-1) for k, v in h.items()}
    lasts = {k: _t(batch['last'][k]) for k in h}
    row_k, row_n, tot = mixed_rows(batch)
    rows = th.zeros(len(row_k), E.shape[1])
    for k in range(1, K + 1):
        sel = th.from_numpy(np.nonzero(row_k == k)[0])
        rows = rows.index_copy(0, sel, h[k][th.from_numpy(row_n[row_k == k])])
    ids = _seg_ids(tot)
    srs = []
    for k in range(1, K + 1):
        i = k - 1
        u = rows @ params[f'readout.fc_u.{i}.weight'].t() + params[f'readout.fc_u.{i}.bias']
        v = h[k][lasts[k]] @ params[f'readout.fc_v.{i}.weight'].t()
        e = th.sigmoid(u + v[ids]) @ params[f'readout.fc_e.{i}.weight'].t()
        g = _seg_sum(rows * _seg_softmax(e, ids, B), ids, B)
        srs.append(th.cat([h[k][lasts[k]], g], 1) @ params[f'fc_sr.{i}.weight'].t())
    sr = th.stack(srs, 1)                                                # [B, K, d]
    if norm:
        sr = TF.normalize(sr, dim=-1)
    renorm_rows_(E, th.arange(E.shape[0]))
    target = TF.normalize(E, dim=-1) if norm else E
    logits = sr @ target.t()                                             # [B, K, V]
    if extra:
        # REnorm (`msgifsr.py:281-305`): items of the session (its order-1 nodes) and all other items get their own
        # soft-max; a 2-way gate phi = sc_sr[0](sr) mixes the two distributions.
        hid = th.relu(sr @ params['sc_sr.0.0.weight'].t() + params['sc_sr.0.0.bias'])
        phi = th.softmax(hid @ params['sc_sr.0.2.weight'].t(), -1)       # [B, K, 2]
        seg1, iid1 = batch['seg'][1], _t(batch['iid'][1])
        mask = th.zeros(B, E.shape[0], dtype=th.bool)
        for b in range(B):
            mask[b, iid1[seg1[b]:seg1[b + 1]]] = True
        m3 = mask.unsqueeze(1)
        s_in = th.softmax(SCALE * logits.masked_fill(~m3, float('-inf')), -1)
        s_ex = th.softmax(SCALE * logits.masked_fill(m3, float('-inf')), -1)
        s_ex = s_ex.masked_fill(s_ex != s_ex, 0)                         # a session that holds the whole catalog
        score = phi[..., 0:1] * s_in + phi[..., 1:2] * s_ex
    else:
        score = th.softmax(SCALE * logits, -1)
    if K > 1 and fusion:
        score = (score * th.softmax(params['alpha'], -1).view(1, K, 1)).sum(1)
    else:
        score = score[:, 0]
    res = th.log(score)
    return (res, dict(sr=sr, feats=feats, h=h)) if return_aux else res


# ---- training-step helpers (the body of `TrainRunner.train`, `utils/train.py:95-101`) --------------


def nll(logp, labels):
    return TF.nll_loss(logp, _t(labels))


def decay_split(names):
    """`fix_weight_decay` (`utils/train.py:12-23`): names containing bias / batch_norm / activation get
    no L2 term."""
    no = [n for n in names if any(t in n for t in ('bias', 'batch_norm', 'activation'))]
    return [n for n in names if n not in no], no


def make_adam(params, lr=1e-3, weight_decay=1e-4):
    names = [n for n, p in params.items() if p.requires_grad]
    if weight_decay > 0:
        dec, no = decay_split(names)
        groups = [{'params': [params[n] for n in dec]}, {'params': [params[n] for n in no], 'weight_decay': 0}]
    else:
        groups = [params[n] for n in names]
    return th.optim.Adam(groups, lr=lr, weight_decay=weight_decay)


def topk_metrics(logp, labels, cutoff=20):
    """`evaluate` (`utils/train.py:36-55`) for one batch: (sum of reciprocal ranks, hits)."""
    topk = logp.topk(k=cutoff)[1]
    ranks = th.where(topk == _t(labels).unsqueeze(-1))[1] + 1
    return ranks.float().reciprocal().sum().item(), ranks.numel()
