"""Minimal pure-PyTorch stand-in for the ~25 DGL 0.7.2 API points the reference touches.

TEST INFRASTRUCTURE ONLY.  DGL is an un-vendored dependency of the reference
(`/root/reference/environment.yaml:14-15`, dgl 0.7.2) and is not installable in this image, so the
reference's *unmodified* model / collate code is run over this shim to produce the golden vectors in
`tests/golden/` (see `oracle/make_golden.py`).  Semantics restated from DGL 0.7.2's published
behaviour; call sites: `src/models/srgnn.py:37-41,82-86,139`, `src/models/msgifsr.py:60-87,131-146,264`,
`src/models/gnn_models/gatconv.py:256,294-304`, `src/utils/data/collate.py:41,78,197-211,225,246`.

Nothing in the product package imports this.
"""
from contextlib import contextmanager

import torch as th

from . import function  # noqa: F401
from . import base  # noqa: F401

_N, _E = '_N', '_E'


def _as_ids(x):
    if isinstance(x, th.Tensor):
        return x.long().reshape(-1)
    return th.tensor(list(x), dtype=th.long).reshape(-1)


class _EdgeBatch:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class _NodeBatch:
    def __init__(self, data, mailbox=None):
        self.data, self.mailbox = data, mailbox


class _NodeView:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        g = self._g

        class _V:
            data = g._nframes[ntype]
        return _V()


class _Frame(dict):
    """Feature dict bound to an entity count (rows are checked on assignment)."""

    def __init__(self, owner, kind, key, *a):
        super().__init__(*a)
        self._owner, self._kind, self._key = owner, kind, key

    def __setitem__(self, k, v):
        n = (self._owner._num_nodes[self._key] if self._kind == 'n'
             else self._owner._rels[self._key][0].numel())
        assert v.shape[0] == n, f'feature {k!r}: {v.shape[0]} rows for {n} entities'
        super().__setitem__(k, v)

    def update(self, other):
        for k, v in other.items():
            self[k] = v


class DGLGraph:
    """Heterograph container; a homogeneous graph is the single-type special case."""

    def __init__(self, num_nodes, rels, bnn=None, bne=None):
        self._num_nodes = dict(num_nodes)                       # ntype -> int
        self._rels = {k: (v[0], v[1]) for k, v in rels.items()}  # (s,e,t) -> (src, dst)
        self._nframes = {nt: _Frame(self, 'n', nt) for nt in self._num_nodes}
        self._eframes = {r: _Frame(self, 'e', r) for r in self._rels}
        self._bnn = bnn   # ntype -> int64[B] or None (unbatched)
        self._bne = bne

    # ---- structure -------------------------------------------------------------------------
    @property
    def ntypes(self):
        return sorted(self._num_nodes)

    @property
    def canonical_etypes(self):
        return sorted(self._rels)

    @property
    def dsttypes(self):
        return self.ntypes

    @property
    def is_block(self):
        return False

    def _only_ntype(self):
        assert len(self._num_nodes) == 1, 'ntype required on a heterograph'
        return next(iter(self._num_nodes))

    def _only_rel(self):
        assert len(self._rels) == 1, 'etype required on a heterograph'
        return next(iter(self._rels))

    def num_nodes(self, ntype=None):
        if ntype is None:
            return sum(self._num_nodes.values())
        return self._num_nodes[ntype]

    number_of_nodes = num_nodes

    def number_of_edges(self, etype=None):
        if etype is None:
            return sum(v[0].numel() for v in self._rels.values())
        return self._rels[etype][0].numel()

    num_edges = number_of_edges

    def number_of_dst_nodes(self):
        r = self._only_rel()
        return self._num_nodes[r[2]]

    def batch_num_nodes(self, ntype=None):
        nt = self._only_ntype() if ntype is None else ntype
        if self._bnn is None:
            return th.tensor([self._num_nodes[nt]], dtype=th.long)
        return self._bnn[nt]

    def in_degrees(self):
        s, e, t = self._only_rel()
        dst = self._rels[(s, e, t)][1]
        return th.bincount(dst, minlength=self._num_nodes[t])

    def add_nodes(self, k, ntype=None):
        nt = self._only_ntype() if ntype is None else ntype
        old = self._nframes[nt]
        self._num_nodes[nt] += k
        for key in list(old):
            v = dict.__getitem__(old, key)
            pad = th.zeros((k,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            dict.__setitem__(old, key, th.cat([v, pad], 0))

    def reverse(self, copy_ndata=True, copy_edata=False):
        rels = {(t, e, s): (dst, src) for (s, e, t), (src, dst) in self._rels.items()}
        g = DGLGraph(self._num_nodes, rels, self._bnn, self._bne)
        if copy_ndata:
            for nt, fr in self._nframes.items():
                dict.update(g._nframes[nt], fr)
        if copy_edata:
            for (s, e, t), fr in self._eframes.items():
                dict.update(g._eframes[(t, e, s)], fr)
        return g

    def __getitem__(self, key):
        s, e, t = key
        src, dst = self._rels[(s, e, t)]
        g = DGLGraph.__new__(DGLGraph)
        g._num_nodes = {s: self._num_nodes[s], t: self._num_nodes[t]}
        g._rels = {(s, e, t): (src, dst)}
        g._nframes = {s: self._nframes[s], t: self._nframes[t]}  # shared with the parent, like DGL
        g._eframes = {(s, e, t): self._eframes[(s, e, t)]}
        g._bnn, g._bne = self._bnn, self._bne
        return g

    def to(self, device, **kw):
        g = DGLGraph(self._num_nodes,
                     {k: (v[0].to(device), v[1].to(device)) for k, v in self._rels.items()},
                     None if self._bnn is None else {k: v.to(device) for k, v in self._bnn.items()},
                     self._bne)
        for nt, fr in self._nframes.items():
            dict.update(g._nframes[nt], {k: v.to(device) for k, v in fr.items()})
        for r, fr in self._eframes.items():
            dict.update(g._eframes[r], {k: v.to(device) for k, v in fr.items()})
        return g

    def pin_memory(self):
        return self

    # ---- data views ------------------------------------------------------------------------
    @property
    def nodes(self):
        return _NodeView(self)

    @property
    def ndata(self):
        return self._nframes[self._only_ntype()]

    @property
    def edata(self):
        return self._eframes[self._only_rel()]

    @property
    def srcdata(self):
        return self._nframes[self._only_rel()[0]]

    @property
    def dstdata(self):
        return self._nframes[self._only_rel()[2]]

    @contextmanager
    def local_scope(self):
        saved_n = {k: dict(v) for k, v in self._nframes.items()}
        saved_e = {k: dict(v) for k, v in self._eframes.items()}
        try:
            yield
        finally:
            for k, v in saved_n.items():
                self._nframes[k].clear()
                dict.update(self._nframes[k], v)
            for k, v in saved_e.items():
                self._eframes[k].clear()
                dict.update(self._eframes[k], v)

    def filter_nodes(self, predicate, ntype=None):
        nt = self._only_ntype() if ntype is None else ntype
        mask = predicate(_NodeBatch(self._nframes[nt]))
        return th.nonzero(mask.reshape(-1), as_tuple=False).reshape(-1)

    # ---- message passing -------------------------------------------------------------------
    def apply_edges(self, func):
        s, e, t = self._only_rel()
        src, dst = self._rels[(s, e, t)]
        assert isinstance(func, function._Binary) and func.name == 'u_add_v'
        self._eframes[(s, e, t)][func.out] = self._nframes[s][func.lhs][src] + self._nframes[t][func.rhs][dst]

    def update_all(self, message_func, reduce_func):
        s, e, t = self._only_rel()
        src, dst = self._rels[(s, e, t)]
        n_dst = self._num_nodes[t]
        sdata, ddata, edata = self._nframes[s], self._nframes[t], self._eframes[(s, e, t)]
        if isinstance(message_func, function._Binary):
            assert isinstance(reduce_func, function._Reduce) and reduce_func.name == 'sum'
            if message_func.name == 'u_mul_e':
                m = sdata[message_func.lhs][src] * edata[message_func.rhs]
            elif message_func.name == 'copy_u':
                m = sdata[message_func.lhs][src]
            else:
                raise NotImplementedError(message_func.name)
            out = th.zeros((n_dst,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
            ddata[reduce_func.out] = out.index_add(0, dst, m)
            return
        # UDF path = DGL degree bucketing: mailbox (n_bucket, deg, ...) ordered by edge id,
        # zero-in-degree nodes keep zeros.
        msgs = message_func(_EdgeBatch({k: v[src] for k, v in sdata.items()},
                                       {k: v[dst] for k, v in ddata.items()}, dict(edata)))
        deg = th.bincount(dst, minlength=n_dst)
        order = th.argsort(dst, stable=True)           # edges grouped by dst, edge-id order inside
        start = th.cumsum(deg, 0) - deg
        results = {}
        for d in th.unique(deg).tolist():
            if d == 0:
                continue
            nodes = th.nonzero(deg == d, as_tuple=False).reshape(-1)
            eidx = order[(start[nodes].unsqueeze(1) + th.arange(d).unsqueeze(0)).reshape(-1)]
            mailbox = {k: v[eidx].reshape((nodes.numel(), d) + tuple(v.shape[1:])) for k, v in msgs.items()}
            red = reduce_func(_NodeBatch({k: v[nodes] for k, v in ddata.items()}, mailbox))
            for k, v in red.items():
                if k not in results:
                    results[k] = th.zeros((n_dst,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                results[k] = results[k].index_copy(0, nodes, v)
        for k, v in results.items():
            ddata[k] = v


def graph(data, num_nodes=None, **kw):
    src, dst = _as_ids(data[0]), _as_ids(data[1])
    if num_nodes is None:
        num_nodes = int(max(src.max().item(), dst.max().item())) + 1 if src.numel() else 0
    return DGLGraph({_N: num_nodes}, {(_N, _E, _N): (src, dst)})


def heterograph(data_dict, num_nodes_dict=None, **kw):
    rels, nn_ = {}, {}
    for (s, e, t), (src, dst) in data_dict.items():
        src, dst = _as_ids(src), _as_ids(dst)
        rels[(s, e, t)] = (src, dst)
        nn_[s] = max(nn_.get(s, 0), int(src.max().item()) + 1 if src.numel() else 0)
        nn_[t] = max(nn_.get(t, 0), int(dst.max().item()) + 1 if dst.numel() else 0)
    if num_nodes_dict:
        nn_.update(num_nodes_dict)
    return DGLGraph(nn_, rels)


def batch(graphs):
    g0 = graphs[0]
    ntypes, rels = g0.ntypes, g0.canonical_etypes
    for g in graphs:
        assert g.ntypes == ntypes and g.canonical_etypes == rels
    counts = {nt: th.tensor([g._num_nodes[nt] for g in graphs], dtype=th.long) for nt in ntypes}
    offs = {nt: th.cumsum(counts[nt], 0) - counts[nt] for nt in ntypes}
    brels = {}
    for (s, e, t) in rels:
        srcs = [g._rels[(s, e, t)][0] + offs[s][i] for i, g in enumerate(graphs)]
        dsts = [g._rels[(s, e, t)][1] + offs[t][i] for i, g in enumerate(graphs)]
        brels[(s, e, t)] = (th.cat(srcs), th.cat(dsts))
    bg = DGLGraph({nt: int(counts[nt].sum()) for nt in ntypes}, brels, counts, None)
    for nt in ntypes:
        keys = set()
        for g in graphs:
            if g._num_nodes[nt] > 0:
                keys |= set(g._nframes[nt])
        for k in keys:
            parts = [g._nframes[nt][k] for g in graphs if g._num_nodes[nt] > 0]
            dict.__setitem__(bg._nframes[nt], k, th.cat(parts, 0))
    for r in rels:
        keys = set()
        for g in graphs:
            keys |= set(g._eframes[r])
        for k in keys:
            parts = [g._eframes[r][k] for g in graphs if k in g._eframes[r]]
            dict.__setitem__(bg._eframes[r], k, th.cat(parts, 0))
    return bg


def broadcast_nodes(g, feat, ntype=None):
    return th.repeat_interleave(feat, g.batch_num_nodes(ntype).to(feat.device), dim=0)


from . import ops  # noqa: E402,F401
from . import utils  # noqa: E402,F401
from . import nn  # noqa: E402,F401
