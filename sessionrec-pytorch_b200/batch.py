"""SessionBatch: the flat int32 batch produced by the native builder (csrc/batch_builder.cu) and consumed by
the kernels.  It plays the role of the batched DGLGraph in the reference's collate -> model handshake
(`src/utils/data/collate.py:219-256`, `src/utils/train.py:26-32`): it supports `.to(device)` and
`.pin_memory()`, so `prepare_batch` works unchanged.  One contiguous buffer = one H2D copy per batch."""
import ctypes

import numpy as np
import torch

from . import _lib

_TYPE_TAB, _REL_TAB, _TAB_W, _MAXK = 16, 80, 16, 4
_DATA0 = _REL_TAB + _TAB_W * (3 * _MAXK - 2)
_MAGIC = 0x53524B31


def _rel_name(code):
    if code < 100:
        return f'intra{code}'
    return f'inter1_{code - 100}' if code < 200 else f'inter{code - 200}_1'


class SessionBatch:
    """Header fields (B, K, node / edge counts) are plain ints read from the host copy of the header; the per-section
    tensor views (`labels`, `types[k][...]`, `rels[r][...]`) are built on first use - the native training step only needs
    `buf` + `hdr`, and `.to(device)` sits inside the timed end-to-end loop."""
    _LAZY = ('labels', 'row_seg', 'row_type', 'row_node', 'types', 'rels')

    def __init__(self, buf, hdr=None):
        self.buf = buf                                            # int32 tensor (host or device)
        if hdr is None:                                           # host copy of the header words
            assert buf.numel() >= _DATA0, 'not a SessionBatch buffer'
            hdr = (np.ctypeslib.as_array((ctypes.c_int32 * _DATA0).from_address(buf.data_ptr())) if not buf.is_cuda
                   else buf[:_DATA0].cpu().numpy())
        self.hdr = np.array(hdr, dtype=np.int32)
        h = self.hdr
        assert int(h[0]) == _MAGIC, 'not a SessionBatch buffer'
        self.B, self.K, self.R = int(h[1]), int(h[3]), int(h[6])
        self.kind = 'session' if int(h[2]) == 0 else 'ccs'
        self.N1 = int(h[_TYPE_TAB])                               # nodes of type s1
        self.M1 = int(h[_REL_TAB + 2]) if int(h[5]) > 0 else 0    # edges of the first relation

    def rel_edge_counts(self):
        """{relation name: number of edges}, read from the host header only (no tensor views are built)."""
        h = self.hdr
        return {_rel_name(int(h[_REL_TAB + _TAB_W * r + 12])): int(h[_REL_TAB + _TAB_W * r + 2]) for r in range(int(h[5]))}

    def empty_relations(self):
        """Names of the relations without a single edge in this batch (hashable; cached)."""
        e = self.__dict__.get('_empty_rels')
        if e is None:
            e = self.__dict__['_empty_rels'] = frozenset(n for n, m in self.rel_edge_counts().items() if m == 0)
        return e

    def __getattr__(self, name):
        if name in SessionBatch._LAZY:
            self._build_views()
            return self.__dict__[name]
        raise AttributeError(name)

    def _build_views(self):
        h = self.hdr
        self.labels = self._sec(h[7], self.B)
        self.row_seg = self._sec(h[8], self.B + 1)
        self.row_type = self._sec(h[9], self.R)
        self.row_node = self._sec(h[10], self.R)
        self.types = {}
        for k in range(1, self.K + 1):
            t = h[_TYPE_TAB + _TAB_W * (k - 1):]
            N, U = int(t[0]), int(t[8])
            self.types[k] = dict(N=N, U=U, P=N * k, iid=self._sec(t[1], N * k), seg=self._sec(t[2], self.B + 1),
                                 last=self._sec(t[3], self.B), node2seg=self._sec(t[4], N),
                                 perm=self._sec(t[5], N * k), uoff=self._sec(t[6], U + 1), uid=self._sec(t[7], U),
                                 last_row=self._sec(t[9], self.B), row_of=self._sec(t[10], N))
        self.rels = []
        for r in range(int(h[5])):
            t = h[_REL_TAB + _TAB_W * r:]
            st, dt, M = int(t[0]), int(t[1]), int(t[2])
            Ns, Nd = self.types[st]['N'], self.types[dt]['N']
            rel = dict(st=st, dt=dt, M=M, code=int(t[12]), name=_rel_name(int(t[12])),
                       src=self._sec(t[3], M), dst=self._sec(t[4], M),
                       in_ptr=self._sec(t[5], Nd + 1), in_src=self._sec(t[6], M), in_eid=self._sec(t[7], M),
                       out_ptr=self._sec(t[8], Ns + 1), out_dst=self._sec(t[9], M), out_eid=self._sec(t[10], M),
                       w=self._sec(t[11], M).view(torch.float32) if self.kind == 'session' else None)
            self.rels.append(rel)

    def _sec(self, off, n):
        off = int(off)
        return self.buf[off:off + int(n)]

    # ---- construction ------------------------------------------------------------------------------
    @staticmethod
    def build(seqs, labels, kind='session', order=1, pin=False):
        """Native equivalent of `collate_fn` (`collate.py:219-230,232-256`)."""
        lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
        offs = np.zeros(len(seqs) + 1, np.int32)
        np.cumsum(lens, out=offs[1:])
        items = np.fromiter((int(i) for s in seqs for i in s), dtype=np.int32, count=int(offs[-1]))
        return SessionBatch.build_flat(items, offs, np.asarray(labels, np.int32), kind, order, pin)

    @staticmethod
    def build_flat(items, offs, labels, kind='session', order=1, pin=False, out=None):
        """items int32[T] (all sessions back to back), offs int32[B + 1], labels int32[B].  `out`: a host int32 tensor to
        build into (e.g. a reused pinned buffer, see loader.BatchPrefetcher); it must hold `batch_words(...)` words."""
        L = _lib.lib()
        B = len(offs) - 1
        kind_i = 0 if kind == 'session' else 1
        items = np.ascontiguousarray(items, np.int32)
        offs = np.ascontiguousarray(offs, np.int32)
        labels = np.ascontiguousarray(labels, np.int32)
        ip, op, lp = (a.__array_interface__['data'][0] for a in (items, offs, labels))      # plain ints: c_void_p arguments
        if out is None:
            cap = L.call('srk_batch_size', ip, op, B, kind_i, order)
            buf = torch.empty(cap, dtype=torch.int32, pin_memory=pin)
        else:
            if out.dtype != torch.int32 or out.is_cuda or not out.is_contiguous():
                raise _lib.SessRecError('build_flat: `out` must be a contiguous host int32 tensor')
            buf, cap = out, out.numel()
        used = L.call('srk_batch_build', ip, op, lp, B, kind_i, order, buf.data_ptr(), cap)
        return SessionBatch(buf[:used])

    @staticmethod
    def batch_words(n_items, B, kind='session', order=1):
        """Upper bound of the buffer size (int32 words) of a batch of B sessions with n_items clicks in total."""
        offs = np.zeros(B + 1, np.int32)
        offs[B] = n_items
        return int(_lib.lib().call('srk_batch_size', None, offs.__array_interface__['data'][0], B,
                                   0 if kind == 'session' else 1, order))

    # ---- the DGLGraph-like surface `prepare_batch` relies on ----------------------------------------
    def to(self, device, non_blocking=True):
        device = torch.device(device)
        if self.buf.device == device:
            return self
        return SessionBatch(self.buf.to(device, non_blocking=non_blocking), self.hdr)

    def pin_memory(self):
        return SessionBatch(self.buf.pin_memory(), self.hdr)

    @property
    def device(self):
        return self.buf.device

    @property
    def nbytes(self):
        return self.buf.numel() * 4

    def batch_num_nodes(self, k=1):
        seg = self.types[k]['seg']
        return (seg[1:] - seg[:-1]).long()

    # ---- export to the oracle's flat dict (tests only use this; cheap, host side) --------------------
    def to_flat_dict(self):
        c = SessionBatch(self.buf.cpu(), self.hdr)
        out = dict(B=c.B, K=c.K, kind=c.kind, labels=c.labels.numpy().astype(np.int64), iid={}, seg={}, last={}, rel={})
        for k, t in c.types.items():
            iid = t['iid'].numpy().astype(np.int64)
            out['iid'][k] = iid if k == 1 else iid.reshape(-1, k)
            out['seg'][k] = t['seg'].numpy().astype(np.int64)
            out['last'][k] = t['last'].numpy().astype(np.int64)
        for r in c.rels:
            out['rel'][r['name']] = (r['st'], r['dt'], r['src'].numpy().astype(np.int64), r['dst'].numpy().astype(np.int64))
        if c.kind == 'session':
            out['w'] = c.rels[0]['w'].numpy().astype(np.int64)
        return out
