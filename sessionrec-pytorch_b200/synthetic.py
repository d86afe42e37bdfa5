"""Synthetic session batches of the shapes BASELINE.json names (no public dataset is reachable offline).
Recipe: prefix length 1 + min(Geom(0.25), 18); items ~ Zipf(1.05) over a random
permutation of the catalog; with probability 0.15 a click repeats an earlier item of the session; labels from
the same Zipf."""
import numpy as np

CONFIGS = {
    # name: model, V, d, B, order, layers, dropout  (BASELINE.json configs[0..4])
    'cfg0': dict(model='SRGNN', V=3429, d=256, B=32, order=1, layers=1, dropout=0.1, note='datasets/sample shape'),
    'cfg1': dict(model='MSGIFSR', V=43097, d=96, B=512, order=1, layers=1, dropout=0.1, note='Diginetica shape'),
    'cfg2': dict(model='SRGNN', V=17000, d=256, B=2048, order=1, layers=1, dropout=0.1, note='Yoochoose1/64 shape'),
    'cfg3': dict(model='NISER', V=29510, d=64, B=128, order=1, layers=2, dropout=0.5, note='Gowalla shape, per rank'),
    'cfg4': dict(model='MSGIFSR', V=29618, d=256, B=512, order=1, layers=1, dropout=0.1, note='Yoochoose1/4 shape'),
    # not a BASELINE config: the reference's argparse default order (main_msgifsr.py:84) at the cfg1 shape; K > 1 runs through
    # the general native step (csrc/step_k.cu: k-gram node types, intra / inter relations, ~330 launches per step)
    'cfg1k3': dict(model='MSGIFSR', V=43097, d=96, B=512, order=3, layers=1, dropout=0.1, note='Diginetica shape, order 3'),
}


class SessionSampler:
    def __init__(self, V, seed=123, zipf_s=1.05, p_repeat=0.15, p_len=0.25, max_len=19):
        self.rng = np.random.default_rng(seed)
        self.V, self.p_repeat, self.p_len, self.max_len = V, p_repeat, p_len, max_len
        w = 1.0 / np.arange(1, V + 1) ** zipf_s
        self.cdf = np.cumsum(w / w.sum())
        self.perm = self.rng.permutation(V)

    def _items(self, n):
        return self.perm[np.minimum(np.searchsorted(self.cdf, self.rng.random(n)), self.V - 1)]

    def batch(self, B):
        """(items int32[T], offs int32[B+1], labels int32[B])"""
        lens = 1 + np.minimum(self.rng.geometric(self.p_len, B) - 1, self.max_len - 1)
        offs = np.zeros(B + 1, np.int32)
        np.cumsum(lens, out=offs[1:])
        items = self._items(int(offs[-1])).astype(np.int32)
        rep = self.rng.random(items.shape[0]) < self.p_repeat
        for b in range(B):
            lo, hi = offs[b], offs[b + 1]
            for i in range(lo + 1, hi):
                if rep[i]:
                    items[i] = items[self.rng.integers(lo, i)]
        return items, offs, self._items(B).astype(np.int32)

    def sessions(self, B):
        items, offs, labels = self.batch(B)
        return [items[offs[b]:offs[b + 1]].tolist() for b in range(B)], labels.tolist()
