"""CPU restatement of the reference's batch collation (session -> graph arrays -> batched arrays).

TEST INFRASTRUCTURE ONLY: the checker for the native batch builder and the input generator for the
torch oracle in `oracle/models.py`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU
baseline may import this; the product package never does.

Pinned against the reference's own `collate.py` (run unmodified over `oracle/dgl_shim`) by
`tests/golden/collate_golden.json` (`oracle/make_golden.py`, `tests/test_oracle_collate.py`).

Output layout ("flat batch", all ids batch-global, numpy int64):
  B, labels[B], order K, kind in {'session', 'ccs'}
  per node type k = 1..K:  iid[k] [N_k] (k == 1) or [N_k, k];  seg[k] [B+1] node offsets;  last[k] [B]
  relations: name -> (src_type, dst_type, src[M], dst[M])   names 'intra{k}', 'inter1_{k}', 'inter{k}_1'
  kind == 'session': w [M] multiplicities of relation 'intra1'
"""
import numpy as np


def _dedup_pairs(pairs):
    """First-occurrence-ordered unique pairs with counts (`collections.Counter` semantics,
    reference `collate.py:66-69`)."""
    seen, out, cnt = {}, [], []
    for p in pairs:
        j = seen.get(p)
        if j is None:
            seen[p] = len(out)
            out.append(p)
            cnt.append(1)
        else:
            cnt[j] += 1
    return out, cnt


def session_graph(seq):
    """`seq_to_session_graph`, reference `src/utils/data/collate.py:61-85`."""
    items = sorted(set(seq))
    nid = {it: i for i, it in enumerate(items)}
    s = [nid[it] for it in seq]
    pairs, cnt = _dedup_pairs(list(zip(s[:-1], s[1:])))
    if not pairs:                       # single click: self-loop with weight 1 (`collate.py:74-76`)
        pairs, cnt = [(0, 0)], [1]
    return dict(iid=[np.asarray(items, np.int64)], last=[s[-1]],
                rel={'intra1': (1, 1, [p[0] for p in pairs], [p[1] for p in pairs])}, w=cnt)


def ccs_graph(seq, order):
    """`seq_to_ccs_graph`, reference `src/utils/data/collate.py:87-217`: item nodes (s1) plus one node
    per distinct consecutive k-gram (s_k, k = 2..order), intra-order and inter-order relations."""
    L = len(seq)
    items = sorted(set(seq))
    nid = {it: i for i, it in enumerate(items)}
    s = [nid[it] for it in seq]
    iid = [np.asarray(items, np.int64)]
    last = [s[-1]]
    gid = [s]                                            # gid[k-1][j] = node id of the k-gram starting at j
    for k in range(2, order + 1):
        table, ids, rows = {}, [], []
        for j in range(L - k + 1):
            key = tuple(seq[j:j + k])
            if key not in table:
                table[key] = len(rows)
                rows.append(list(key))
            ids.append(table[key])
        gid.append(ids)
        if rows:
            iid.append(np.asarray(rows, np.int64))
            last.append(ids[-1])
        else:                                           # session shorter than k: one dummy node made of
            iid.append(np.full((1, k), items[0], np.int64))   # the smallest item id (`collate.py:203-207`)
            last.append(0)
    rel = {}
    for k in range(1, order + 1):
        g = gid[k - 1]
        pairs, _ = _dedup_pairs(list(zip(g[:-1], g[1:])))
        rel[f'intra{k}'] = (k, k, [p[0] for p in pairs], [p[1] for p in pairs])
    for k in range(2, order + 1):
        g = gid[k - 1]
        n = max(L - k, 0)
        fwd, _ = _dedup_pairs([(s[i], g[i + 1]) for i in range(n)])      # item preceding the gram -> gram
        bwd, _ = _dedup_pairs([(g[i], s[i + k]) for i in range(n)])      # gram -> item following it
        rel[f'inter1_{k}'] = (1, k, [p[0] for p in fwd], [p[1] for p in fwd])
        rel[f'inter{k}_1'] = (k, 1, [p[0] for p in bwd], [p[1] for p in bwd])
    return dict(iid=iid, last=last, rel=rel, w=None)


def build_batch(seqs, labels, kind='session', order=1):
    """`collate_fn_factory` / `collate_fn_factory_ccs` + `dgl.batch`, reference `collate.py:219-256`."""
    assert kind in ('session', 'ccs')
    K = 1 if kind == 'session' else order
    graphs = [session_graph(list(q)) if kind == 'session' else ccs_graph(list(q), K) for q in seqs]
    B = len(graphs)
    out = dict(B=B, K=K, kind=kind, labels=np.asarray(labels, np.int64), iid={}, seg={}, last={}, rel={})
    for k in range(1, K + 1):
        cnt = np.asarray([len(g['iid'][k - 1]) for g in graphs], np.int64)
        seg = np.zeros(B + 1, np.int64)
        np.cumsum(cnt, out=seg[1:])
        out['seg'][k] = seg
        out['iid'][k] = np.concatenate([g['iid'][k - 1] for g in graphs], 0)
        out['last'][k] = seg[:-1] + np.asarray([g['last'][k - 1] for g in graphs], np.int64)
    for name in graphs[0]['rel']:
        st, dt = graphs[0]['rel'][name][:2]
        src = [np.asarray(g['rel'][name][2], np.int64) + out['seg'][st][b] for b, g in enumerate(graphs)]
        dst = [np.asarray(g['rel'][name][3], np.int64) + out['seg'][dt][b] for b, g in enumerate(graphs)]
        out['rel'][name] = (st, dt, np.concatenate(src), np.concatenate(dst))
    if kind == 'session':
        out['w'] = np.concatenate([np.asarray(g['w'], np.int64) for g in graphs])
    return out


def create_index(sessions):
    """All-prefix augmentation index, reference `src/utils/data/dataset.py:6-13`: one sample per
    (session, label position >= 1)."""
    idx = [(sid, l) for sid, s in enumerate(sessions) for l in range(1, len(s))]
    return np.asarray(idx, np.int64).reshape(-1, 2)


def augmented_samples(sessions):
    """`AugmentedDataset.__getitem__`, reference `dataset.py:42-47`."""
    return [(list(sessions[sid][:l]), int(sessions[sid][l])) for sid, l in create_index(sessions)]


def read_sessions(path):
    """`read_sessions`, reference `dataset.py:16-19` (one comma-separated session per line)."""
    out = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                out.append([int(t) for t in line.split(',')])
    return out
