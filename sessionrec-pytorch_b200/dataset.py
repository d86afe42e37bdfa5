"""Drop-in for `src/utils/data/dataset.py` (`:6-50`): session files and the all-prefix augmentation, same public names
(`create_index`, `read_sessions`, `read_dataset`, `AugmentedDataset`).

Stored flat: all clicks of all sessions in one int32 array plus session offsets, and the sample index as two parallel
arrays (session, label position).  `AugmentedDataset.flat()` hands every sample of the dataset to the native batch builder
as (items, offs, labels) without a Python loop (`loader.EpochBatches`); `__getitem__` still yields the reference's
`(prefix, label)` pairs for a `DataLoader` + `collate_fn`.  Text files are parsed directly (the reference's
`pandas.read_csv(..., squeeze=True)` no longer exists in pandas >= 2)."""
from pathlib import Path

import numpy as np


def create_index(sessions):
    """Rows (session id, label position) of every sample: session s of length L contributes label positions 1 .. L - 1, in
    session order (`dataset.py:6-13`)."""
    n_labels = np.maximum(np.fromiter((len(s) for s in sessions), dtype=np.int64, count=len(sessions)) - 1, 0)
    first = np.cumsum(n_labels) - n_labels                      # index of each session's first sample
    sid = np.repeat(np.arange(len(sessions), dtype=np.int64), n_labels)
    pos = np.arange(int(n_labels.sum()), dtype=np.int64) - first[sid] + 1
    return np.stack([sid, pos], axis=1)


def read_sessions(filepath):
    """One session per line, item ids separated by commas."""
    with open(filepath) as f:
        return [[int(tok) for tok in line.split(',')] for line in (ln.strip() for ln in f) if line]


def read_dataset(dataset_dir):
    root = Path(dataset_dir)
    num_items = int((root / 'num_items.txt').read_text().split()[0])
    return read_sessions(root / 'train.txt'), read_sessions(root / 'test.txt'), num_items


class AugmentedDataset:
    """Every proper prefix of every session is a sample whose label is the next click (`dataset.py:29-50`);
    `sort_by_length=True` orders the samples by decreasing prefix length like the reference (`:36-39`)."""

    def __init__(self, sessions, sort_by_length=False):
        self.sessions = sessions
        lens = np.fromiter((len(s) for s in sessions), dtype=np.int64, count=len(sessions))
        self._start = np.zeros(len(sessions) + 1, np.int64)
        np.cumsum(lens, out=self._start[1:])
        self._clicks = np.fromiter((int(i) for s in sessions for i in s), dtype=np.int32, count=int(self._start[-1]))
        index = create_index(sessions)
        if sort_by_length:
            index = index[np.argsort(index[:, 1])[::-1]]
        self.index = index

    def __len__(self):
        return self.index.shape[0]

    def __getitem__(self, idx):
        s, p = (int(x) for x in self.index[idx])
        lo = int(self._start[s])
        return self._clicks[lo:lo + p].tolist(), int(self._clicks[lo + p])

    def flat(self, order=None):
        """(items int32[T], offs int32[n + 1], labels int32[n]) of all samples in index order: what
        `SessionBatch.build_flat` / `loader.EpochBatches` take.  `order`: sample ids in the order a sampler would draw them
        (e.g. `torch.randperm(len(ds))` for the shuffled loader of `main_niser.py:86`)."""
        index = self.index if order is None else self.index[np.asarray(order, dtype=np.int64)]
        sid, pos = index[:, 0], index[:, 1]
        offs = np.zeros(len(pos) + 1, np.int64)
        np.cumsum(pos, out=offs[1:])
        assert offs[-1] < 2 ** 31, 'more than 2^31 clicks in one pass: build it in pieces'
        within = np.arange(int(offs[-1]), dtype=np.int64) - np.repeat(offs[:-1], pos)
        items = self._clicks[np.repeat(self._start[sid], pos) + within]
        labels = self._clicks[self._start[sid] + pos]
        return items, offs.astype(np.int32), labels.astype(np.int32)
