"""Drop-in for `src/utils/data/dataset.py` (`:6-50`): session files, all-prefix augmentation.  Plain-Python reader
(the reference's pandas `squeeze=True` call no longer exists in pandas >= 2)."""
import numpy as np


def create_index(sessions):
    """(session id, label position) for every prefix of length >= 1 (`dataset.py:6-13`)."""
    lens = np.fromiter(map(len, sessions), dtype=np.int64)
    session_idx = np.repeat(np.arange(len(sessions)), lens - 1)
    label_idx = np.concatenate([np.arange(1, l) for l in lens]) if len(lens) else np.zeros(0, np.int64)
    return np.column_stack((session_idx, label_idx))


def read_sessions(filepath):
    out = []
    with open(filepath) as f:
        for line in f:
            line = line.strip()
            if line:
                out.append(list(map(int, line.split(','))))
    return out


def read_dataset(dataset_dir):
    from pathlib import Path
    dataset_dir = Path(dataset_dir)
    train_sessions = read_sessions(dataset_dir / 'train.txt')
    test_sessions = read_sessions(dataset_dir / 'test.txt')
    with open(dataset_dir / 'num_items.txt', 'r') as f:
        num_items = int(f.readline())
    return train_sessions, test_sessions, num_items


class AugmentedDataset:
    def __init__(self, sessions, sort_by_length=False):
        self.sessions = sessions
        index = create_index(sessions)
        if sort_by_length:
            index = index[np.argsort(index[:, 1])[::-1]]
        self.index = index

    def __getitem__(self, idx):
        sid, lidx = self.index[idx]
        return self.sessions[sid][:lidx], self.sessions[sid][lidx]

    def __len__(self):
        return len(self.index)
