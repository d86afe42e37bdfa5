"""Import the UNMODIFIED reference (`/root/reference/src/...`) on top of `oracle/dgl_shim`.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: `/root/reference` does not exist on
the GPU box, so nothing under `tests/ -m gpu`, `bench.py` or `__graft_entry__.smoke()` may call this.
It exists to (1) validate the restatement in `oracle/models.py` / `oracle/collate.py` and (2) generate
the golden vectors committed under `tests/golden/` (`oracle/make_golden.py`).
"""
import os
import sys
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get('SESSREC_REFERENCE_ROOT', '/root/reference'))
_SHIM = Path(__file__).resolve().parent / 'dgl_shim'


def available():
    return (REFERENCE_ROOT / 'src' / 'models' / 'srgnn.py').is_file()


def load():
    """Returns a namespace with the reference's model classes, collate fns and train helpers."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    sys.dont_write_bytecode = True          # /root/reference is read-only
    for p in (str(_SHIM), str(REFERENCE_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dgl  # noqa: F401  (the shim)
    assert 'dgl_shim' in dgl.__file__, 'a real dgl shadowed the shim: ' + dgl.__file__
    os.environ.setdefault('WANDB_MODE', 'disabled')
    from types import SimpleNamespace
    from src.models.srgnn import SRGNN, SRGNNLayer
    from src.models.niser import NISER
    from src.models.msgifsr import MSGIFSR
    from src.models.gnn_models.gatconv import GATConv
    from src.utils.data import collate
    from src.utils.data.dataset import AugmentedDataset
    from src.utils import train
    return SimpleNamespace(SRGNN=SRGNN, SRGNNLayer=SRGNNLayer, NISER=NISER, MSGIFSR=MSGIFSR, GATConv=GATConv,
                           collate=collate, AugmentedDataset=AugmentedDataset, train=train, dgl=dgl)


def read_sessions(path):
    """`src/utils/data/dataset.py:16-19` restated without pandas' removed `squeeze=` keyword."""
    out = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                out.append([int(t) for t in line.split(',')])
    return out
