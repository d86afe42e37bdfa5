"""Import the UNMODIFIED reference (`/root/reference/src/...`) on top of `oracle/dgl_shim`.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  `/root/reference` exists in the build container only; on the GPU box
the one caller is the reference arm of `bench.py` (`--impl reference` / `cpu_baseline`), which times the reference's
own code from the byte-identical copy under the git-ignored `oracle/_ref/` (`oracle/build_ref.py`).  Nothing under
`tests/ -m gpu`, `__graft_entry__.smoke()` or the product package calls this.  It exists to (1) validate the
restatement in `oracle/models.py` / `oracle/collate.py`, (2) generate the golden vectors committed under
`tests/golden/` (`oracle/make_golden.py`, `oracle/make_convergence_golden.py`) and (3) be the CPU baseline.
"""
import os
import sys
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_SHIM = _HERE / 'dgl_shim'
# /root/reference in the build container; on the GPU box the byte-identical copy of the hot-path sources that
# oracle/build_ref.py placed under the git-ignored oracle/_ref/ (bench.py's reference arm only)
REFERENCE_ROOT = Path(os.environ.get('SESSREC_REFERENCE_ROOT', '/root/reference'))
if not (REFERENCE_ROOT / 'src' / 'models' / 'srgnn.py').is_file() and (_HERE / '_ref' / 'src' / 'models' / 'srgnn.py').is_file():
    REFERENCE_ROOT = _HERE / '_ref'


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
