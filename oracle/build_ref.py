"""Recipe for `oracle/_ref/`: the UNMODIFIED reference's Python sources for the hot path, taken from where they lie under
`/root/reference` into the git-ignored (NOT gpurun-ignored) directory `oracle/_ref/`, so that the reference's own model,
collate and training-loop code can be timed on the GPU box's host cores (`bench.py --impl reference`, `cpu_baseline.kind =
"reference"`) - `/root/reference` does not exist there.

TEST / MEASUREMENT INFRASTRUCTURE: nothing in the product package imports this.  Nothing under `oracle/_ref/` is ever
committed (`.gitignore`); `__graft_entry__.build()` calls `populate()` whenever `/root/reference` is present, which is in the
build container only.  The reference is a script tree without `setup.py` / `pyproject.toml`, so there is nothing to `pip
install`: the files are copied byte for byte (a manifest with their SHA-256 is written next to them) and imported over
`oracle/dgl_shim`, the pure-PyTorch stand-in for the un-installable DGL 0.7.2 (`environment.yaml:14-15`)."""
import hashlib
import json
import shutil
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path('/root/reference')
DST = HERE / '_ref'
FILES = ['src/models/__init__.py', 'src/models/srgnn.py', 'src/models/niser.py', 'src/models/msgifsr.py', 'src/models/lessr.py',
         'src/models/gnn_models/gatconv.py', 'src/utils/train.py', 'src/utils/data/collate.py', 'src/utils/data/dataset.py']
OPTIONAL = ['src/__init__.py', 'src/models/gnn_models/__init__.py', 'src/utils/__init__.py', 'src/utils/data/__init__.py',
            'LICENSE']


def populate():
    """Copy the reference's hot-path sources into oracle/_ref/ (no-op without /root/reference).  Returns the manifest."""
    if not (REF / 'src' / 'models' / 'srgnn.py').is_file():
        return None
    man = {}
    for rel in FILES + OPTIONAL:
        src = REF / rel
        if not src.is_file():
            if rel in FILES:
                raise FileNotFoundError(src)
            continue
        dst = DST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        data = src.read_bytes()
        if not dst.exists() or dst.read_bytes() != data:
            dst.write_bytes(data)
        man[rel] = hashlib.sha256(data).hexdigest()
    sub = REF / '.SUBMODULES.json'
    commit = None
    if sub.is_file():
        try:
            commit = json.loads(sub.read_text()).get('commit')
        except Exception:                         # noqa: BLE001
            commit = None
    (DST / 'MANIFEST.json').write_text(json.dumps(dict(source=str(REF), commit=commit, sha256=man), indent=1))
    return man


def available():
    return (DST / 'src' / 'models' / 'srgnn.py').is_file()


if __name__ == '__main__':
    m = populate()
    print('oracle/_ref:', 'reference not present, nothing copied' if m is None else f'{len(m)} files')
