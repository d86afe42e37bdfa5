"""sessionrec-pytorch_b200: B200-native (sm_100a) training hot path for SRGNN / NISER+ / MSGIFSR behind the
reference's `src/models` forward()/loss contract.  The directory name is not a Python identifier; load it with
`__graft_entry__.load_package()` (module name `sessionrec_pytorch_b200`)."""
from . import _lib  # noqa: F401
from .batch import SessionBatch  # noqa: F401
from . import ops  # noqa: F401,E402
