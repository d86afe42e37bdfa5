from . import functional  # noqa: F401
from . import pytorch  # noqa: F401
