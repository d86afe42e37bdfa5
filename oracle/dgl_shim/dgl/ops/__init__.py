"""`dgl.ops`: contiguous-segment reductions (DGL `segment_reduce` / `segment_softmax`)."""
import torch as th

from . import segment  # noqa: F401
from .segment import segment_reduce, segment_softmax  # noqa: F401
