"""`dgl.nn.pytorch`: HeteroGraphConv restated from DGL 0.7.2 (skips relations with no edges, sums
per-destination-type outputs with stack(...).sum(0), flips allow_zero_in_degree on its modules)."""
import torch as th
from torch import nn

from . import utils  # noqa: F401


class HeteroGraphConv(nn.Module):
    def __init__(self, mods, aggregate='sum'):
        super().__init__()
        self.mods = nn.ModuleDict(mods)
        for _, v in self.mods.items():
            fn_ = getattr(v, 'set_allow_zero_in_degree', None)
            if callable(fn_):
                fn_(True)
        assert aggregate == 'sum'

    def forward(self, g, inputs, mod_args=None, mod_kwargs=None):
        outputs = {nty: [] for nty in g.dsttypes}
        if isinstance(inputs, tuple):
            src_inputs, dst_inputs = inputs
        else:
            src_inputs = dst_inputs = inputs
        for stype, etype, dtype in g.canonical_etypes:
            rel_graph = g[stype, etype, dtype]
            if rel_graph.number_of_edges() == 0:
                continue
            if stype not in src_inputs or dtype not in dst_inputs:
                continue
            outputs[dtype].append(self.mods[etype](rel_graph, (src_inputs[stype], dst_inputs[dtype])))
        rsts = {}
        for nty, alist in outputs.items():
            if len(alist) != 0:
                rsts[nty] = th.stack(alist, 0).sum(0)
        return rsts
