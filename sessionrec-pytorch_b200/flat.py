"""Flat parameter storage: every nn.Parameter of a model becomes a view into one contiguous fp32 buffer (and its
gradient a view into a twin buffer), so that the optimizer step is ONE fused kernel and the data-parallel
gradient exchange ONE NCCL all-reduce.  Names / shapes / state_dict keys are untouched."""
import torch

ALIGN = 64   # floats (256 B): keeps every tensor 128-bit loadable and TMA-friendly


class FlatParams:
    def __init__(self, module):
        named = [(n, p) for n, p in module.named_parameters()]
        assert named, 'module has no parameters'
        dev = named[0][1].device
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.offsets, off = [], 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self.data = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            self.data[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.data[o:o + p.numel()].view(p.shape)
        self.grad = torch.zeros_like(self.data)
        self.index = {n: i for i, n in enumerate(self.names)}
        # where every parameter hangs in the module tree: lets valid() notice swapped Parameter objects in O(1)
        owner = {}
        for sub in module.modules():
            for key, p in sub._parameters.items():
                if p is not None:
                    owner.setdefault(id(p), (sub, key))
        self._owner = [owner[id(p)] for p in self.params]

    def valid(self, module=None):
        """The parameters still ARE the views into the flat buffer: same Parameter objects (load_state_dict(assign=True)
        swaps them), same storage (model.to() / p.data = ... re-point them).  Checked on the first, middle and last parameter."""
        base = self.data.data_ptr()
        n = len(self.params)
        probe = sorted({0, n // 2, n - 1})
        for i in probe:
            if self.params[i].data_ptr() != base + 4 * self.offsets[i]:
                return False
        for i in probe:
            sub, key = self._owner[i]
            if sub._parameters.get(key) is not self.params[i]:
                return False
        return True

    def views(self, flat):
        """Per-parameter views (same order as `params`) into another flat buffer of the same layout."""
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def view(self, flat, name):
        i = self.index[name]
        p, o = self.params[i], self.offsets[i]
        return flat[o:o + p.numel()].view(p.shape)

    def decay_segments(self, weight_decay, inactive=(), owned_rows=None):
        """(seg_off int64[S+1], seg_decay fp32[S]) on the device: `fix_weight_decay` semantics of the reference
        (`src/utils/train.py:12-23`): names containing bias / batch_norm / activation get no L2 term.  Names in `inactive`
        get -1: the Adam kernels leave such a segment untouched, like torch.optim.Adam skips a parameter whose grad is None.
        owned_rows = (name, lo, hi): catalog sharding - of that [V, d] table only rows [lo, hi) are this rank's to update;
        the table becomes three segments with the outer two inactive (their owners broadcast the updated rows)."""
        offs, dec = [], []
        for n, p, o in zip(self.names, self.params, self.offsets):
            dcy = -1.0 if n in inactive else (0.0 if any(t in n for t in ('bias', 'batch_norm', 'activation')) else float(weight_decay))
            if owned_rows is not None and n == owned_rows[0]:
                lo, hi, d = owned_rows[1], owned_rows[2], p.shape[1]
                offs += [o, o + lo * d, o + hi * d]
                dec += [-1.0, dcy, -1.0]
            else:
                offs.append(o)
                dec.append(dcy)
        offs.append(self.total)
        dev = self.data.device
        return (torch.tensor(offs, dtype=torch.int64, device=dev), torch.tensor(dec, dtype=torch.float32, device=dev))
