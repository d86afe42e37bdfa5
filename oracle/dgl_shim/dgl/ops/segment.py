import torch as th


def _seg_ids(seglen):
    seglen = seglen.long()
    return th.repeat_interleave(th.arange(seglen.numel(), device=seglen.device), seglen)


def segment_reduce(seglen, value, reducer='sum'):
    """Reduce rows of `value` over contiguous segments; mean divides by max(len, 1); empty max -> 0."""
    seglen = seglen.to(value.device)
    ids = _seg_ids(seglen)
    B = seglen.numel()
    shape = (B,) + tuple(value.shape[1:])
    if reducer in ('sum', 'mean'):
        out = th.zeros(shape, dtype=value.dtype, device=value.device).index_add(0, ids, value)
        if reducer == 'mean':
            n = seglen.clamp(min=1).to(value.dtype)
            out = out / n.reshape((B,) + (1,) * (value.dim() - 1))
        return out
    if reducer in ('max', 'min'):
        idx = ids.reshape((-1,) + (1,) * (value.dim() - 1)).expand_as(value)
        out = th.zeros(shape, dtype=value.dtype, device=value.device)
        return out.scatter_reduce(0, idx, value, 'amax' if reducer == 'max' else 'amin', include_self=False)
    raise NotImplementedError(reducer)


def segment_softmax(seglen, value):
    seglen = seglen.to(value.device)
    ids = _seg_ids(seglen)
    vmax = segment_reduce(seglen, value, 'max').detach()
    ex = th.exp(value - vmax[ids])
    return ex / segment_reduce(seglen, ex, 'sum')[ids]
