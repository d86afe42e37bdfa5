import torch as th


def edge_softmax(graph, logits, eids=None, norm_by='dst'):
    """Softmax of `logits` over the in-edges of every destination node, per trailing dim."""
    (s, e, t), = graph._rels.keys()
    dst = graph._rels[(s, e, t)][1]
    n = graph._num_nodes[t]
    idx = dst.reshape((-1,) + (1,) * (logits.dim() - 1)).expand_as(logits)
    shape = (n,) + tuple(logits.shape[1:])
    mx = th.full(shape, float('-inf'), dtype=logits.dtype, device=logits.device)
    mx = mx.scatter_reduce(0, idx, logits.detach(), 'amax', include_self=True)
    ex = th.exp(logits - mx[dst])
    den = th.zeros(shape, dtype=logits.dtype, device=logits.device).index_add(0, dst, ex)
    return ex / den[dst]
