"""Drop-in for the names the reference's scripts import from `src/utils/data/collate.py` (`:61-85,87-217,219-256`).
The per-session graph functions are markers only: the whole batch is built natively in one call
(`csrc/batch_builder.cu`), which is what replaces `list(map(seq_to_graph, seqs))` + `dgl.batch`."""
import torch

from .batch import SessionBatch


def seq_to_session_graph(seq):
    raise TypeError('pass seq_to_session_graph to collate_fn_factory(); graphs are built per batch, not per session')


def seq_to_ccs_graph(seq, order=1, coaDict=None):
    raise TypeError('pass seq_to_ccs_graph to collate_fn_factory_ccs(); graphs are built per batch, not per session')


def _kind(fn):
    if fn is seq_to_session_graph:
        return 'session'
    if fn is seq_to_ccs_graph:
        return 'ccs'
    raise ValueError(f'unsupported graph constructor {fn!r} (session and ccs graphs are built)')


def collate_fn_factory(*seq_to_graph_fns):
    """`collate_fn_factory` (`collate.py:219-230`): samples -> ([SessionBatch, ...], LongTensor labels)."""
    kinds = [_kind(f) for f in seq_to_graph_fns]

    def collate_fn(samples):
        seqs, labels = zip(*samples)
        return [SessionBatch.build(seqs, labels, k, 1) for k in kinds], torch.LongTensor(labels)
    return collate_fn


def collate_fn_factory_ccs(seq_to_graph_fns, order):
    """`collate_fn_factory_ccs` (`collate.py:232-256`)."""
    kinds = [_kind(f) for f in seq_to_graph_fns]

    def collate_fn(samples):
        seqs, labels = zip(*samples)
        return [SessionBatch.build(seqs, labels, k, order if k == 'ccs' else 1) for k in kinds], torch.LongTensor(labels)
    return collate_fn
