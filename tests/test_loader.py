"""BatchPrefetcher (host logic; the CUDA copy path is covered by bench.py's end-to-end-with-build figure): batches come
out in order and equal to direct builds although the ring has fewer slots than batches, buffers are reused, a builder
error reaches the consumer."""
import numpy as np
import pytest
import torch


def _raw(n, seed=0):
    from sessionrec_pytorch_b200.synthetic import SessionSampler
    smp = SessionSampler(500, seed=seed)
    return [smp.batch(int(b)) for b in np.random.default_rng(seed).integers(1, 70, size=n)]


@pytest.mark.parametrize('kind,order', [('session', 1), ('ccs', 1), ('ccs', 3)])
def test_prefetcher_yields_the_same_batches_in_order(pkg, kind, order):
    from sessionrec_pytorch_b200.loader import BatchPrefetcher
    raw = _raw(17, seed=order)
    ptrs = set()
    n = 0
    for (items, offs, labels), got in zip(raw, BatchPrefetcher(raw, kind, order, depth=2)):
        ref = pkg.SessionBatch.build_flat(items, offs, labels, kind, order)
        assert got.B == ref.B and got.K == ref.K and got.kind == ref.kind
        assert torch.equal(got.buf, ref.buf)                      # valid until the next batch is requested
        ptrs.add(got.buf.data_ptr())
        n += 1
    assert n == len(raw)
    assert len(ptrs) <= 4, 'host buffers are reused (2 slots, regrown at most once each)'


def test_prefetcher_is_exhausted_once_and_forwards_builder_errors(pkg):
    from sessionrec_pytorch_b200.loader import BatchPrefetcher
    raw = _raw(3)
    it = BatchPrefetcher(raw, 'ccs', 1)
    assert len(list(it)) == 3
    with pytest.raises(StopIteration):
        next(it)
    items, offs, labels = raw[0]
    bad = (items, np.zeros_like(offs), labels)                  # every session empty
    it = BatchPrefetcher([raw[1], bad, raw[2]], 'ccs', 1)
    next(it)
    with pytest.raises(pkg._lib.SessRecError, match='empty'):
        next(it)


def test_build_flat_into_caller_buffer(pkg):
    items, offs, labels = _raw(1)[0]
    ref = pkg.SessionBatch.build_flat(items, offs, labels, 'ccs', 2)
    words = pkg.SessionBatch.batch_words(int(offs[-1]), len(offs) - 1, 'ccs', 2)
    out = torch.empty(words, dtype=torch.int32)
    got = pkg.SessionBatch.build_flat(items, offs, labels, 'ccs', 2, out=out)
    assert got.buf.data_ptr() == out.data_ptr() and torch.equal(got.buf, ref.buf)
    with pytest.raises(pkg._lib.SessRecError, match='too small'):
        pkg.SessionBatch.build_flat(items, offs, labels, 'ccs', 2, out=torch.empty(100, dtype=torch.int32))
    with pytest.raises(pkg._lib.SessRecError):
        pkg.SessionBatch.build_flat(items, offs, labels, 'ccs', 2, out=torch.empty(words, dtype=torch.int64))


@pytest.mark.parametrize('kind,order,threads', [('session', 1, 1), ('ccs', 1, 4), ('ccs', 3, 3)])
def test_epoch_batches_equal_per_batch_builds(pkg, kind, order, threads):
    """One buffer for the whole pass: every slice is bit-identical to building that batch alone (also the short last one),
    starts on a 256-byte boundary, and the pass survives `.to()` as one copy."""
    from sessionrec_pytorch_b200.dataset import AugmentedDataset
    from sessionrec_pytorch_b200.loader import EpochBatches, flatten_samples
    from sessionrec_pytorch_b200.synthetic import SessionSampler
    seqs, _ = SessionSampler(300, seed=9).sessions(40)
    ds = AugmentedDataset([s + [1] for s in seqs])                 # every prefix of every session
    samples = [ds[i] for i in range(len(ds))]
    items, offs, labels = flatten_samples(ds)
    assert len(offs) - 1 == len(ds) == len(labels)
    ep = EpochBatches(items, offs, labels, 32, kind, order, threads=threads)
    assert len(ep) == (len(ds) + 31) // 32
    for i, b in enumerate(ep):
        chunk = samples[i * 32:(i + 1) * 32]
        ref = pkg.SessionBatch.build([s for s, _ in chunk], [l for _, l in chunk], kind, order)
        assert b.B == len(chunk) == ref.B and torch.equal(b.buf, ref.buf)
        assert (b.buf.data_ptr() - ep.buf.data_ptr()) % 256 == 0
    moved = ep.to('cpu')
    assert len(moved) == len(ep) and torch.equal(moved[len(ep) - 1].buf, ep[len(ep) - 1].buf)
    assert len(EpochBatches(items, offs, labels, 32, kind, order, drop_last=True)) == len(ds) // 32
