"""BatchPrefetcher with a CUDA device: pinned ring, non-blocking H2D, slot reuse gated by the copy's event."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_device_batches_equal_direct_builds(pkg):
    from sessionrec_pytorch_b200.loader import BatchPrefetcher
    from sessionrec_pytorch_b200.synthetic import SessionSampler
    smp = SessionSampler(2000, seed=5)
    raw = [smp.batch(int(b)) for b in np.random.default_rng(5).integers(1, 600, size=40)]
    got = list(BatchPrefetcher(raw, 'ccs', 1, device='cuda', depth=2))      # 40 batches through 2 pinned slots
    torch.cuda.synchronize()
    assert len(got) == len(raw)
    for (items, offs, labels), g in zip(raw, got):
        ref = pkg.SessionBatch.build_flat(items, offs, labels, 'ccs', 1)
        assert g.device.type == 'cuda' and g.B == ref.B and g.N1 == ref.N1
        assert torch.equal(g.buf.cpu(), ref.buf)
