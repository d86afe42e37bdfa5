"""Host-side logic that needs no GPU: flat parameter layout (one buffer -> one fused Adam launch / one all-reduce), the
`fix_weight_decay` name routing (`src/utils/train.py:12-23`) and parameter registration order of the drop-in modules."""
import pytest
import torch

from oracle import models as OM


def _models(pkg):
    from sessionrec_pytorch_b200.msgifsr import MSGIFSR
    from sessionrec_pytorch_b200.srgnn import NISER, SRGNN
    return {'srgnn': SRGNN(50, 16, 2), 'niser': NISER(50, 16, 1),
            'msgifsr_k1': MSGIFSR(50, 'x', 16, 1, order=1, extra=False, fusion=False),
            'msgifsr_k3': MSGIFSR(50, 'x', 16, 2, order=3, extra=True, fusion=True)}


@pytest.mark.parametrize('name', ['srgnn', 'niser', 'msgifsr_k1', 'msgifsr_k3'])
def test_flat_params_layout_and_views(pkg, name):
    from sessionrec_pytorch_b200.flat import ALIGN, FlatParams
    m = _models(pkg)[name]
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    keys = list(m.state_dict().keys())
    fp = FlatParams(m)
    assert fp.valid() and fp.names == [n for n, _ in m.named_parameters()]
    assert list(m.state_dict().keys()) == keys, 'flattening must not touch the state_dict keys'
    end = 0
    for (n, p), off in zip(m.named_parameters(), fp.offsets):
        assert off % ALIGN == 0 and off >= end, (n, off)                       # 256-byte aligned, no overlap
        end = off + p.numel()
        assert p.data_ptr() == fp.data.data_ptr() + 4 * off                    # the Parameter IS a view of the flat buffer
        assert torch.equal(p.detach(), before[n])
        assert fp.view(fp.grad, n).shape == p.shape
    assert end <= fp.total == fp.data.numel() == fp.grad.numel()
    # a write through the flat buffer is seen by the module, a load_state_dict lands in the flat buffer
    fp.data.zero_()
    assert all(float(p.detach().abs().max()) == 0.0 for p in m.parameters())
    m.load_state_dict({k: (before[k] if k in before else v) for k, v in m.state_dict().items()})
    assert fp.valid() and torch.equal(fp.view(fp.data, fp.names[-1]), before[fp.names[-1]])


@pytest.mark.parametrize('name', ['srgnn', 'niser', 'msgifsr_k1', 'msgifsr_k3'])
def test_decay_segments_follow_fix_weight_decay(pkg, name):
    from sessionrec_pytorch_b200.flat import FlatParams
    m = _models(pkg)[name]
    fp = FlatParams(m)
    seg_off, seg_decay = fp.decay_segments(1e-4)
    dec, no = OM.decay_split(fp.names)                                         # the oracle's restatement of train.py:12-23
    assert seg_off.dtype == torch.int64 and seg_off.tolist() == fp.offsets + [fp.total]
    for n, w in zip(fp.names, seg_decay.tolist()):
        assert (w == 0.0) == (n in no), n
        assert w == 0.0 or abs(w - 1e-4) < 1e-10                          # fp32 rounding of 1e-4
    assert no and all(('bias' in n) or ('activation' in n) or ('batch_norm' in n) for n in no)
    assert len(dec) + len(no) == len(fp.names)


def test_models_refuse_to_run_on_cpu(pkg):
    m = _models(pkg)['msgifsr_k1']
    b = pkg.SessionBatch.build([[1, 2, 3], [4]], [5, 6], 'ccs', 1)
    with pytest.raises(pkg._lib.SessRecError, match='no CPU fallback'):
        m(b)
    with pytest.raises(pkg._lib.SessRecError, match='no CPU fallback'):
        m.train_step(b)
