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


class _FakeBatch:
    def __init__(self, labels):
        self.labels = labels

    def to(self, device):
        return self


class _FakeModel:
    """Stands in for a drop-in module: records the learning rate of every step, ranks the label at a scripted position."""

    def __init__(self, rank_per_eval):
        self.rank_per_eval, self.evals, self.lrs, self._opt = rank_per_eval, 0, [], None
        self.training = True

    def configure_optimizer(self, lr=1e-3, weight_decay=1e-4):
        self._opt = dict(lr=lr, weight_decay=weight_decay)
        return self._opt

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def set_lr(self, lr):
        self._opt['lr'] = lr

    def train_step(self, batch):
        self.lrs.append(self._opt['lr'])
        return torch.tensor(float(len(self.lrs)))

    def topk(self, batch, k=20):
        r = self.rank_per_eval[min(self.evals, len(self.rank_per_eval) - 1)]
        out = torch.full((len(batch.labels), k), -1, dtype=torch.long)
        if r <= k:
            out[:, r - 1] = batch.labels
        return out


def _loaders(n_train=5, n_test=2):
    lab = torch.arange(4)
    return ([([_FakeBatch(lab)], lab) for _ in range(n_train)], [([_FakeBatch(lab)], lab) for _ in range(n_test)])


def test_train_runner_follows_the_reference_epoch_loop(pkg):
    """StepLR(3, 0.1) per epoch, evaluation before and after every epoch, early stop after `patience` epochs in which MRR and
    HR both fell (`utils/train.py:57-127`)."""
    from sessionrec_pytorch_b200.train import TrainRunner

    class Counting(_FakeModel):
        def topk(self, batch, k=20):
            out = super().topk(batch, k)
            self.calls = getattr(self, 'calls', 0) + 1
            if self.calls % 2 == 0:            # two test batches per evaluate()
                self.evals += 1
            return out

    # label rank per evaluate(): initial, then after epochs 0..: improving, then three epochs where the label leaves the top 20
    m = Counting([5, 4, 2, 1, 30, 30, 30, 30, 30])
    train, test = _loaders()
    logs = []
    r = TrainRunner('x', m, train, test, 'cpu', lr=1e-2, weight_decay=1e-4, patience=3)
    assert m._opt == dict(lr=1e-2, weight_decay=1e-4)
    mrr, hit = r.train(20, log_interval=2, log=logs.append)
    assert (mrr, hit) == (1.0, 1.0)                                   # best epoch: rank 1
    # epochs 0-2 improve, epochs 3-5 are bad -> stop inside epoch index 5 (6 epochs trained)
    assert len(m.lrs) == 6 * 5 and r.epoch == 5
    expect = [1e-2] * 15 + [1e-3] * 15
    assert all(abs(a - b) < 1e-12 for a, b in zip(m.lrs, expect)), m.lrs
    assert sum(s.startswith('Epoch') for s in logs) == 6 and any(s.startswith('Batch 2:') for s in logs)
    assert m.training is False                                        # left in eval mode by the last evaluate()


def test_train_runner_accepts_bare_batches(pkg):
    from sessionrec_pytorch_b200.train import TrainRunner
    m = _FakeModel([1])
    lab = torch.arange(3)
    r = TrainRunner('x', m, [_FakeBatch(lab)] * 4, [([_FakeBatch(lab)], lab)], 'cpu', lr=1e-3, weight_decay=0, patience=2)
    r.train(1, log=lambda s: None)
    assert len(m.lrs) == 4 and m._opt['weight_decay'] == 0 and abs(m._opt['lr'] - 1e-3) < 1e-15


def test_inactive_parameters_match_the_reference_goldens(pkg):
    """Parameters the reference's forward never reaches keep grad None, so torch.optim.Adam never touches them: the static
    rule of the drop-ins (`_inactive_params`) names exactly the parameters whose value did not change over the reference's own
    training steps in the trajectory goldens (plus biases whose gradient is exactly zero: no decay, no update either way)."""
    import torch
    from sessionrec_pytorch_b200.msgifsr import MSGIFSR
    from sessionrec_pytorch_b200.srgnn import NISER, SRGNN
    from tests.util import golden
    seen = 0
    for f in ('train_golden.pt', 'train_extra_golden.pt'):
        for name, c in golden(f).items():
            untouched = {n for n, v in c['final_state'].items() if torch.equal(v, c['params'][n])}
            if c['model'] == 'MSGIFSR':
                m = MSGIFSR(c['V'], 'g', c['d'], 1, dropout=0.0, order=c['K'], extra=c.get('extra', False), fusion=False)
            else:
                m = {'SRGNN': SRGNN, 'NISER': NISER}[c['model']](c['V'], c['d'], 1, 0.0)
            mine = set(m._inactive_params(None))
            assert mine <= untouched, (name, sorted(mine - untouched))
            assert all('bias' in n for n in untouched - mine), (name, sorted(untouched - mine))
            seen += 1
    assert seen >= 5


def test_adam_segments_for_inactive_parameters_and_owned_rows(pkg):
    import torch
    from sessionrec_pytorch_b200.flat import FlatParams
    m = torch.nn.Module()
    m.emb = torch.nn.Embedding(10, 8)
    m.lin = torch.nn.Linear(8, 4)
    fp = FlatParams(m)
    off, dec = fp.decay_segments(1e-4, inactive={'lin.weight'}, owned_rows=('emb.weight', 3, 7))
    names = dict(zip(fp.names, fp.offsets))
    e0 = names['emb.weight']
    assert off.tolist() == [e0, e0 + 3 * 8, e0 + 7 * 8, names['lin.weight'], names['lin.bias'], fp.total]
    assert dec.tolist()[0] == -1.0 and abs(dec.tolist()[1] - 1e-4) < 1e-9 and dec.tolist()[2] == -1.0      # only rows [3, 7) are ours
    assert dec.tolist()[3] == -1.0 and dec.tolist()[4] == 0.0                                                # inactive weight, bias without decay


@pytest.mark.parametrize('K,L', [(2, 1), (3, 2), (4, 1)])
def test_order_k_native_step_slot_table(K, L):
    """The slot table handed to srk_msgifsr_k_train_step (csrc/step_k.cu): every name is a parameter of the module, no name
    twice, and the count is the one the C side checks (1 + L*2*(K+1)*4 + (K-1)*4 + 5)."""
    from __graft_entry__ import load_package
    load_package()
    from sessionrec_pytorch_b200.msgifsr import MSGIFSR
    m = MSGIFSR(300, 'x', 16, L, dropout=0.1, order=K, extra=False, fusion=False)
    names = m._slot_names_k()
    params = dict(m.named_parameters())
    assert all(n in params for n in names), [n for n in names if n not in params]
    assert len(set(names)) == len(names) == 1 + L * 2 * (K + 1) * 4 + (K - 1) * 4 + 5
    # shapes the step relies on
    d = 16
    assert tuple(params[f'expander.GRUs.{K - 2}.weight_ih_l0'].shape) == (3 * d, d)
    assert tuple(params['layers.0.conv1.mods.inter.fc.weight'].shape) == (8 * d, d)
    assert tuple(params['fc_sr.0.weight'].shape) == (d, 2 * d)
