"""Shared helpers of the parity tests (oracle = checker, never the product)."""
import json
from pathlib import Path

import numpy as np
import torch

from oracle import collate as OC
from oracle import models as OM

GOLD = Path(__file__).resolve().parent / 'golden'
RTOL = 1e-4      # north_star: fp32 logits / loss within 1e-4 relative


def golden(name):
    return torch.load(GOLD / name, weights_only=False)


def golden_json(name):
    return json.loads((GOLD / name).read_text())


def oracle_params(sd, grad=True):
    p = {}
    for k, v in sd.items():
        t = v.detach().clone().cpu()
        if grad and t.is_floating_point():
            t.requires_grad_(True)
        p[k] = t
    return p


def oracle_batch(samples, kind, K):
    return OC.build_batch([s for s, _ in samples], [l for _, l in samples], kind, K)


def assert_close(name, got, ref, rtol=RTOL, floor=1.0):
    """|got - ref| <= rtol * max(|ref|, floor) elementwise."""
    got = got.detach().cpu().double()
    ref = ref.detach().cpu().double()
    assert got.shape == ref.shape, f'{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}'
    assert torch.isfinite(got).all(), f'{name}: non-finite values'
    err = (got - ref).abs()
    tol = rtol * ref.abs().clamp(min=floor)
    bad = err > tol
    if bad.any():
        i = int((err / tol).argmax())
        raise AssertionError(f'{name}: {int(bad.sum())}/{bad.numel()} elements off; worst |d|={err.flatten()[i]:.3e} '
                             f'ref={ref.flatten()[i]:.6e} got={got.flatten()[i]:.6e} (flat index {i})')


def assert_close_after_adam(name, got, ref, lr, steps, rtol=RTOL, floor=0.1, max_frac=2e-3):
    """Parameters after `steps` Adam updates.  Adam's update lr * m / (sqrt(v) + eps) is sign-like: an element whose
    gradient is within rounding of zero (|g| ~ eps = 1e-8) moves by up to lr per step in a direction that 1e-9 of
    gradient noise decides.  So: every element within rtol * max(|ref|, floor), except at most `max_frac` of them,
    and those no further than steps * lr away."""
    g, r = got.detach().cpu().double(), ref.detach().cpu().double()
    assert g.shape == r.shape and torch.isfinite(g).all(), name
    err = (g - r).abs()
    bad = err > rtol * r.abs().clamp(min=floor)
    nbad = int(bad.sum())
    assert nbad <= max_frac * bad.numel(), f'{name}: {nbad}/{bad.numel()} elements off (worst {float(err.max()):.3e})'
    assert float(err.max()) <= steps * lr * 1.001, f'{name}: an element moved {float(err.max()):.3e} > steps * lr'


def assert_grad_close(name, got, ref, rtol=RTOL):
    """Max-norm relative: max|got - ref| <= rtol * max|ref| (absolute 1e-7 for numerically-zero grads)."""
    if ref is None:
        assert got is None or float(got.abs().max()) == 0.0, f'{name}: reference has no gradient'
        return
    assert got is not None, f'{name}: missing gradient'
    got = got.detach().cpu().double()
    ref = ref.detach().cpu().double()
    assert got.shape == ref.shape, f'{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}'
    assert torch.isfinite(got).all(), f'{name}: non-finite gradient'
    err = float((got - ref).abs().max())
    scale = max(float(ref.abs().max()), 1e-3)
    assert err <= rtol * scale, f'{name}: max|d|={err:.3e} vs max|ref|={float(ref.abs().max()):.3e} (rel {err / scale:.2e})'


def assert_grad_close_robust(name, got, ref, rtol=RTOL, l2_rtol=1e-2):
    """Max-norm criterion of assert_grad_close, or - when a few elements miss it - a relative L2 error below 1 %.
    With dropout a near-tie (< 1e-6) in the max over the 8 GAT heads can resolve differently in fp32 on the two sides,
    which re-routes ONE node's gradient to another head and perturbs every upstream parameter slightly; a wrong
    formula would show up as an O(1) L2 error."""
    if ref is None:
        return assert_grad_close(name, got, ref, rtol)
    g, r = got.detach().cpu().double(), ref.detach().cpu().double()
    assert g.shape == r.shape and torch.isfinite(g).all(), name
    scale = max(float(r.abs().max()), 1e-3)
    err = (g - r).abs()
    if float(err.max()) <= rtol * scale:
        return
    l2 = float((g - r).norm() / r.norm().clamp(min=1e-12))
    assert l2 <= l2_rtol, f'{name}: max|d|={float(err.max()):.3e} vs max|ref|={scale:.3e}, relative L2 error {l2:.2e}'


def run_oracle(case_model, params, ob, L=1, fusion=False, drop=OM.NO_DROPOUT, extra=False):
    if case_model == 'MSGIFSR':
        return OM.msgifsr_forward(params, ob, drop=drop, num_layers=L, fusion=fusion, extra=extra)
    return OM.srgnn_forward(params, ob, drop=drop, num_layers=L, niser=(case_model == 'NISER'))
