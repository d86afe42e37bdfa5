#!/usr/bin/env python
"""Runs only the fused scoring + CE head kernels (csrc/flash_ce.cu) at a BASELINE shape: CUDA-event timing per kernel,
and a short command line for `ncu --set full -k regex:fce_`.  Usage: head_probe.py [B V d] [--iters n]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from __graft_entry__ import build, load_package  # noqa: E402

build()
load_package()
from sessionrec_pytorch_b200 import ops  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
B, V, d = (int(x) for x in args[:3]) if len(args) >= 3 else (512, 43097, 96)
iters = int(sys.argv[sys.argv.index('--iters') + 1]) if '--iters' in sys.argv else 10
dev = 'cuda'
g = torch.Generator().manual_seed(1)
s = torch.nn.functional.normalize(torch.randn(B, d, generator=g), dim=-1).to(dev)
E = torch.nn.functional.normalize(torch.randn(V, d, generator=g), dim=-1).to(dev)
Sh, Sl = (torch.empty(B, d, dtype=torch.int16, device=dev) for _ in range(2))
Eh, El = (torch.empty(V, d, dtype=torch.int16, device=dev) for _ in range(2))
ops.split_bf16(s, d, B, d, Sh, Sl, d)
ops.split_bf16(E, d, V, d, Eh, El, d)
lab = torch.randint(0, V, (B,), generator=g).int().to(dev)
lse, nll = torch.empty(B, device=dev), torch.empty(B, device=dev)
part = torch.empty(ops.flash_ce_part_floats(B, V), device=dev)
parts = ops.flash_ce_bwd_parts(B)
dS = torch.empty(B, d, device=dev)
dEp = torch.empty(parts, V, d, device=dev)
one = torch.ones(1, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def fwd():
    ops.flash_ce_fwd(B, V, d, Sh, Sl, d, Eh, El, d, 12.0, lab, lse, nll, part)


def bwd():
    ops.flash_ce_bwd(B, V, d, Sh, Sl, d, Eh, El, d, 12.0, lab, lse, one, dS, dEp)


for name, fn in (('fwd', fwd), ('bwd', bwd)):
    for _ in range(3):
        fn()
    for mode in ('L2 flushed', 'warm'):
        ts = []
        for _ in range(iters):
            if mode == 'L2 flushed':
                flush.fill_(0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f'{name} B={B} V={V} d={d} [{mode}]: median {np.median(ts):.1f} us, min {min(ts):.1f} us')
flops = 2.0 * B * V * d
print(f'algorithmic GFLOP: fwd {flops / 1e9:.2f}, bwd {2 * flops / 1e9:.2f}')
