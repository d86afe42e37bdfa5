#!/usr/bin/env python
"""Per-tile timeline (SM clock cycles) of CTA 0 of the fused head kernels: where producer / MMA issuer / epilogue wait."""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from __graft_entry__ import build, load_package  # noqa: E402

build()
load_package()
from sessionrec_pytorch_b200 import _lib, ops  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
B, V, d = (int(x) for x in args[:3]) if len(args) >= 3 else (512, 43097, 96)
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
dS = torch.empty(B, d, device=dev)
dEp = torch.empty(ops.flash_ce_bwd_parts(B), V, d, device=dev)
one = torch.ones(1, device=dev)
trace = torch.zeros(11 * 64 * 8, dtype=torch.int64, device=dev)
fn = _lib.lib().functions['srk_flash_ce_set_trace']


def run(name, f, cols):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    trace.zero_()
    fn(ctypes.c_void_p(trace.data_ptr()))
    f()
    torch.cuda.synchronize()
    fn(None)
    t = trace.cpu().view(11, 64, 8)
    t0 = int(t[t > 0].min())
    print(f'== {name}: CTA 0, cycles since its first stamp')
    for role, rn in enumerate(('tma ', 'mma ') + tuple(f'ep{w} ' for w in range(8)) + ('drn ',)):
        for it in range(64):
            row = t[role, it]
            if int(row.max()) == 0:
                continue
            print(f'  {rn} tile {it:2d}: ' + '  '.join(f'{c}={int(row[k]) - t0:7d}' for k, c in enumerate(cols[2 if 2 <= role < 10 else min(role, 3)]) if int(row[k]) > 0))


run('fwd', lambda: ops.flash_ce_fwd(B, V, d, Sh, Sl, d, Eh, El, d, 12.0, lab, lse, nll, part),
    (('slot_free', 'issued'), ('e_full', 'z_empty', 'mma_issued'), ('z_full', 'done'), ()))
run('bwd', lambda: ops.flash_ce_bwd(B, V, d, Sh, Sl, d, Eh, El, d, 12.0, lab, lse, one, dS, dEp),
    (('slot_free', 'issued'), ('e_full', 'logits_issued', 'd_full', 'grads_issued'),
     ('z_full', 'math_done', 'd_free', 'dz_written'), ('w0_start', 'w1_start', 'w2_start', 'w3_start')))
