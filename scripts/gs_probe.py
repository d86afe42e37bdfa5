#!/usr/bin/env python
"""Runs only the embedding gather (K1) and gradient scatter-add (K8) kernels on the stress shape (table >> L2) for
`ncu --set full -k regex:(gather_fwd|scatter_bwd)`; prints their CUDA-event bandwidths too."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from __graft_entry__ import build, load_package  # noqa: E402

build()
load_package()
shapes = [('cfg2 shape: V=17000 d=256 N=9000 (table 17 MB, L2-resident)', 17000, 256, 9000),
          ('stress: V=4M d=256 N=1M (table 4.1 GB >> 126 MB L2)', 4_000_000, 256, 1_000_000)]
if '--stress-only' in sys.argv:
    shapes = shapes[1:]
print(json.dumps(bench.gather_scatter_probe(torch.device('cuda', 0), bench.peaks(), shapes)))
