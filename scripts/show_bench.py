#!/usr/bin/env python
"""One-line digest of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get('roofline') or {}
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'host_enqueue_ms_per_step') if k in d}, 'e2e', d['e2e']['value'],
      'roofline', r.get('achieved'), r.get('frac'), r.get('ms_per_kernel'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
