#!/usr/bin/env python
"""Digest of a bench.py JSON line (headline workload + the `workloads` block)."""
import json, sys


def show(tag, d):
    if 'error' in d:
        print(tag, 'ERROR', d['error'])
        return
    r = d.get('roofline') or {}
    cb = d.get('cpu_baseline') or {}
    print(tag, {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'host_enqueue_ms_per_step') if k in d}, 'e2e', d['e2e']['value'],
          'roofline', r.get('achieved'), r.get('frac'), r.get('ms_per_kernel'), 'cpu', cb.get('value'), cb.get('kind'), cb.get('cores'),
          'collate ms', cb.get('collate_ms_per_batch'))


d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
show(d.get('headline_workload', 'main'), d)
for k, v in (d.get('workloads') or {}).items():
    show(k, v)
