#!/usr/bin/env python
"""bench.py - train sessions/sec of the B200-native session-rec training path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1] [--also cfg2] [--parallelism dp|shard]
                  [--impl reference]

A "step" = one TrainRunner iteration (`src/utils/train.py:95-101`): forward + nll_loss + backward + Adam(L2) step
over one synthetic batch of the named BASELINE.json configuration.  One process per GPU (torchrun sets RANK /
LOCAL_RANK / WORLD_SIZE); data-parallel ranks own independent batches and exchange the flat gradient with one NCCL
all-reduce per step (weak scaling).  Prints ONE JSON line on rank 0.

`value`     device-timed (CUDA events per step, L2 flushed between steps), batches already resident in HBM.
`e2e`       the same metric through the public API with HOST (pinned) batch buffers: H2D copy of the batch and a
            D2H read of the loss inside every timed step.
            `e2e_loss_read_one_step_late` (N = 1): the same loop with every step's loss copied D2H asynchronously and read one
            step later - shows how much of the value / e2e gap is the per-step synchronisation; never replaces `e2e`.
`roofline`  the dominant kernel family (catalog scoring GEMMs), timed live in isolation with CUDA events.
`cpu_baseline` the reference's own CPU implementation of the path timed on this box's host cores on a bounded sample
            of the same workload: the UNMODIFIED reference sources (oracle/_ref, a byte-identical copy of
            /root/reference/src made by oracle/build_ref.py) over oracle/dgl_shim, kind "reference"; the restated port
            (oracle/models.py, kind "port") only when that copy is absent.  `--impl reference` times only that.
`workloads` (N = 1) the same measurements for the other single-GPU configuration of BASELINE.json (cfg2 = SRGNN on the
            Yoochoose1/64 shape, B = 2048, d = 256); the top-level keys are the headline workload named in `config`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import build, load_package  # noqa: E402

METRIC = 'train sessions/sec'
UNIT = 'sessions/s'


def peaks():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [v for v in (num(r[0]) for r in self.rows if r) if v is not None]
        mx = [v for v in (num(r[1]) for r in self.rows if len(r) > 1) if v is not None]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == 'Active'})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def build_model(cfg, device):
    from sessionrec_pytorch_b200.msgifsr import MSGIFSR
    from sessionrec_pytorch_b200.srgnn import NISER, SRGNN
    torch.manual_seed(123)
    if cfg['model'] == 'MSGIFSR':
        m = MSGIFSR(cfg['V'], 'synthetic', cfg['d'], cfg['layers'], dropout=cfg['dropout'], order=cfg['order'],
                    extra=False, fusion=False)
    else:
        m = {'SRGNN': SRGNN, 'NISER': NISER}[cfg['model']](cfg['V'], cfg['d'], cfg['layers'], cfg['dropout'])
        if os.environ.get('SESSREC_BENCH_NO_DEAD') == '1':      # diagnosis only (is the dead-layer stream the critical path?)
            m.compute_dead_layers = False
            print('[bench] DIAGNOSTIC RUN: dead GGNN layers skipped - not a valid bench line', file=sys.stderr)
    return m.to(device).train()


def kind_of(cfg):
    return 'session' if cfg['model'] in ('SRGNN', 'NISER') else 'ccs'


def cpu_oracle_steps(cfg, sessions, steps, warmup):
    """The reference loop body on the host cores via the oracle port; returns (sessions/s, cores, seconds)."""
    from oracle import collate as OC
    from oracle import models as OM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = build_model(cfg, 'cpu')
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}
    opt = OM.make_adam({k: v for k, v in params.items() if v.is_floating_point()}, 1e-3, 1e-4)
    drop = OM.Dropout(cfg['dropout'], True, None)
    batches = [OC.build_batch(s, l, kind_of(cfg), cfg['order']) for s, l in sessions]
    t_total, n = 0.0, 0
    for it in range(warmup + steps):
        ob = batches[it % len(batches)]
        t0 = time.perf_counter()
        opt.zero_grad()
        if cfg['model'] == 'MSGIFSR':
            out = OM.msgifsr_forward(params, ob, drop=drop, num_layers=cfg['layers'])
        else:
            out = OM.srgnn_forward(params, ob, drop=drop, num_layers=cfg['layers'], niser=cfg['model'] == 'NISER')
        loss = OM.nll(out, ob['labels'])
        loss.backward()
        opt.step()
        float(loss.detach())
        if it >= warmup:
            t_total += time.perf_counter() - t0
            n += ob['B']
    return n / t_total, cores, t_total


def cpu_reference_steps(cfg, raw, steps, warmup):
    """The UNMODIFIED reference on the host cores: its own model classes, its own collate_fn and the loop body of its own
    `TrainRunner.train` (`src/utils/train.py:95-103`: zero_grad / model(*inputs) / isnan assert / nll_loss / backward /
    optimizer.step / loss.item()) with the optimizer its `TrainRunner.__init__` builds, imported from oracle/_ref over
    oracle/dgl_shim.  raw: [(seqs, labels)] python lists.  Returns dict(compute-only sessions/s, collate ms/batch, ...)."""
    from oracle import ref_import
    R = ref_import.load()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(123)
    if cfg['model'] == 'MSGIFSR':
        m = R.MSGIFSR(cfg['V'], 'synthetic', cfg['d'], cfg['layers'], dropout=cfg['dropout'], order=cfg['order'], extra=False,
                      fusion=False)
        fn = R.collate.collate_fn_factory_ccs((R.collate.seq_to_ccs_graph,), order=cfg['order'])
    else:
        m = getattr(R, cfg['model'])(cfg['V'], cfg['d'], cfg['layers'], cfg['dropout'])
        fn = R.collate.collate_fn_factory(R.collate.seq_to_session_graph)
    runner = R.train.TrainRunner('synthetic', m, [], [], torch.device('cpu'), lr=1e-3, weight_decay=1e-4, patience=3)
    t0 = time.perf_counter()
    batches = [fn(list(zip(s, l))) for s, l in raw]
    collate_s = (time.perf_counter() - t0) / len(raw)
    m.train()
    t_total, n = 0.0, 0
    for it in range(warmup + steps):
        batch = batches[it % len(batches)]
        t0 = time.perf_counter()
        inputs, labels = R.train.prepare_batch(batch, torch.device('cpu'))
        runner.optimizer.zero_grad()
        scores = m(*inputs)
        assert not torch.isnan(scores).any()
        loss = torch.nn.functional.nll_loss(scores, labels)
        loss.backward()
        runner.optimizer.step()
        loss.item()
        if it >= warmup:
            t_total += time.perf_counter() - t0
            n += len(labels)
    step_s = t_total / steps
    return dict(value=n / t_total, cores=cores, seconds=t_total, collate_ms_per_batch=1e3 * collate_s,
                with_collate=cfg['B'] / (step_s + collate_s), root=str(ref_import.REFERENCE_ROOT))


def cpu_arm(cfg, workload, steps, warmup):
    """cpu_baseline object for one workload: the unmodified reference when oracle/_ref (or /root/reference) is there, else
    the restated port."""
    from sessionrec_pytorch_b200.synthetic import SessionSampler
    from oracle import ref_import
    smp = SessionSampler(cfg['V'], seed=123)
    raw = [smp.sessions(cfg['B']) for _ in range(min(4, steps + warmup))]
    if ref_import.available():
        r = cpu_reference_steps(cfg, raw, steps, warmup)
        return dict(value=round(r['value'], 2), unit=UNIT, cores=r['cores'], kind='reference',
                    sample=f"{steps} full steps (B={cfg['B']}) of {workload} after {warmup} warm-up, {r['seconds']:.1f} s of CPU work; "
                           f"batches pre-collated (compute only, like the GPU arm's resident batches)",
                    what='unmodified reference src/models + TrainRunner loop body + torch.optim.Adam over oracle/dgl_shim '
                         '(DGL 0.7.2 is not installable offline), imported from ' + r['root'],
                    collate_ms_per_batch=round(r['collate_ms_per_batch'], 2),
                    value_with_reference_collate=round(r['with_collate'], 2))
    v, cores, secs = cpu_oracle_steps(cfg, raw, steps, warmup)
    return dict(value=round(v, 2), unit=UNIT, cores=cores, kind='port',
                sample=f"{steps} full steps (B={cfg['B']}) of {workload} after {warmup} warm-up, {secs:.1f} s",
                what='oracle/models.py restatement (oracle/_ref absent)')


def roofline_probe(cfg, device, pk):
    """Times the three catalog GEMM launches of one step (Z = s E^T, dS = dZ E, dE = dZ^T s) in isolation, L2 flushed:
    tcgen05 3xTF32 `umma_gemm_kernel` when the embedding dim fits one UMMA N tile, else the fp32 `sgemm_kernel`."""
    from sessionrec_pytorch_b200 import ops
    B, V, d = cfg['B'], cfg['V'], cfg['d']
    ldz = (V + 3) // 4 * 4
    g = torch.Generator(device='cpu').manual_seed(1)
    s = torch.randn(B, d, generator=g).to(device)
    E = torch.randn(V, d, generator=g).to(device)
    Z = torch.randn(B, ldz, generator=g).to(device) * 1e-3
    dS = torch.zeros(B, d, device=device)
    dE = torch.zeros(V, d, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    umma = d <= 256 and os.environ.get('SESSREC_NO_UMMA', '0') != '1'
    flash = umma and ops.flash_ce_supported(d) and os.environ.get('SESSREC_NO_FLASH_CE', '0') != '1'
    per_kernel = None
    if flash:
        # fused head (csrc/flash_ce.cu): forward = soft-max statistics only, backward recomputes the logit tiles and
        # runs both gradient products from shared memory; the (B, V) logits never exist in HBM
        sn = torch.nn.functional.normalize(s, dim=-1)
        En = torch.nn.functional.normalize(E, dim=-1)
        Sh, Sl = (torch.empty(B, d, dtype=torch.int16, device=device) for _ in range(2))
        Eh, El = (torch.empty(V, d, dtype=torch.int16, device=device) for _ in range(2))
        ops.split_bf16(sn, d, B, d, Sh, Sl, d)
        ops.split_bf16(En, d, V, d, Eh, El, d)
        lse, nll = torch.empty(B, device=device), torch.empty(B, device=device)
        part = torch.empty(ops.flash_ce_part_floats(B, V), device=device)
        lab = torch.zeros(B, dtype=torch.int32, device=device)
        parts = ops.flash_ce_bwd_parts(B)
        dEp = torch.empty(parts, V, d, device=device)
        one = torch.ones(1, device=device)

        def run_f():
            ops.flash_ce_fwd(B, V, d, Sh, Sl, d, Eh, El, d, 12.0, lab, lse, nll, part)

        def run_b():
            ops.flash_ce_bwd(B, V, d, Sh, Sl, d, Eh, El, d, 12.0, lab, lse, one, dS, dEp)

        def run():
            run_f()
            run_b()
        per_kernel = {}
        for nm, fn in (('fce_fwd_kernel (+ finalize)', run_f), ('fce_bwd_kernel (+ dS memset)', run_b)):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                flush.fill_(0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            per_kernel[nm] = round(float(np.median(ts)), 4)
        kname = ('fce_fwd_kernel + fce_bwd_kernel (flash CE: logits recomputed in the backward, never stored): tcgen05.mma '
                 'kind::f16 on bf16 hi/lo pairs (3 products), TMA SWIZZLE_128B operands, TMEM accumulators, dZ fed back '
                 'to the tensor cores from shared memory')
    elif umma:
        sh, sl, Eh, El = (torch.empty_like(x) for x in (s, s, E, E))
        Zh, Zl = torch.empty_like(Z), torch.empty_like(Z)
        ops.split_tf32(s, d, B, d, sh, sl, d)
        ops.split_tf32(E, d, V, d, Eh, El, d)
        ops.split_tf32(Z, ldz, B, V, Zh, Zl, ldz)
        split = max(1, min((V + 31) // 32, 148 // ((B + 127) // 128)))

        lse, nll = torch.empty(B, device=device), torch.empty(B, device=device)
        part = torch.empty(4 * ((V + 255) // 256) * B + B, device=device)
        lab = torch.zeros(B, dtype=torch.int32, device=device)

        def run():
            # forward: persistent kernel, Z + fused row log-sum-exp / label logit; backward: dS (split-K) and dE
            ops.umma_score_fwd(B, V, d, sh, sl, d, Eh, El, d, Z, ldz, 12.0, lab, lse, nll, part)
            ops.umma_gemm(1, B, d, V, Zh, Zl, ldz, Eh, El, d, dS, d, accumulate=True, split_k=split)
            ops.umma_gemm(2, V, d, B, Zh, Zl, ldz, sh, sl, d, dE, d)
        kname = ('umma_score_fwd_kernel (persistent, fused LSE) + umma_gemm_kernel x2: tcgen05.mma kind::tf32 (3xTF32 split), '
                 'TMA SWIZZLE_128B operands, TMEM accumulators')
    else:
        def run():
            ops.gemm(B, V, d, s, d, 1, E, 1, d, Z, ldz, alpha=12.0)
            ops.gemm(B, d, V, Z, ldz, 1, E, d, 1, dS, d, accumulate=True, split_k=0)
            ops.gemm(V, d, B, Z, 1, ldz, s, d, 1, dE, d, accumulate=True, split_k=0)
        kname = 'sgemm_kernel: fp32 FFMA'
    for _ in range(3):
        run()
    times = []
    for _ in range(10):
        flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = float(np.median(times))
    flops = 6.0 * B * V * d
    ach = flops / (ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    if flash:
        # dram__bytes_read.sum + dram__bytes_write.sum of the forward + backward launch from the committed `ncu --set full`
        # capture; only valid for the shape a capture was taken on (narrow kernels: cfg1 shape, wide kernels: cfg2 shape)
        caps = sorted((ROOT / 'profiles').glob('*_ncu_full_flash_ce.json'))
        wide = {(512, 43097, 96): False, (2048, 17000, 256): True}.get((B, V, d))
        if caps and wide is not None:
            rows = [r for r in json.loads(caps[-1].read_text()) if ('wide' in r['kernel']) == wide]
            if len(rows) == 2:
                traffic = round(sum(float(r['dram_bytes_read']) + float(r['dram_bytes_write']) for r in rows))
                traffic_src = (f'profiles/{caps[-1].name} (bytes of the forward + backward launch; tensor pipe active '
                               + ' / '.join(f"{r['tensor_pipe_pct_active']:.1f} %" for r in rows) + ')')
    elif umma and (B, V, d) == (512, 43097, 96):
        # dram__bytes_read.sum + dram__bytes_write.sum of the same three launches from the committed `ncu --set full`
        # capture (profiles/); only valid for the shape it was captured on
        caps = sorted((ROOT / 'profiles').glob('*_ncu_full_umma_gemm.json'))
        if caps:
            rows = json.loads(caps[-1].read_text())
            traffic = round(sum(float(r['dram__bytes_read.sum'][0]) + float(r['dram__bytes_write.sum'][0]) for r in rows) * 1e6)
            traffic_src = f'profiles/{caps[-1].name} (bytes for the 3 launches; algorithmic operand bytes ~ 0.47 GB with the hi/lo split)'
    return dict(bound='tensor', kernel='catalog scoring GEMMs (Z = s E^T, dS = dZ E, dE = dZ^T s): ' + kname,
                achieved=round(ach, 3), peak=pk['tf_burst'], unit='TFLOP/s', frac=round(ach / pk['tf_burst'], 5),
                traffic=traffic, traffic_source=traffic_src, ms_for_the_head=round(ms, 4), ms_per_kernel=per_kernel,
                algorithmic_flops=flops,
                note='algorithmic FLOPs 6*B*V*d (Z = s E^T, dS = dZ E, dE = dZ^T s) counted once; the 3 passes of the hi/lo '
                     'split and the logit recomputation in the backward are the kernels\' own cost.  Timed here (and captured by '
                     'ncu) is the public srk_flash_ce_bwd, which writes one partial table gradient per 128-session tile; the '
                     'native steps run the same kernel in its accumulating form (one L2-resident [V, d] buffer, no partial tables)',
                peak_source=f"{pk['src']} bf16 burst (kernel timed alone)")


def gather_scatter_probe(device, pk, shapes=None):
    """HBM roofline of the embedding gather (K1) and the gradient scatter-add (K8) kernels, timed alone with CUDA
    events: the BASELINE config-2 shape (table L2-resident) and a stress shape whose table is far larger than L2.
    Algorithmic bytes: gather N*(4 + 2*4d), scatter N*(4 + 4d) + 2*U*4d (int32 indices)."""
    from sessionrec_pytorch_b200 import ops
    out = []
    shapes = shapes or [('cfg2 shape: V=17000 d=256 N=9000 (table 17 MB, L2-resident)', 17000, 256, 9000),
                        ('stress: V=4M d=256 N=1M (table 4.1 GB >> 126 MB L2)', 4_000_000, 256, 1_000_000)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    for name, V, d, N in shapes:
        g = torch.Generator().manual_seed(7)
        E = torch.empty(V, d, device=device).normal_()
        iid_h = torch.randint(0, V, (N,), generator=g, dtype=torch.int32)
        order = torch.argsort(iid_h.long(), stable=True).int()
        uid_h, cnt = torch.unique(iid_h.long(), return_counts=True)
        uoff_h = torch.zeros(uid_h.numel() + 1, dtype=torch.int32)
        uoff_h[1:] = torch.cumsum(cnt, 0).int()
        t = dict(iid=iid_h.to(device), perm=order.to(device), uoff=uoff_h.to(device), uid=uid_h.int().to(device),
                 U=int(uid_h.numel()), P=N)
        X = torch.empty(N, d, device=device)
        rn = torch.empty(N, device=device)
        dE = torch.zeros(V, d, device=device)

        def timeit(fn):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(7):
                flush.fill_(0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return float(np.median(ts))
        ms_g = timeit(lambda: ops.embed_gather_fwd(E, t['iid'], N, d, 2, None, X, rn))
        ws = torch.empty(max(1, ops.embed_scatter_ws_floats(N, d)), device=device)
        ms_s = timeit(lambda: ops.embed_scatter_bwd(E, t, d, 2, None, rn, X, None, dE, ws=ws))
        ms_p = timeit(lambda: ops.embed_scatter_bwd(E, t, d, 0, None, None, X, None, dE, ws=ws))
        bg = N * (4 + 2 * 4 * d)
        bp = N * (4 + 4 * d) + 2 * t['U'] * 4 * d           # plain scatter-add (SRGNN): BASELINE.md section 4
        bs = bp + t['U'] * 4 * d                             # + one table row per distinct item for the normalise-backward
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the
        # stress shape (profiles/*_ncu_full_gather_scatter.json); the L2-resident shape has no capture
        caps = sorted((ROOT / 'profiles').glob('*_ncu_full_gather_scatter.json'))
        cap = json.loads(caps[-1].read_text()) if caps and name.startswith('stress') else []
        for kname, ms, by in (('gather_tma_kernel (K1: TMA bulk-staged gather + L2 normalise)', ms_g, bg),
                              ('scatter_bwd_kernel (K8: normalise-backward + deterministic scatter-add; reads one table row per distinct item)', ms_s, bs),
                              ('scatter_bwd_kernel (K8 plain: deterministic scatter-add, SRGNN)', ms_p, bp)):
            ach = by / (ms * 1e-3) / 1e9
            tr = [round(r['dram_bytes_read'] + r['dram_bytes_write']) for r in cap
                  if kname.split(' ')[0][:-7] in r['kernel'] and 'plain' not in kname]
            out.append(dict(bound='hbm', kernel=kname, shape=name, achieved=round(ach, 1), peak=pk['hbm'], unit='GB/s',
                            frac=round(ach / pk['hbm'], 4), ms=round(ms, 4), algorithmic_bytes=by,
                            traffic=tr[0] if tr else None, traffic_source=f'profiles/{caps[-1].name}' if tr else None))
        del E, X, dE
        torch.cuda.empty_cache()
    return out


def builder_probe(pkg, cfg, n=40):
    """Host-side cost of the native batch builder (`srk_batch_build`, the collate_fn replacement, collate.py:219-256) on
    one thread: the end-to-end loop above starts from built, pinned batches (a DataLoader worker's output), so this says
    whether one worker keeps up with the device step.  Never fatal: returns None on any problem."""
    try:
        from sessionrec_pytorch_b200.synthetic import SessionSampler
        smp = SessionSampler(cfg['V'], seed=321)
        raw = [smp.batch(cfg['B']) for _ in range(n)]
        pkg.SessionBatch.build_flat(*raw[0], kind_of(cfg), cfg['order'], pin=False)
        t0 = time.perf_counter()
        for items, offs, labels in raw:
            pkg.SessionBatch.build_flat(items, offs, labels, kind_of(cfg), cfg['order'], pin=False)
        dt = (time.perf_counter() - t0) / n
        return dict(ms_per_batch=round(1e3 * dt, 4), sessions_per_s=round(cfg['B'] / dt, 1), threads=1,
                    what='SessionBatch.build_flat: flat item ids -> graph batch buffer (host, not inside any timed region)')
    except Exception as e:                                        # noqa: BLE001
        return dict(error=f'{type(e).__name__}: {e}')


def reference_arm(args, cfg, rank, guard):
    if rank != 0:
        return
    cb = cpu_arm(cfg, args.workload, args.steps, args.warmup)
    v = cb['value']
    guard.emit({
        'impl': 'reference', 'metric': METRIC, 'value': round(v, 2), 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(1e3 * cfg['B'] / v, 3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(workload=workload_name(args.workload, cfg), **{k: cfg[k] for k in ('V', 'd', 'B', 'order', 'layers', 'dropout')}),
        'cpu_baseline': cb,
        'e2e': dict(value=round(v, 2), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        'note': cb['what']})


def graph_counters():
    """CUDA-graph replay statistics of the native steps in this process (whole run, not only the timed region)."""
    from sessionrec_pytorch_b200._lib import lib
    f = lib().functions
    return dict(steps_replayed=int(f['srk_graph_launches']()), fallbacks=int(f['srk_graph_fallbacks']()),
                nodes_rewritten=int(f['srk_graph_node_updates']()))


def workload_name(key, cfg):
    head = f"BASELINE.json configs[{int(key[3:])}]" if key[3:].isdigit() else f"{key} (not a BASELINE config)"
    return (f"{head}: {cfg['model']} order {cfg['order']}, {cfg['layers']} layer(s), {cfg['note']} "
            f"V={cfg['V']} d={cfg['d']} B={cfg['B']}")


class StdoutGuard:
    """Library chatter on fd 1 (NCCL prints its version banner there) must not precede the ONE JSON line: everything
    written to stdout while the guard is active goes to stderr; `emit` restores fd 1 and prints the line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)


def measure(key, cfg, args, pkg, device, group, world, rank, pk, primary):
    """Device-timed loop, end-to-end loop, roofline and CPU arm of ONE workload; returns the JSON object of that workload
    (None on the ranks that do not print).  primary: also sample the clocks over a continuation, time the native batch
    builder and the gather / scatter kernels."""
    from sessionrec_pytorch_b200 import ops
    from sessionrec_pytorch_b200.synthetic import SessionSampler
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shard = args.parallelism == 'shard' and world > 1
    model = build_model(cfg, device)
    model.configure_optimizer(lr=1e-3, weight_decay=1e-4)
    # The library's own NCCL communicator (collectives enqueued inside the native step): always for the catalog-sharded
    # step; for data parallelism only with SESSREC_NATIVE_COMM=1 - measured equal to the torch.distributed all-reduce at 2
    # GPUs, but back-to-back steps stall on it at 8 GPUs (profiles/r2o_*_8gpu.json), so torch.distributed stays the default
    native_comm = False
    if world > 1 and (shard or os.environ.get('SESSREC_NATIVE_COMM', '0') == '1'):
        from sessionrec_pytorch_b200 import parallel
        parallel.init_comm(group)
        native_comm = True
        model.dp_allreduce_inside = True
    if shard:
        model.shard_catalog(group)
    step_group = None if shard else group
    gb = None if (shard or world == 1) else world * cfg['B']      # equal shards: B_global is known on the host
    # data parallel: every rank draws its own sessions (weak scaling); catalog sharding: every rank sees the SAME global
    # batch and scores its slice of the catalog (strong scaling)
    smp = SessionSampler(cfg['V'], seed=123 + (0 if shard else rank))
    n_batches = int(os.environ.get('SESSREC_BENCH_BATCHES', '8'))      # experiment knob: 1 = the same batch every step
    host = []
    for _ in range(n_batches):
        items, offs, labels = smp.batch(cfg['B'])
        host.append(pkg.SessionBatch.build_flat(items, offs, labels, kind_of(cfg), cfg['order'], pin=True))
    resident = [b.to(device, non_blocking=False) for b in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)        # 2x the 126 MB L2

    # ---- device-timed region: K steps, inputs resident in HBM -------------------------------------------------
    for i in range(args.warmup):
        model.train_step(resident[i % n_batches], step_group, gb)
    barrier()
    clocks = ClockSampler(device.index)
    clocks.start()
    evs = []
    launches = 0
    enqueue_s = 0.0
    t_wall0 = time.perf_counter()
    nat0 = list(ops.native_call_s)
    for i in range(args.steps):
        flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kk = ops.kernel_launches()
        a.record()
        tq = time.perf_counter()
        model.train_step(resident[(args.warmup + i) % n_batches], step_group, gb)
        enqueue_s += time.perf_counter() - tq
        b.record()
        launches += ops.kernel_launches() - kk
        evs.append((a, b))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if ops.HOST_PROFILE and rank == 0:
        print(f'[host profile] train_step {1e3 * enqueue_s / args.steps:.4f} ms/step, of which inside the native C call '
              f'{1e3 * (ops.native_call_s[0] - nat0[0]) / max(1, ops.native_call_s[1] - nat0[1]):.4f} ms', file=sys.stderr)
    clk = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    units = cfg['B'] if shard else world * cfg['B']              # sessions all ranks processed per step
    value = units * args.steps / (total_ms * 1e-3)
    # The timed region lasts a few milliseconds, shorter than one nvidia-smi sampling period: take the clocks line again
    # over an untimed continuation of the same step loop (single process only - no collective may depend on it).
    if world == 1 and primary:
        try:
            more = ClockSampler(device.index)
            more.start()
            for i in range(1500):
                model.train_step(resident[i % n_batches])
                if i % 100 == 99:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            m = more.stop()
            if (m.get('samples') or 0) > (clk.get('samples') or 0):
                m['window'] = ('untimed continuation of the timed loop (1500 more steps of the same workload): the timed region '
                               'itself is shorter than one sampling period')
                m['in_timed_region'] = clk
                clk = m
        except Exception as e:                                  # noqa: BLE001 - keep the first reading
            clk['continuation_error'] = f'{type(e).__name__}: {e}'

    # ---- end-to-end through the public API with host buffers -----------------------------------------------------
    for i in range(min(3, args.warmup)):                 # untimed warm-up of THIS path: first pinned H2D, first D2H read of the loss
        model.train_step(host[i % n_batches].to(device, non_blocking=True), step_group, gb).item()
    barrier()
    t0 = time.perf_counter()
    h2d = 0
    for i in range(args.steps):
        hb = host[i % n_batches]
        db = hb.to(device, non_blocking=True)           # H2D of the whole batch (one pinned buffer)
        loss = model.train_step(db, step_group, gb)
        loss.item()                                     # D2H read of the step's loss (what TrainRunner does, train.py:103)
        h2d += hb.nbytes
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = units * args.steps / float(e2e_s)

    # ---- the same loop with the loss read one step late (asynchronous D2H into pinned memory): what the package's own
    # TrainRunner does between log lines.  Reported beside `e2e`, never instead of it.
    e2e_lagged = None
    if world == 1 and primary:
        try:
            slots = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
            evs2 = [torch.cuda.Event() for _ in range(2)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            acc = 0.0
            for i in range(args.steps):
                db = host[i % n_batches].to(device, non_blocking=True)
                loss = model.train_step(db)
                if i > 0:
                    evs2[(i - 1) & 1].synchronize()
                    acc += float(slots[(i - 1) & 1])                 # loss of the PREVIOUS step
                slots[i & 1].copy_(loss, non_blocking=True)
                evs2[i & 1].record()
            evs2[(args.steps - 1) & 1].synchronize()
            acc += float(slots[(args.steps - 1) & 1])
            torch.cuda.synchronize()
            e2e_lagged = dict(value=round(cfg['B'] * args.steps / (time.perf_counter() - t0), 1), unit=UNIT, mean_loss=round(acc / args.steps, 5),
                              timing='wall clock over K steps: H2D batch copy + train_step per step, every step\'s loss copied D2H '
                                     'asynchronously and read one step later')
        except Exception as e:                                  # noqa: BLE001 - a secondary figure, never fatal
            e2e_lagged = dict(error=f'{type(e).__name__}: {e}')
            torch.cuda.synchronize()

    # ---- the same loop fed by the native batch builder (raw clicks -> graph batch on one background thread) --------
    e2e_build = None
    if world == 1 and primary:          # single process only: no collective inside, so a failure here cannot desynchronise ranks
        try:
            from sessionrec_pytorch_b200.loader import BatchPrefetcher
            raw_smp = SessionSampler(cfg['V'], seed=777)
            raw = [raw_smp.batch(cfg['B']) for _ in range(n_batches)]
            it = BatchPrefetcher((raw[i % n_batches] for i in range(args.steps + 4)), kind_of(cfg), cfg['order'], device=device,
                                 depth=3, timeout=30.0)
            for _ in range(4):                              # every slot of the ring has made its pinned allocation
                model.train_step(next(it)).item()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_done = 0
            for db in it:
                model.train_step(db).item()
                n_done += 1
            torch.cuda.synchronize()
            eb_s = time.perf_counter() - t0
            e2e_build = dict(value=round(cfg['B'] * n_done / eb_s, 1), unit=UNIT, steps=n_done, builder_threads=1,
                             timing='wall clock: native batch build (one background thread, ring of pinned buffers) + H2D + '
                                    'train_step + loss.item() per step')
        except Exception as e:                                  # noqa: BLE001 - a secondary figure, never fatal
            e2e_build = dict(error=f'{type(e).__name__}: {e}')
            torch.cuda.synchronize()

    out = None
    if rank == 0:
        par = f'catalog-sharded x{world} (session encoder replicated, every rank scores V/{world} rows)' if shard else f'dp{world}'
        out = {
            'metric': METRIC, 'value': round(value, 1), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(total_ms / args.steps, 4), 'higher_is_better': True,
            'scaling': 'strong' if shard else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'dtype_note': 'fp32 parameters, activations and accumulation; each fp32 product of the catalog head runs as 3 bf16 '
                          'tcgen05 MMAs (hi/lo split, ~1e-5 relative against fp64; parity bar 1e-4)',
            'config': dict(workload=workload_name(key, cfg), parallelism=par, global_batch=units,
                           l2='flushed between timed steps (256 MB write)', timing='CUDA events per step, max over ranks',
                           **{k: cfg[k] for k in ('V', 'd', 'B', 'order', 'layers', 'dropout')}),
            'clocks': clk, 'gpu_launches': int(launches), 'launches_per_step': round(launches / args.steps, 1),
            'e2e': dict(value=round(e2e, 1), unit=UNIT, h2d_bytes_per_step=int(h2d / args.steps), d2h_bytes_per_step=4,
                        timing='wall clock over K steps incl. H2D batch copy + loss.item() per step'),
            'wall_s_timed_region': round(t_wall, 4), 'host_enqueue_ms_per_step': round(1e3 * enqueue_s / args.steps, 4),
            'graph': graph_counters(),
            'roofline': roofline_probe(cfg, device, pk),
        }
        if world > 1:
            nflat = int(model._flat.data.numel())
            where = ('enqueued by the native step itself on the library\'s own NCCL communicator (csrc/comm.cu), inside its CUDA-graph replay'
                     if native_comm else 'torch.distributed (NCCL) between the two halves of the native step')
            if shard:
                out['collectives'] = dict(per_step=[f'ncclAllReduce(sum) [2, B={cfg["B"]}] fp32: per-session sum exp(logit - 12) + label logit',
                                                    f'ncclAllReduce(sum) dS [B={cfg["B"]}, d={cfg["d"]}] fp32',
                                                    f'2 x ncclAllReduce(avg) of the replicated (non-table) gradients, {nflat - cfg["V"] * cfg["d"]} floats',
                                                    f'grouped ncclBroadcast of the owners\' updated table rows, V*d = {cfg["V"] * cfg["d"]} floats'],
                                          where=where, table_gradient_allreduce=False)
            else:
                out['collectives'] = dict(per_step=[f'ncclAllReduce(sum) of the flat fp32 gradient buffer, {nflat} floats'], where=where)
        if primary:
            out['roofline_gather_scatter'] = gather_scatter_probe(device, pk) if not args.no_gather_probe else None
            out['batch_builder'] = builder_probe(pkg, cfg) if world == 1 else None
            out['e2e_with_batch_build'] = e2e_build
            out['e2e_loss_read_one_step_late'] = e2e_lagged
    del model, resident, flush
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        # bounded sample: ~10 s of CPU work, sized from the reference's indicative step times
        guess_s = {'cfg1': 0.16, 'cfg2': 0.13, 'cfg3': 0.03, 'cfg4': 0.35, 'cfg1k3': 0.8}.get(key, 0.2)
        nst = max(3, min(120, int(10.0 / guess_s)))
        out['cpu_baseline'] = cpu_arm(cfg, key, nst, 1)
    return out


def main():
    guard = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--workload', default=None, help='cfg1 (default; cfg4 with --parallelism shard)')
    ap.add_argument('--also', default=None, help='comma list of further workloads measured into the `workloads` block '
                                                 '(default: cfg2 on a single-GPU run of cfg1)')
    ap.add_argument('--parallelism', default='dp', choices=['dp', 'shard'],
                    help='N > 1: dp = data parallel, weak scaling (default); shard = item-catalog rows sharded across the ranks, '
                         'strong scaling on BASELINE configs[4]')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--secondary', action='store_true', help='internal: this process measures one workload of the `workloads` block')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gather-probe', action='store_true')
    args = ap.parse_args()
    if args.workload is None:
        args.workload = 'cfg4' if args.parallelism == 'shard' else 'cfg1'

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    load_package()
    from sessionrec_pytorch_b200.synthetic import CONFIGS
    cfg = dict(CONFIGS[args.workload])
    if args.impl == 'reference':
        if args.steps == 30 and args.warmup == 5:
            args.steps, args.warmup = 5, 1
        reference_arm(args, cfg, rank, guard)
        return
    assert args.warmup >= 3, 'timing rules: at least 3 warm-up steps'
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback exists for the product path)'
    build()
    pkg = load_package()
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
        group = dist.group.WORLD
    pk = peaks()
    also = args.also
    if also is None:
        also = 'cfg2,cfg4,cfg1k3' if (world == 1 and args.workload == 'cfg1') else ''
    out = measure(args.workload, cfg, args, pkg, device, group, world, rank, pk, not args.secondary)
    extra = {}
    for key in [k for k in also.split(',') if k and k != args.workload]:
        if world > 1:
            extra[key] = measure(key, dict(CONFIGS[key]), args, pkg, device, group, world, rank, pk, False)
            continue
        # single GPU: every further workload runs in its own process (a fresh allocator, no host threads left over from the
        # CPU arm of the previous workload, and a failure there never costs the headline line)
        cmd = [sys.executable, str(ROOT / 'bench.py'), '--workload', key, '--also', '', '--secondary', '--steps', str(args.steps),
               '--warmup', str(args.warmup), '--no-gather-probe'] + (['--no-cpu-baseline'] if args.no_cpu_baseline else [])
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
            lines = [ln for ln in p.stdout.splitlines() if ln.strip().startswith('{')]
            extra[key] = json.loads(lines[-1]) if (p.returncode == 0 and lines) else dict(error=f'exit {p.returncode}: {p.stderr[-300:]}')
        except Exception as e:                                  # noqa: BLE001
            extra[key] = dict(error=f'{type(e).__name__}: {e}')
    if rank == 0:
        if extra:
            out['headline_workload'] = args.workload
            out['workloads'] = extra
        guard.emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
