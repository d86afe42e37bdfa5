"""Multi-GPU host logic (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in the CPU tests).

The reference is single-device; the path shards two ways:
  * sessions (data parallel): rank r owns a contiguous slice of every global batch; ONE all-reduce of the flat
    gradient buffer per step, the 1/world mean folded into the Adam kernel (SessRecModule.train_step(batch, group));
  * catalog rows: rank r owns rows [lo, hi) of the item table for the scoring head; per session only the partial
    soft-max statistics (and the label logit) cross the links: ONE all-reduce of a [B, 2] tensor when the logits are
    bounded (NISER / MSGIFSR: |z| <= scale), a max + sum pair otherwise; backward needs one all-reduce of dS [B, d].
Everything here is device-agnostic tensor plumbing so that it can be exercised on CPU with gloo."""
import ctypes

import torch
import torch.distributed as dist

_COMM = dict(ready=False, rank=0, world=1)


def init_comm(group=None, nccl_path=None):
    """Creates this process's own NCCL communicator inside libsessrec_b200.so (csrc/comm.cu) so that the collectives of a
    training step are enqueued by the native step itself, between its kernels.  Collective over `group` (default: WORLD):
    rank 0 draws the NCCL unique id, torch.distributed only carries those 128 bytes (bootstrap).  The CUDA device of this
    process must be current.  Idempotent; returns (rank, world)."""
    if _COMM['ready']:
        return _COMM['rank'], _COMM['world']
    from ._lib import lib
    L = lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    path = None if nccl_path is None else str(nccl_path).encode()
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        L.call('srk_comm_unique_id', buf, path)
    dev = torch.device('cuda', torch.cuda.current_device())
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = bytes(t.cpu().tolist())
    L.call('srk_comm_init', ctypes.create_string_buffer(raw, 128), rank, world, path)
    _COMM.update(ready=True, rank=rank, world=world)
    return rank, world


def comm_ready():
    return _COMM['ready']


def destroy_comm():
    if _COMM['ready']:
        from ._lib import lib
        lib().call('srk_comm_destroy')
        _COMM.update(ready=False, rank=0, world=1)


def shard_slice(n, rank, world):
    """Contiguous, balanced [lo, hi) of n units for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(seqs, labels, rank, world):
    """Data parallel: the slice of a global batch this rank trains on."""
    lo, hi = shard_slice(len(seqs), rank, world)
    return seqs[lo:hi], labels[lo:hi]


def allreduce_mean_grads(flat_grad, group=None, weight=1.0):
    """Sum the (optionally weighted) flat gradient over the ranks; returns the factor the optimizer must apply
    (1 / sum of weights) so that unequal shard sizes still give the global-batch mean."""
    w = torch.tensor([float(weight)], dtype=torch.float64, device=flat_grad.device)
    if weight != 1.0:
        flat_grad.mul_(weight)
    dist.all_reduce(flat_grad, group=group)
    dist.all_reduce(w, group=group)
    return 1.0 / float(w)


def combine_lse(local_max, local_sumexp, group=None, bound=None):
    """log-sum-exp over a catalog whose rows are sharded across ranks.

    local_max[b], local_sumexp[b] = sum_v exp(z[b, v] - local_max[b]) over this rank's rows.
    bound: if the logits are known to satisfy |z| <= bound (cosine heads, scale 12) the shift is a constant and a
    single SUM all-reduce suffices (local_max is then ignored and local_sumexp must be relative to `bound`)."""
    if bound is not None:
        s = local_sumexp.clone()
        dist.all_reduce(s, group=group)
        return bound + torch.log(s)
    m = local_max.clone()
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    s = local_sumexp * torch.exp(local_max - m)
    dist.all_reduce(s, group=group)
    return m + torch.log(s)


def sharded_ce(z_local, labels, lo, hi, group=None, bound=None):
    """Mean NLL and d loss / d z_local for logits whose columns [lo, hi) live on this rank.

    Communication: the soft-max statistics and the label logit ride in ONE [B, 2] SUM all-reduce when `bound` is
    given (otherwise one extra MAX all-reduce)."""
    B = z_local.shape[0]
    own = (labels >= lo) & (labels < hi)
    idx = (labels - lo).clamp(0, max(hi - lo - 1, 0))
    zl = torch.where(own, z_local.gather(1, idx.unsqueeze(1)).squeeze(1) if hi > lo else torch.zeros_like(labels, dtype=z_local.dtype),
                     torch.zeros((), dtype=z_local.dtype, device=z_local.device))
    if bound is not None:
        pack = torch.stack([torch.exp(z_local - bound).sum(1), zl], 1)
        dist.all_reduce(pack, group=group)
        lse, zlab = bound + torch.log(pack[:, 0]), pack[:, 1]
    else:
        m = z_local.max(1)[0] if hi > lo else torch.full((B,), -3e38, dtype=z_local.dtype, device=z_local.device)
        lse = combine_lse(m, torch.exp(z_local - m.unsqueeze(1)).sum(1), group)
        zlab = zl.clone()
        dist.all_reduce(zlab, group=group)
    loss = (lse - zlab).mean()
    dz = torch.exp(z_local - lse.unsqueeze(1))
    if hi > lo:
        dz[own, idx[own]] -= 1.0
    return loss, dz / B, lse
