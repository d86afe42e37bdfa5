"""BatchPrefetcher: the collate side of the reference's `DataLoader(..., collate_fn=collate_fn, num_workers=...)`
(`src/scripts/main_*.py`, `utils/data/collate.py:219-256`) for the native builder.  One background thread turns raw
click arrays into `SessionBatch` buffers inside a small ring of reused (pinned) host buffers while the device works on
the previous batch; the consumer thread only issues the one H2D copy per batch.  The builder call releases the GIL
(ctypes), so the two threads really overlap."""
import queue
import threading

import torch

from .batch import SessionBatch

_END = object()


class _Slot:
    __slots__ = ('buf', 'event')

    def __init__(self):
        self.buf, self.event = None, None


class BatchPrefetcher:
    """for batch in BatchPrefetcher(source, kind, order, device='cuda'): ...

    source: iterable of (items int32[T], offs int32[B + 1], labels int32[B]) triples (what `SessionBatch.build_flat`
    takes).  device=None yields host batches that stay valid until the next one is requested; with a CUDA device the
    batch is copied (non-blocking, current stream) and the host slot is reused once that copy has completed."""

    def __init__(self, source, kind='session', order=1, device=None, depth=3, pin=None, timeout=60.0):
        if depth < 2:
            raise ValueError('BatchPrefetcher needs depth >= 2 (one slot being built, one being consumed)')
        self.kind, self.order, self.timeout = kind, order, timeout
        self.device = None if device is None else torch.device(device)
        self.pin = (self.device is not None and self.device.type == 'cuda') if pin is None else pin
        self._free, self._ready = queue.Queue(), queue.Queue()
        for _ in range(depth):
            self._free.put(_Slot())
        self._held = None
        self._stop = False
        self._thread = threading.Thread(target=self._work, args=(iter(source),), daemon=True, name='sessrec-batch-builder')
        self._thread.start()

    # ---- builder thread ---------------------------------------------------------------------------------------------
    def _work(self, it):
        try:
            for items, offs, labels in it:
                slot = self._free.get()
                if slot is _END or self._stop:
                    return
                if slot.event is not None:                       # the H2D copy that read this slot last
                    slot.event.synchronize()
                    slot.event = None
                need = SessionBatch.batch_words(int(offs[-1]) - int(offs[0]), len(offs) - 1, self.kind, self.order)
                if slot.buf is None or slot.buf.numel() < need:
                    slot.buf = torch.empty(int(need * 1.25), dtype=torch.int32, pin_memory=self.pin)
                batch = SessionBatch.build_flat(items, offs, labels, self.kind, self.order, out=slot.buf)
                self._ready.put((slot, batch))
            self._ready.put((None, _END))
        except BaseException as e:                               # noqa: BLE001 - forwarded to the consumer
            self._ready.put((None, e))

    # ---- consumer -----------------------------------------------------------------------------------------------------
    def __iter__(self):
        return self

    def __next__(self):
        if self._held is not None:                               # host mode: the previous batch is no longer in use
            self._free.put(self._held)
            self._held = None
        try:
            slot, batch = self._ready.get(timeout=self.timeout)
        except queue.Empty:
            raise RuntimeError(f'BatchPrefetcher: no batch within {self.timeout} s (builder thread stuck?)') from None
        if batch is _END:
            self._ready.put((None, _END))                        # stay exhausted
            raise StopIteration
        if isinstance(batch, BaseException):
            self._ready.put((None, batch))
            raise batch
        if self.device is None:
            self._held = slot
            return batch
        dev = batch.to(self.device, non_blocking=True)
        if self.device.type == 'cuda':
            slot.event = torch.cuda.Event()
            slot.event.record()
        self._free.put(slot)
        return dev

    def close(self):
        self._stop = True
        self._free.put(_END)

    def __del__(self):
        try:
            self.close()
        except Exception:                                        # noqa: BLE001
            pass


class EpochBatches:
    """Every batch of one pass over a dataset, built ONCE into one host buffer (optionally pinned) by a few threads, and
    movable to the device in ONE copy: the batches of the pass are then contiguous slices of device memory and the training
    loop does no per-step collate and no per-step H2D at all.  The reference's train loaders for SRGNN / MSGIFSR sample
    sequentially (`main_msgifsr.py:156`), so the pass is the same every epoch; for a shuffled loader (`main_niser.py:86`)
    build the next epoch's pass in the background from the permuted sample arrays.

    items int32[T] / offs int32[n + 1] / labels int32[n]: all samples back to back (see `flatten_samples`)."""

    ALIGN = 64                       # words: every batch starts on a 256-byte boundary of the buffer

    def __init__(self, items, offs, labels, batch_size, kind='session', order=1, threads=4, pin=False, drop_last=False):
        import numpy as np
        from concurrent.futures import ThreadPoolExecutor
        items = np.ascontiguousarray(items, np.int32)
        offs = np.ascontiguousarray(offs, np.int32)
        labels = np.ascontiguousarray(labels, np.int32)
        n = len(offs) - 1
        assert len(labels) == n and batch_size >= 1
        bounds = [(lo, min(lo + batch_size, n)) for lo in range(0, n, batch_size)]
        if drop_last and bounds and bounds[-1][1] - bounds[-1][0] < batch_size:
            bounds.pop()
        starts, pos = [], 0
        for lo, hi in bounds:
            starts.append(pos)
            need = SessionBatch.batch_words(int(offs[hi]) - int(offs[lo]), hi - lo, kind, order)
            pos += (need + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.buf = torch.empty(max(pos, 1), dtype=torch.int32, pin_memory=pin)
        ends = starts[1:] + [pos]

        def one(i):
            lo, hi = bounds[i]
            # offs keeps its absolute click positions: the builder indexes `items` with them, no per-batch copy
            return SessionBatch.build_flat(items, offs[lo:hi + 1], labels[lo:hi], kind, order, out=self.buf[starts[i]:ends[i]])

        if threads > 1 and len(bounds) > 1:
            with ThreadPoolExecutor(max_workers=threads) as ex:
                built = list(ex.map(one, range(len(bounds))))
        else:
            built = [one(i) for i in range(len(bounds))]
        self._spans = [(s, s + b.buf.numel()) for s, b in zip(starts, built)]
        self._hdrs = [b.hdr for b in built]

    def __len__(self):
        return len(self._spans)

    def __getitem__(self, i):
        s, e = self._spans[i]
        return SessionBatch(self.buf[s:e], self._hdrs[i])

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    @property
    def nbytes(self):
        return self.buf.numel() * 4

    def to(self, device, non_blocking=True):
        out = object.__new__(EpochBatches)
        out.buf = self.buf.to(device, non_blocking=non_blocking)
        out._spans, out._hdrs = self._spans, self._hdrs
        return out


def flatten_samples(samples):
    """[(click sequence, label), ...] (e.g. an `AugmentedDataset`) -> (items int32[T], offs int32[n + 1], labels int32[n])."""
    import numpy as np
    n = len(samples)
    lens = np.fromiter((len(samples[i][0]) for i in range(n)), dtype=np.int64, count=n)
    offs = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=offs[1:])
    items = np.fromiter((int(x) for i in range(n) for x in samples[i][0]), dtype=np.int32, count=int(offs[-1]))
    labels = np.fromiter((int(samples[i][1]) for i in range(n)), dtype=np.int32, count=n)
    return items, offs, labels
