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
