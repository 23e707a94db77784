"""Host -> device input pipeline stage: the copy of batch i+1 overlaps iteration i.

The reference's loop hands `tf.data` batches to the step (train.py:174-194) and TensorFlow's runtime prefetches them
to the device; here the equivalent is explicit: batches from the (host-side) loader are copied with ``non_blocking``
transfers on a dedicated copy stream, from pinned memory when the loader provides it, one batch ahead of the consumer.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional

import torch


class DevicePrefetcher:
    """Iterator over ``batches`` (tuples of tensors / scalars) whose tensors arrive on ``device`` one batch ahead.

    ``pin``: stage pageable host tensors through pinned memory first (a pageable source would make the copy synchronous).
    On a CPU ``device`` it is a plain pass-through."""

    def __init__(self, batches: Iterable, device, pin: bool = True):
        self.it: Iterator = iter(batches)
        self.device = torch.device(device)
        self.pin = bool(pin)
        self.cuda = self.device.type == "cuda"
        self.stream: Optional[torch.cuda.Stream] = torch.cuda.Stream(self.device) if self.cuda else None
        self._next = None
        self._event = None
        self._preload()

    def _to_device(self, t):
        if not torch.is_tensor(t) or t.device == self.device:
            return t
        if self.cuda and self.pin and t.device.type == "cpu" and not t.is_pinned():
            t = t.pin_memory()
        return t.to(self.device, non_blocking=True)

    def _preload(self) -> None:
        try:
            batch = next(self.it)
        except StopIteration:
            self._next = None
            return
        if not self.cuda:
            self._next = tuple(self._to_device(t) for t in batch)
            return
        with torch.cuda.stream(self.stream):
            self._next = tuple(self._to_device(t) for t in batch)
            self._event = torch.cuda.Event()
            self._event.record(self.stream)

    def __iter__(self) -> "DevicePrefetcher":
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        batch = self._next
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._event)
            for t in batch:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(cur)          # the copy stream's allocator must not reuse it while `cur` reads it
        self._preload()
        return batch
