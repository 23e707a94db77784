"""One-process-per-GPU stand-in for ``tf.distribute.MirroredStrategy`` (config/config.py:140).

The reference drives all replicas from one Python thread; here every rank runs the same script
(``torchrun``) and this shim keeps the call sites of training_step.py / train.py unchanged:
``run`` calls the function on the local replica, ``reduce(SUM)`` is an all-reduce over NCCL (gloo
on CPU), ``experimental_distribute_dataset`` shards each global batch on axis 0.
"""
from __future__ import annotations

import contextlib
import os
from typing import Any, Callable, Iterable, Iterator, Sequence

import torch
import torch.distributed as dist


class ReduceOp:
    SUM = "SUM"
    MEAN = "MEAN"


class Strategy:
    def __init__(self, init_process_group: bool = True, backend: str | None = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1 and init_process_group and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend)
        self.world_size = dist.get_world_size() if dist.is_initialized() else 1
        if dist.is_initialized():
            self.rank = dist.get_rank()

    @property
    def num_replicas_in_sync(self) -> int:
        return self.world_size

    @contextlib.contextmanager
    def scope(self):
        yield self

    def run(self, fn: Callable, args: Sequence[Any] = (), kwargs: dict | None = None):
        return fn(*args, **(kwargs or {}))

    def reduce(self, reduce_op: str, value, axis=None):
        t = value if torch.is_tensor(value) else torch.as_tensor(value, dtype=torch.float32)
        if self.world_size > 1:
            t = t.detach().clone()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            if reduce_op == ReduceOp.MEAN:
                t = t / self.world_size
        return t

    def reduce_many(self, values: Sequence[torch.Tensor]) -> list:
        """SUM-reduce several scalars with ONE packed all-reduce (the reference issues seven,
        training_step.py:106-134)."""
        if self.world_size == 1:
            return [v.detach() for v in values]
        packed = torch.stack([v.detach().float().reshape(()) for v in values])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        return list(packed.unbind(0))

    def experimental_distribute_dataset(self, dataset: Iterable) -> Iterator:
        """Each element is a tuple of global-batch tensors; yield this rank's contiguous shard."""
        for batch in dataset:
            yield tuple(self.shard(t) for t in batch)

    def shard(self, t):
        if not torch.is_tensor(t) or t.dim() == 0 or self.world_size == 1:
            return t
        per = t.shape[0] // self.world_size
        return t[self.rank * per: (self.rank + 1) * per]

    def experimental_local_results(self, value):
        return (value,)
