"""Running means of the training losses between two prints — the service utils/loss_tracker.py:11-82 gives train.py.

Semantics kept from the reference:
* a value is accumulated only when it is strictly positive (:41-43): the regularisation penalties are identically
  0.0 on the steps that do not compute them (training_step.py:115,261,294), so their mean runs over regularised steps;
* every ``increment_losses`` call also records the wall time since the previous call; ``print_losses`` reports the
  mean step time and the number of steps (per replica) since the last ``reinitialize_tracker``;
* the printed line has the reference's exact format.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Sequence


@dataclass
class RunningMean:
    """Scalar stand-in for ``tf.keras.metrics.Mean``: call to add a value, ``result()`` is 0.0 while empty."""

    name: str
    total: float = 0.0
    count: int = 0

    def __call__(self, value) -> None:
        self.count += 1
        self.total += float(value)

    def result(self) -> float:
        return self.total / self.count if self.count else 0.0


class LossTracker:
    def __init__(self, loss_names: Sequence[str], print_step: Optional[int] = None, log_losses: Optional[bool] = None,
                 num_replicas: int = 1, printer: Callable[[str], None] = print):
        self.loss_names = list(loss_names)
        self.print_step, self.log_losses = print_step, log_losses
        self.num_replicas = max(1, int(num_replicas))
        self._emit = printer
        self.reinitialize_tracker()

    def reinitialize_tracker(self) -> None:
        self.losses: Dict[str, RunningMean] = {name: RunningMean(name) for name in self.loss_names}
        self.timer = RunningMean("timer")
        self._last = time.time()

    def increment_losses(self, losses: dict) -> None:
        """``float(v)`` on a device scalar is one host read per tracked loss, like the reference's ``loss_value > 0``."""
        for name, value in losses.items():
            value = float(value)
            if value > 0.0:
                self.losses[name](value)
        now = time.time()
        self.timer(now - self._last)
        self._last = now

    def print_losses(self, step) -> str:
        n_steps = self.timer.count // self.num_replicas
        head = f"Step: {step}. Avg over the last {n_steps:d} steps. {self.timer.result():.2f} s/step. Losses:"
        body = ", ".join(f"- {name}: {self.losses[name].result():.4f}" for name in self.loss_names)
        self._emit(head + body)
        return head + body
