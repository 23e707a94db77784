"""LossTracker — running means of the training losses between two prints (mirror of
utils/loss_tracker.py:11-82).  Values ``<= 0`` are skipped exactly like the reference's ``if loss_value > 0``
(:41-43), so the regularisation penalties — identically 0.0 on the steps that do not compute them
(training_step.py:115,261,294) — average over regularised steps only."""
from __future__ import annotations

from time import time
from typing import Dict, List, Optional


class _Mean:
    """tf.keras.metrics.Mean for scalars: ``m(value)`` accumulates, ``result()`` is the mean (0.0 when empty)."""

    def __init__(self, name: str):
        self.name = name
        self.total = 0.0
        self.count = 0

    def __call__(self, value) -> None:
        self.total += float(value)
        self.count += 1

    def result(self) -> float:
        return self.total / self.count if self.count else 0.0


class LossTracker:
    """Tracks the different losses to monitor the performance of the model."""

    def __init__(self, loss_names: List[str], print_step: Optional[int] = None, log_losses: Optional[bool] = None,
                 num_replicas: int = 1, printer=print):
        self.print_step = print_step
        self.log_losses = log_losses
        self.loss_names = loss_names
        self.num_replicas = num_replicas
        self._print = printer
        self._initiate_loss_tracking()

    def _initiate_loss_tracking(self) -> None:
        self.losses: Dict[str, _Mean] = {n: _Mean(n) for n in self.loss_names}
        self.timer = _Mean("timer")
        self.start_time = time()

    def increment_losses(self, losses: dict) -> None:
        """utils/loss_tracker.py:32-46.  ``float(loss) > 0`` reads a device scalar: one host sync per tracked
        value, as in the reference (``.numpy()`` behind ``if loss_value > 0``)."""
        for loss_name, loss_value in losses.items():
            v = float(loss_value)
            if v > 0:
                self.losses[loss_name](v)
        self.timer(time() - self.start_time)
        self.start_time = time()

    def print_losses(self, step) -> str:
        """utils/loss_tracker.py:48-79"""
        start_print = "Step: {}. Avg over the last {:d} steps. {:.2f} s/step. Losses:".format(
            step, int(self.timer.count / self.num_replicas), self.timer.result())
        loss_print = ", ".join("- {:s}: {:.4f}".format(n, self.losses[n].result()) for n in self.loss_names)
        line = start_print + loss_print
        self._print(line)
        return line

    def reinitialize_tracker(self) -> None:
        self._initiate_loss_tracking()
