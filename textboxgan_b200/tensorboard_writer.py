"""TensorboardWriter — scalar and config logging of the training loop (mirror of utils/tensorboard_writer.py:12-44;
image summaries, :46-137, are not built).  Usable as ``Trainer(scalar_writer=TensorboardWriter(log_dir).log_scalars)``."""
from __future__ import annotations

import dataclasses
from typing import Optional

from .config import Config


class TensorboardWriter:
    """Log data related to the performance of the model on a file which can be visualised on tensorboard."""

    def __init__(self, log_dir: str, cfg: Optional[Config] = None):
        from torch.utils.tensorboard import SummaryWriter

        self.cfg = cfg
        self.train_summary_writer = SummaryWriter(log_dir=log_dir)

    def log_scalars(self, loss_dict: dict, step: int) -> None:
        """utils/tensorboard_writer.py:24-37.  Values may be running means (``.result()``) or plain numbers."""
        for loss_name, metric in loss_dict.items():
            value = metric.result() if hasattr(metric, "result") else metric
            self.train_summary_writer.add_scalar(loss_name, float(value), global_step=step)
        self.train_summary_writer.flush()

    def log_config_file(self, step: int) -> None:
        """utils/tensorboard_writer.py:39-44 — the configuration as a two-column text table."""
        if self.cfg is None:
            return
        rows = "\n".join(f"| {f.name} | {getattr(self.cfg, f.name)} |" for f in dataclasses.fields(self.cfg)
                         if f.name != "strategy")
        self.train_summary_writer.add_text("configs", "| key | value |\n|---|---|\n" + rows, global_step=step)
        self.train_summary_writer.flush()
