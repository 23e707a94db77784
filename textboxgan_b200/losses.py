"""GAN and OCR losses (mirror of models/losses/gan_losses.py:8-16, ocr_losses.py:8-20).  All
losses are sums divided by the GLOBAL batch size, so per-replica gradients are summed, not
averaged, across GPUs."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def generator_loss(y_pred: torch.Tensor, batch_size: int) -> torch.Tensor:
    return F.softplus(-y_pred).sum() / batch_size


def discriminator_loss(y_pred: torch.Tensor, y_true: torch.Tensor, batch_size: int) -> torch.Tensor:
    return (F.softplus(y_pred) + F.softplus(-y_true)).sum() / batch_size


def softmax_cross_entropy_loss(y_pred: torch.Tensor, y_true: torch.Tensor, batch_size: int) -> torch.Tensor:
    loss = F.cross_entropy(y_pred.reshape(-1, y_pred.shape[-1]).float(), y_true.reshape(-1).long(), reduction="none")
    return loss.sum() / batch_size


def mean_squared_loss(y_with_noise: torch.Tensor, y_without_noise: torch.Tensor, batch_size: int) -> torch.Tensor:
    loss = ((y_with_noise - y_without_noise) ** 2).mean(dim=-1)
    return loss.sum() / batch_size
