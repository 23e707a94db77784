"""Shared builders for the parity tests (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import torch

from oracle import aster as OA
from oracle import stylegan as OS
from oracle import train_step as OT
from textboxgan_b200.config import baseline_config


def small_cfg(batch=4):
    cfg = baseline_config(0)
    cfg.batch_size_per_gpu = batch
    cfg.batch_size = batch
    cfg.aster_synthetic_weights = True      # explicit opt-in: the reference's pretrained ASTER is not in its repository
    return cfg


def fractional_cfg(batch=4):
    """Tiny ladder (64x16) with max_char_number=12: char_width = 16/3, the BASELINE configs[2]/[3] situation."""
    from fractions import Fraction

    from textboxgan_b200.config import Config

    return Config(char_height=16, char_width=Fraction(64, 12), max_char_number=12, z_dim=128, style_dim=128,
                  batch_size_per_gpu=batch, num_replicas=1)


def perturbed_params(cfg, seed=1, noise_strength=0.2):
    """Reference-initialised parameters with non-zero biases / noise strengths / w_avg so that
    every term of the forward pass is exercised (the reference initialises them to zero)."""
    g = torch.Generator().manual_seed(seed)
    GP = OS.init_generator_params(cfg, g)
    DP = OS.init_discriminator_params(cfg, g)
    for k in GP:
        if "noise_" in k:
            GP[k] = torch.tensor(noise_strength)
        elif k.endswith("/b") or k.endswith("/bias"):
            GP[k] = torch.randn(GP[k].shape, generator=g) * 0.1
    GP["latent_encoder/w_avg"] = torch.randn(cfg.style_dim, generator=g) * 0.1
    for k in DP:
        if k.endswith("/b"):
            DP[k] = torch.randn(DP[k].shape, generator=g) * 0.1
    return GP, DP, g


def to64(draws):
    out = {}
    for k, v in draws.items():
        if torch.is_tensor(v) and v.is_floating_point():
            out[k] = v.double()
        elif isinstance(v, list):
            out[k] = [t.double() for t in v]
        else:
            out[k] = v
    return out


def rel_err(a, b):
    import numpy as np

    a = torch.as_tensor(np.asarray(a)) if not torch.is_tensor(a) else a
    b = torch.as_tensor(np.asarray(b)) if not torch.is_tensor(b) else b
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
