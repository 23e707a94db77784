"""One TextBoxGAN training iteration restated on the CPU (TEST INFRASTRUCTURE — oracle/__init__.py).

Follows training_step.py:138-402 (``TrainingStep._train_step`` and helpers), train.py:110-129
(optimiser hyper-parameters), generator.py:48-59 (EMA) and tf.keras 2.8 ``optimizer_v2.Adam``
semantics.  All randomness is injected through ``draws``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import aster as A
from . import stylegan as S

Params = Dict[str, torch.Tensor]


def mask_text_box(fake_images: torch.Tensor, input_words: torch.Tensor, char_width: int) -> torch.Tensor:
    """utils/utils.py:11-45"""
    keep = torch.where(input_words == 0, 0.0, 1.0).to(fake_images.dtype)
    if isinstance(char_width, int):
        mask = torch.repeat_interleave(keep, char_width, dim=1)[:, None, None, :]
    else:
        # EXTENSION (outside the reference's domain: tf.repeat needs integer repeats): for a fractional
        # char_width = W/mcn, column x belongs to character floor(x / char_width); equals the line above for ints
        from fractions import Fraction

        cw = Fraction(char_width)
        idx = (torch.arange(fake_images.shape[3]) * cw.denominator) // cw.numerator
        mask = keep[:, idx][:, None, None, :]
    return fake_images * mask


def generator_loss(y_pred, batch_size):
    """gan_losses.py:8-10"""
    return F.softplus(-y_pred).sum() / batch_size


def discriminator_loss(y_pred, y_true, batch_size):
    """gan_losses.py:13-16"""
    return (F.softplus(y_pred) + F.softplus(-y_true)).sum() / batch_size


def update_optimizer_params(params: dict) -> dict:
    """train.py:110-129"""
    p = dict(params)
    mb_ratio = p["reg_interval"] / (p["reg_interval"] + 1)
    p["learning_rate"] = p["learning_rate"] * mb_ratio
    p["beta1"] = p["beta1"] ** mb_ratio
    p["beta2"] = p["beta2"] ** mb_ratio
    return p


@dataclass
class AdamState:
    """tf.keras.optimizers.Adam (optimizer_v2, non-amsgrad) state for a set of named variables."""
    lr: float
    beta1: float
    beta2: float
    eps: float
    iterations: int = 0
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)

    def apply(self, P: Params, grads: Dict[str, torch.Tensor]) -> None:
        """``t = iterations + 1; m,v EMAs; theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps)``
        (epsilon outside the bias-corrected sqrt — unlike torch.optim.Adam)."""
        t = self.iterations + 1
        lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)
        for name, g in grads.items():
            if name not in self.m:
                self.m[name] = torch.zeros_like(P[name])
                self.v[name] = torch.zeros_like(P[name])
            self.m[name] = self.beta1 * self.m[name] + (1.0 - self.beta1) * g
            self.v[name] = self.beta2 * self.v[name] + (1.0 - self.beta2) * g * g
            P[name] = P[name] - lr_t * self.m[name] / (torch.sqrt(self.v[name]) + self.eps)
        self.iterations = t


def make_adam(opt_cfg: dict) -> AdamState:
    p = update_optimizer_params(opt_cfg)
    return AdamState(lr=p["learning_rate"], beta1=p["beta1"], beta2=p["beta2"], eps=p["epsilon"])


def set_as_moving_average_of(clone: Params, src: Params) -> None:
    """generator.py:48-59 — beta 0.99 for every weight (trainable or not), w_avg copied."""
    for name in clone:
        beta = 0.0 if "w_avg" in name else 0.99
        clone[name] = src[name] + (clone[name] - src[name]) * beta


G_SCOPES = ("synthesis/", "latent_encoder/")      # training_step.py:196
OCR_SCOPES = ("synthesis/", "word_encoder/")      # training_step.py:203


@dataclass
class StepState:
    G: Params
    D: Params
    aster: Params
    g_opt: AdamState
    ocr_opt: AdamState
    d_opt: AdamState
    pl_mean: torch.Tensor
    g_reg_interval: int = 8
    d_reg_interval: int = 16


def path_length_reg(st: StepState, cfg, input_words, draws: dict, Gp: Params, fused: bool):
    """training_step.py:300-347.  draws: pl_z [B/2,S], pl_noises (list, B/2), pl_image_noise [B/2,3,H,W]."""
    shrink = 2 if cfg.batch_size_per_gpu // 2 >= 1 else cfg.batch_size_per_gpu      # :41-46
    pl_minibatch = max(1, cfg.batch_size_per_gpu // shrink)                        # :314-316
    pl_z = draws["pl_z"]
    # generator(...) with the default training=False: no dropout, no mixing, psi = 1 (:325-329)
    img, style = S.generator(input_words[:pl_minibatch], pl_z, Gp, cfg, training=False,
                             draws={"noises": draws["pl_noises"]}, ret_style=True, fused=fused)
    pl_noise = draws["pl_image_noise"] * (1.0 / math.sqrt(float(cfg.image_width) * float(cfg.char_height)))  # :53-55,330
    pl_noise_applied = (img * pl_noise).sum()
    (pl_grads,) = torch.autograd.grad(pl_noise_applied, style, create_graph=True)  # :333
    pl_lengths = torch.sqrt((pl_grads ** 2).sum(dim=2).mean(dim=1))                # :334-336
    pl_mean_val = st.pl_mean + 0.01 * (pl_lengths.mean() - st.pl_mean)             # :338-340
    st.pl_mean = pl_mean_val.detach()                                              # :341
    pl_penalty = (pl_lengths - st.pl_mean) ** 2                                    # :344
    pl_penalty = pl_penalty * shrink * st.g_reg_interval                           # :346
    return pl_penalty.sum() / cfg.batch_size                                       # :347


def r1_reg(st: StepState, cfg, real_images, Dp: Params):
    """training_step.py:349-373"""
    real_images = real_images.detach().requires_grad_(True)
    real_scores = S.discriminator(real_images, Dp, cfg)
    real_loss = real_scores.sum()
    (real_grads,) = torch.autograd.grad(real_loss, real_images, create_graph=True)
    r1 = (real_grads ** 2).sum(dim=(1, 2, 3))[:, None]
    r1 = r1 * (0.5 * 10.0) * st.d_reg_interval
    return real_scores, r1.sum() / cfg.batch_size


def ocr_loss_fn(st: StepState, cfg, fake_images, ocr_labels, ocr_images):
    """training_step.py:375-402"""
    x = A.convert_inputs(fake_images, ocr_labels, 1, cfg)
    logits = A.aster_inferer_call(x, st.aster, cfg)
    if cfg.ocr_loss_type == "mse":
        real_logits = A.aster_inferer_call(ocr_images, st.aster, cfg)
        return A.mean_squared_loss(real_logits, logits, cfg.batch_size)
    return A.softmax_cross_entropy_loss(logits, ocr_labels, cfg.batch_size)


def train_step(st: StepState, cfg, real_images, ocr_images, input_words, ocr_labels, do_r1_reg: bool,
               do_pl_reg: bool, ocr_loss_weight: float, draws: dict, *, fused: bool = True,
               with_ocr: bool = True, apply_updates: bool = True, ret_grads: bool = False):
    """training_step.py:138-222 for ONE replica holding the whole (global) batch.

    draws: z, dropout_mask, z2, mix_coin, mix_cutoff, noises (+ pl_* when do_pl_reg).
    Returns ((reg_g, g, pl), (reg_d, d, r1), ocr_loss) [+ the three gradient dicts]."""
    Gp = {k: (v.detach().clone().requires_grad_(True) if k not in S.NON_TRAINABLE else v.detach().clone())
          for k, v in st.G.items()}
    Dp = {k: v.detach().clone().requires_grad_(True) for k, v in st.D.items()}
    state_out: dict = {}
    fake = S.generator(input_words, draws["z"], Gp, cfg, training=True, draws=draws, fused=fused,
                       state_out=state_out)                                              # :178
    fake = mask_text_box(fake, input_words, cfg.char_width)                               # :180
    # _get_generator_losses :268-298
    fake_scores = S.discriminator(fake, Dp, cfg)
    g_loss = generator_loss(fake_scores, cfg.batch_size)
    zero = torch.zeros((), dtype=fake.dtype)
    pl_penalty = path_length_reg(st, cfg, input_words, draws, Gp, fused) if do_pl_reg else zero
    reg_g_loss = g_loss + pl_penalty
    # _get_discriminator_losses :237-266
    if do_r1_reg:
        real_scores, r1_penalty = r1_reg(st, cfg, real_images, Dp)
    else:
        real_scores = S.discriminator(real_images, Dp, cfg)
        r1_penalty = zero
    d_loss = discriminator_loss(fake_scores, real_scores, cfg.batch_size)
    reg_d_loss = d_loss + r1_penalty
    if with_ocr:
        ocr_loss = ocr_loss_fn(st, cfg, fake, ocr_labels, ocr_images) * ocr_loss_weight   # :191-192
    else:
        ocr_loss = zero

    g_names = S.trainable_names(Gp, G_SCOPES)
    o_names = S.trainable_names(Gp, OCR_SCOPES)
    d_names = list(Dp.keys())
    g_grads = dict(zip(g_names, torch.autograd.grad(reg_g_loss, [Gp[n] for n in g_names], retain_graph=True,
                                                    allow_unused=True)))
    if with_ocr:
        o_grads = dict(zip(o_names, torch.autograd.grad(ocr_loss, [Gp[n] for n in o_names], retain_graph=True,
                                                        allow_unused=True)))
    else:
        o_grads = {}
    d_grads = dict(zip(d_names, torch.autograd.grad(reg_d_loss, [Dp[n] for n in d_names], allow_unused=True)))
    g_grads = {k: v for k, v in g_grads.items() if v is not None}
    o_grads = {k: v for k, v in o_grads.items() if v is not None}
    d_grads = {k: v for k, v in d_grads.items() if v is not None}

    if apply_updates:
        st.g_opt.apply(st.G, g_grads)           # :194-199
        if with_ocr:
            st.ocr_opt.apply(st.G, o_grads)     # :201-206 (synthesis receives a second update)
        st.d_opt.apply(st.D, d_grads)           # :208-213
        if "w_avg" in state_out:
            st.G["latent_encoder/w_avg"] = state_out["w_avg"]

    out = ((reg_g_loss.detach(), g_loss.detach(), pl_penalty.detach()),
           (reg_d_loss.detach(), d_loss.detach(), r1_penalty.detach()),
           (ocr_loss / ocr_loss_weight).detach() if with_ocr else zero)                    # :215-222
    if ret_grads:
        return out, (g_grads, o_grads, d_grads), fake.detach()
    return out


# ----------------------------------------------------------------------------------------------
# synthetic inputs and draws
# ----------------------------------------------------------------------------------------------
def synthetic_batch(cfg, batch: int, gen: torch.Generator):
    """Seeded synthetic word batch with the loader's tensor contract
    (dataset_utils/training_data_loader.py:56-97): words, ASTER labels, real images zero right of
    ``len * char_width``."""
    import numpy as np

    from .tokens import main_to_aster_ids

    mcn = cfg.max_char_number
    lens = torch.randint(1, mcn + 1, (batch,), generator=gen)
    chars = torch.randint(1, 70, (batch, mcn), generator=gen)
    pos = torch.arange(mcn)[None, :]
    words = torch.where(pos < lens[:, None], chars, torch.zeros_like(chars)).to(torch.int32)
    labels = torch.from_numpy(main_to_aster_ids(words.numpy().astype(np.int64))).to(torch.int32)
    real = torch.rand(batch, 3, cfg.char_height, cfg.image_width, generator=gen) * 2 - 1
    real = mask_text_box(real, words, cfg.char_width)
    return real, words, labels


def make_draws(cfg, batch: int, gen: torch.Generator, *, with_pl: bool = False, dtype=torch.float32) -> dict:
    res = cfg.generator_resolutions
    n_style = 3 * (len(res) - 1)

    def randn(*shape):
        return torch.randn(*shape, generator=gen, dtype=torch.float64).to(dtype)

    d = {
        "z": randn(batch, cfg.z_dim),
        "dropout_mask": (torch.rand(batch, cfg.max_char_number, cfg.embedding_out_dim, generator=gen) < 0.7).to(dtype),
        "z2": randn(batch, cfg.z_dim),
        "mix_coin": float(torch.rand((), generator=gen)),
        "mix_cutoff": int(torch.randint(1, n_style, (), generator=gen)),
        "noises": [randn(batch, 1, h, w) for (h, w) in res[1:] for _ in range(2)],
    }
    if with_pl:
        pb = max(1, batch // 2)
        d["pl_z"] = randn(pb, cfg.z_dim).requires_grad_(False)
        d["pl_noises"] = [randn(pb, 1, h, w) for (h, w) in res[1:] for _ in range(2)]
        d["pl_image_noise"] = randn(pb, 3, cfg.char_height, cfg.image_width)
    return d
