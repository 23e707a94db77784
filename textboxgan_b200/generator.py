"""Generator = WordEncoder + LatentEncoder + Synthesis (mirror of
models/custom_stylegan2/generator.py:10-59, latent_encoder.py:8-99, layers/synthesis_block.py,
models/word_encoder.py:7-63) on the B200 kernels.

Call surface kept from the reference: ``generator((words, z), batch_size=..., ret_style=False,
truncation_psi=1.0, training=False)``; attributes ``word_encoder``, ``latent_encoder``,
``synthesis``, ``n_style``; ``set_as_moving_average_of``; ``get_weights``/``set_weights``.
One extension: ``draws`` injects every random tensor so that parity tests
can replay the oracle's randomness; when absent, draws come from the device RNG.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import kernels as K
from . import layers as L
from .config import Config
from .model_base import Model, Submodel, fresh_seed

MAIN_VOCAB_SIZE = 70  # len(cfg.char_tokenizer.main.word_index) (word_encoder.py:17): <OOV> + 69 chars


class Generator(Model):
    def __init__(self, cfg: Config, device="cuda", seed: Optional[int] = None):
        super().__init__("generator")
        self.cfg = cfg
        self.device = torch.device(device)
        self.n_style = cfg.n_style                                           # generator.py:16
        self.w_ema_decay = 0.995                                             # latent_encoder.py:18
        self.style_mixing_prob = 0.9                                         # latent_encoder.py:19
        self.dropout_rate = 0.3                                              # word_encoder.py:10
        self._build(seed)
        self.word_encoder = Submodel(self, "word_encoder/")
        self.latent_encoder = Submodel(self, "latent_encoder/")
        self.synthesis = Submodel(self, "synthesis/")

    # ------------------------------------------------------------------------------------------
    def _build(self, seed: Optional[int]) -> None:
        cfg = self.cfg
        g = torch.Generator().manual_seed(seed if seed is not None else fresh_seed())
        S = cfg.style_dim

        def randn(*shape, std=1.0):
            return torch.randn(*shape, generator=g) * std

        # Variable order = flat-buffer order: [w0_embedding | word_encoder | synthesis |
        # latent_encoder | w_avg] makes both optimiser groups of training_step.py:196,203
        # (synthesis + latent_encoder, synthesis + word_encoder) contiguous ranges.
        # word_encoder.py:28-37 (+ Keras Dense(256): glorot-uniform kernel, zero bias)
        self.add_weight("word_encoder/w0_embedding", torch.zeros(1, cfg.embedding_out_dim), trainable=False)
        self.add_weight("word_encoder/w_embedding", randn(MAIN_VOCAB_SIZE - 1, cfg.embedding_out_dim))
        lim = math.sqrt(6.0 / (cfg.embedding_out_dim + cfg.word_encoder_dense_dim))
        self.add_weight("word_encoder/fc/kernel",
                        (torch.rand(cfg.embedding_out_dim, cfg.word_encoder_dense_dim, generator=g) * 2 - 1) * lim)
        self.add_weight("word_encoder/fc/bias", torch.zeros(cfg.word_encoder_dense_dim))

        def modconv(prefix, k, I, O):
            self.add_weight(prefix + "/w", randn(k, k, I, O))                # modulated_conv2d.py:62-63
            self.add_weight(prefix + "/mod_dense/w", randn(S, I))
            self.add_weight(prefix + "/mod_bias/b", torch.zeros(I))

        def torgb(prefix, Cc):
            modconv(prefix + "/conv", 1, Cc, 3)
            self.add_weight(prefix + "/bias/b", torch.zeros(3))

        res, fm = cfg.generator_resolutions, cfg.generator_feat_maps
        torgb(f"synthesis/{res[0][0]}x{res[0][1]}/ToRGB", fm[0])
        prev = fm[0]
        for (h, w), f in zip(res[1:], fm[1:]):
            pb = f"synthesis/{h}x{w}/block"
            modconv(pb + "/conv_0", 3, prev, f)
            self.add_weight(pb + "/noise_0/w", torch.zeros(()))
            self.add_weight(pb + "/bias_0/b", torch.zeros(f))
            modconv(pb + "/conv_1", 3, f, f)
            self.add_weight(pb + "/noise_1/w", torch.zeros(()))
            self.add_weight(pb + "/bias_1/b", torch.zeros(f))
            torgb(f"synthesis/{h}x{w}/ToRGB", f)
            prev = f
        # mapping_block.py:13,24-33 (lrmul 0.01 -> init std 100)
        for i in range(cfg.n_mapping):
            in_dim = cfg.z_dim if i == 0 else S
            self.add_weight(f"latent_encoder/g_mapping/dense_{i}/w", randn(in_dim, S, std=100.0))
            self.add_weight(f"latent_encoder/g_mapping/bias_{i}/b", torch.zeros(S))
        self.add_weight("latent_encoder/w_avg", torch.zeros(S), trainable=False)  # latent_encoder.py:29-37
        self.to(self.device)

    # ------------------------------------------------------------------------------------------
    def _word_encoder(self, words: torch.Tensor, batch_size: int, dropout_mask: Optional[torch.Tensor]):
        """word_encoder.py:39-63 -> NHWC [B, 2, 8, fm0] (the reference returns NCHW [B,fm0,2,8])."""
        cfg, P = self.cfg, self.params
        out_h, out_w = cfg.generator_resolutions[0]
        out_c = cfg.generator_feat_maps[0]
        if L.use_fused():
            from .fused import WordEncoderFn

            return WordEncoderFn.apply(words, P["word_encoder/w0_embedding"], P["word_encoder/w_embedding"], dropout_mask,
                                       P["word_encoder/fc/kernel"], P["word_encoder/fc/bias"], 1.0 - self.dropout_rate,
                                       (out_h, out_w, out_c))
        table = torch.cat([P["word_encoder/w0_embedding"], P["word_encoder/w_embedding"]], dim=0)
        emb = table[words.long()]
        if dropout_mask is not None:
            emb = emb * dropout_mask / (1.0 - self.dropout_rate)
        x = emb.reshape(batch_size * cfg.max_char_number, cfg.embedding_out_dim)
        x = torch.relu(x @ P["word_encoder/fc/kernel"] + P["word_encoder/fc/bias"])
        # reference: reshape [B, out_w, out_c, out_h] then transpose (0,2,3,1) -> [B, c, h, w];
        # NHWC is therefore [B, h, w, c] = permute(0, 3, 1, 2) of the reshaped tensor.
        return x.reshape(batch_size, out_w, out_c, out_h).permute(0, 3, 1, 2).contiguous().to(L.ACT_DTYPE)

    def _mapping(self, z: torch.Tensor) -> torch.Tensor:
        """mapping_block.py:35-45.  lrelu(v)*sqrt2 == lrelu(sqrt2*v) (positive homogeneity), so each
        layer is one addmm (equalised-LR coefficient and sqrt2 folded into alpha) + one leaky-relu."""
        P = self.params
        if L.use_fused():
            from .fused import DenseAct, PixelNorm

            x = PixelNorm.apply(z)
            for i in range(self.cfg.n_mapping):
                w = P[f"latent_encoder/g_mapping/dense_{i}/w"]
                x = DenseAct.apply(x, w, P[f"latent_encoder/g_mapping/bias_{i}/b"], L.runtime_coef(w.shape, 1.0, 0.01), 0.01,
                                   1, L.SQRT2, "mapping")
            return x
        x = z * torch.rsqrt(torch.mean(z * z, dim=1, keepdim=True) + 1e-8)
        for i in range(self.cfg.n_mapping):
            w = P[f"latent_encoder/g_mapping/dense_{i}/w"]
            b = P[f"latent_encoder/g_mapping/bias_{i}/b"]
            pre = torch.addmm(b * (0.01 * L.SQRT2), x, w, alpha=L.runtime_coef(w.shape, 1.0, 0.01) * L.SQRT2)
            x = torch.nn.functional.leaky_relu(pre, 0.2)
        return x

    def _latent_encoder(self, z, training: bool, truncation_psi: float, draws: dict):
        """latent_encoder.py:80-99"""
        P = self.params
        n = self.n_style
        if training:
            # both latents of the style-mixing pair go through the mapping network as one batch
            z2 = draws["z2"].to(z.device) if "z2" in draws else torch.randn_like(z)   # :49
            w12 = self._mapping(torch.cat([z, z2], dim=0))
            wb = w12[: z.shape[0], None, :].expand(-1, n, -1)
        else:
            wb = self._mapping(z)[:, None, :].expand(-1, n, -1)
        if training:
            with torch.no_grad():                                            # :39-45
                batch_avg = wb[:, 0].mean(dim=0)
                w_avg = P["latent_encoder/w_avg"]
                w_avg.copy_(batch_avg + (w_avg - batch_avg) * self.w_ema_decay)
            wb2 = w12[z.shape[0]:, None, :].expand(-1, n, -1)
            # :55-60 — one scalar coin / cutoff per batch, drawn on the device (no host sync, so the
            # whole step can be replayed from a CUDA graph)
            coin = torch.as_tensor(draws["mix_coin"], device=z.device, dtype=torch.float32) if "mix_coin" in draws \
                else torch.rand((), device=z.device)
            cut = torch.as_tensor(draws["mix_cutoff"], device=z.device) if "mix_cutoff" in draws \
                else torch.randint(1, n, (), device=z.device)
            cutoff = torch.where(coin < self.style_mixing_prob, cut, torch.full_like(cut, n))
            idx = torch.arange(n, device=z.device)[None, :, None]
            wb = torch.where(idx < cutoff, wb, wb2)
        if not training:
            w_avg = P["latent_encoder/w_avg"]
            wb = w_avg + (wb - w_avg) * truncation_psi                       # :73-78
        return wb

    def _synthesis(self, x, style, noises, fused_epilogue: bool = False, mask_words=None):
        """synthesis_block.py:137-156; x NHWC bf16, returns NCHW fp32 image (zeroed right of the word when ``mask_words``
        is given: mask_text_box, utils/utils.py:11-45, fused into the last skip sum)."""
        cfg, P = self.cfg, self.params
        res = cfg.generator_resolutions
        # style rows: ToRGB_0 <- 0; block i: conv_0 <- 3i, conv_1 <- 3i+1, ToRGB <- 3i+2 (:140,143-147)
        prefixes = [f"synthesis/{res[0][0]}x{res[0][1]}/ToRGB/conv"]
        idxs = [0]
        for i, (h, w) in enumerate(res[1:]):
            prefixes += [f"synthesis/{h}x{w}/block/conv_0", f"synthesis/{h}x{w}/block/conv_1",
                         f"synthesis/{h}x{w}/ToRGB/conv"]
            idxs += [3 * i, 3 * i + 1, 3 * i + 2]
        if L.use_fused():
            sc = L.all_style_scales(style.float(), P, prefixes, idxs)        # one launch for all layers
        else:
            sc = [None] * len(prefixes)
        fused_rgb = L.use_fused()
        last = len(res) - 2
        y = L.to_rgb(x, style[:, 0], P, f"synthesis/{res[0][0]}x{res[0][1]}/ToRGB", s_pre=sc[0],
                     skip_args=(None, None, False) if fused_rgb else None)
        for i, (h, w) in enumerate(res[1:]):
            pb = f"synthesis/{h}x{w}/block"
            s0, s1, s2 = style[:, 3 * i], style[:, 3 * i + 1], style[:, 3 * i + 2]
            n0, n1 = noises[2 * i], noises[2 * i + 1]
            x = L.modulated_conv2d(x, s0, P, pb + "/conv_0", up=True, noise=n0, noise_strength=P[pb + "/noise_0/w"],
                                   bias=P[pb + "/bias_0/b"], act=True, fused_epilogue=fused_epilogue,
                                   s_pre=sc[3 * i + 1])
            if fused_rgb and L.FUSE_RGB_BACKWARD and not fused_epilogue:
                x, y = L.modulated_conv2d_rgb(x, P, pb + "/conv_1", f"synthesis/{h}x{w}/ToRGB", noise=n1,
                                              noise_strength=P[pb + "/noise_1/w"], bias=P[pb + "/bias_1/b"],
                                              s_conv=sc[3 * i + 2], s_rgb=sc[3 * i + 3], y_prev=y,
                                              mask_words=mask_words if i == last else None, nchw=i == last)
                continue
            x = L.modulated_conv2d(x, s1, P, pb + "/conv_1", up=False, noise=n1, noise_strength=P[pb + "/noise_1/w"],
                                   bias=P[pb + "/bias_1/b"], act=True, fused_epilogue=fused_epilogue,
                                   s_pre=sc[3 * i + 2])
            if fused_rgb:
                # ToRGB + upsampled skip (+ mask + NCHW on the last block) in one launch
                y = L.to_rgb(x, s2, P, f"synthesis/{h}x{w}/ToRGB", s_pre=sc[3 * i + 3],
                             skip_args=(y, mask_words if i == last else None, i == last))
            else:
                y = L.upsample_rgb(y) + L.to_rgb(x, s2, P, f"synthesis/{h}x{w}/ToRGB", s_pre=sc[3 * i + 3])
        if fused_rgb:
            return y
        y = y.permute(0, 3, 1, 2).contiguous()
        if mask_words is not None:
            from .utils import mask_text_box

            y = mask_text_box(y, mask_words, self.cfg.char_width)
        return y

    def _noises(self, batch: int, draws: dict):
        if "noises" in draws:
            # oracle layout [B,1,H,W] or device layout [B,H,W]
            return [n.reshape(n.shape[0], n.shape[-2], n.shape[-1]).to(self.device, torch.float32) for n in draws["noises"]]
        return [torch.randn(batch, h, w, device=self.device) for (h, w) in self.cfg.generator_resolutions[1:]
                for _ in range(2)]

    def __call__(self, inputs, batch_size: Optional[int] = None, ret_style: bool = False,
                 truncation_psi: float = 1.0, training: bool = False, draws: Optional[dict] = None,
                 mask_output: bool = False):
        """generator.py:19-43.  ``mask_output`` (extension): also apply mask_text_box(image, input_words) — the training
        step's next statement (training_step.py:180) — inside the last ToRGB launch."""
        input_words, z_latent = inputs
        draws = draws or {}
        batch_size = batch_size or input_words.shape[0]
        mask = None
        if training:
            mask = draws["dropout_mask"].to(self.device) if "dropout_mask" in draws else (
                torch.rand(batch_size, self.cfg.max_char_number, self.cfg.embedding_out_dim, device=self.device)
                >= self.dropout_rate).float()
        x = self._word_encoder(input_words, batch_size, mask)
        style = self._latent_encoder(z_latent, training, truncation_psi, draws)
        image_out = self._synthesis(x, style, self._noises(batch_size, draws),
                                    fused_epilogue=not torch.is_grad_enabled(),
                                    mask_words=input_words if mask_output else None)
        return (image_out, style) if ret_style else image_out

    @torch.no_grad()
    def set_as_moving_average_of(self, src_net: "Generator") -> None:
        """generator.py:48-59 — ``cw <- lerp(sw, cw, 0.99)`` for every weight (one kernel over the
        flat buffer), ``w_avg`` copied (beta 0); the frozen zero ``w0_embedding`` is unchanged."""
        beta, beta_nontrainable = 0.99, 0.0
        assert self.flat.shape == src_net.flat.shape
        K.ema_step(self.flat, src_net.flat, beta)
        for name in self._non_trainable:
            cw, sw = self.params[name], src_net.params[name]
            b = beta_nontrainable if "w_avg" in name else beta
            cw.copy_(sw + (cw - sw) * b)
