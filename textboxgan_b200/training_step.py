"""TrainingStep — one G + D + OCR training iteration (mirror of training_step.py:14-402).

Same constructor arguments, same ``dist_train_step`` signature and return structure as the
reference, so ``train.py`` drives it unchanged apart from its imports.  Differences that are
consequences of the B200 design, not of behaviour:

* one process per GPU: ``cfg.strategy.run`` is a local call and the seven ``strategy.reduce``
  calls (training_step.py:106-134) are one packed all-reduce;
* the three ``tape.gradient`` + ``apply_gradients`` pairs (:194-213) are three
  ``torch.autograd.grad`` calls on the shared graph followed by three flat-buffer Adam updates
  (gradient all-reduce SUM + one ``tbg_adam_step`` each); all gradients are taken at the
  pre-update weights, exactly as with the reference's single persistent tape;
* ``draws`` optionally injects the random tensors for parity tests.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch

from .aster_inferer import AsterInferer
from .config import Config
from .discriminator import Discriminator
from .generator import Generator
from . import layers as L
from .losses import discriminator_loss, generator_loss, mean_squared_loss, softmax_cross_entropy_loss
from .optimizers import Adam
from .utils import mask_text_box


class TrainingStep:
    """Infer the model, computes the associated losses and backpropagates them."""

    def __init__(
        self,
        generator: Generator,
        discriminator: Discriminator,
        aster_ocr: Optional[AsterInferer],
        g_optimizer: Adam,
        ocr_optimizer: Adam,
        d_optimizer: Adam,
        g_reg_interval: int,
        d_reg_interval: int,
        pl_mean: torch.Tensor,
        cfg: Optional[Config] = None,
    ):
        if cfg is None:
            cfg = generator.cfg
        self.cfg = cfg
        self.generator = generator
        self.discriminator = discriminator
        self.aster_ocr = aster_ocr
        self.g_optimizer = g_optimizer
        self.ocr_optimizer = ocr_optimizer
        self.d_optimizer = d_optimizer
        self.g_reg_interval = g_reg_interval
        self.d_reg_interval = d_reg_interval
        self.batch_size = cfg.batch_size
        self.batch_size_per_gpu = cfg.batch_size_per_gpu
        self.pl_mean = pl_mean

        pl_minibatch_shrink = 2                                               # training_step.py:41-46
        self.pl_minibatch_shrink = pl_minibatch_shrink if self.batch_size_per_gpu // pl_minibatch_shrink >= 1 \
            else self.batch_size_per_gpu
        self.pl_weight = float(self.pl_minibatch_shrink)
        self.pl_decay = 0.01
        self.r1_gamma = 10.0
        self.ocr_loss_type = cfg.ocr_loss_type
        self.z_dim = cfg.z_dim
        self.char_width = cfg.char_width
        self.pl_noise_scaler = 1.0 / math.sqrt(float(cfg.image_width) * float(cfg.char_height))   # :53-55

        # variable groups of :196, :203, :210 — contiguous ranges of the flat buffers
        self._g_names = generator.trainable_names(("synthesis/", "latent_encoder/"))
        self._ocr_names = generator.trainable_names(("word_encoder/", "synthesis/"))
        self._d_names = discriminator.trainable_names()
        # CUDA-graph replay of the step (one graph per (do_r1_reg, do_pl_reg) variant — the
        # reference's tf.function likewise retraces per Python bool, train.py:182-183)
        self.use_cuda_graph = False
        self.batch_d_calls = True       # evaluate D(fake) and D(real) as one concatenated pass
        self.overlap_ocr = True         # OCR branch on a second CUDA stream (parallel sub-graph when captured)
        self.overlap_reg = True         # path-length / R1 regulariser branches on their own streams (lazy-reg iterations)
        self.overlap_comm = True        # start each group's gradient all-reduce as soon as its backward pass is done
        self._reg_streams = {}
        self._side = None
        self._capture = None            # stream of the eager warm-up of every step variant and of its capture
        self._graphs = {}
        self._static = None
        self._step_weights = None          # fused.StepWeights: grouped weight preparation plan

    # ------------------------------------------------------------------------------------------
    def dist_train_step(self, real_images, ocr_images, input_words, ocr_labels, do_r1_reg: bool, do_pl_reg: bool,
                        ocr_loss_weight: float, draws: Optional[dict] = None):
        """training_step.py:57-136.  Returns ``((reg_g, g, pl), (reg_d, d, r1), ocr)`` summed over
        replicas (every loss already carries 1/global_batch)."""
        if self.use_cuda_graph and not draws:
            return self._graphed_step(real_images, ocr_images, input_words, ocr_labels, do_r1_reg, do_pl_reg,
                                      ocr_loss_weight)
        strategy = self.cfg.strategy
        if strategy is None:
            gen_losses, disc_losses, ocr_loss = self._train_step(real_images, ocr_images, input_words, ocr_labels,
                                                                 do_r1_reg, do_pl_reg, ocr_loss_weight, draws)
            return gen_losses, disc_losses, ocr_loss
        gen_losses, disc_losses, ocr_loss = strategy.run(
            self._train_step,
            args=(real_images, ocr_images, input_words, ocr_labels, do_r1_reg, do_pl_reg, ocr_loss_weight, draws))
        red = strategy.reduce_many(list(gen_losses) + list(disc_losses) + [ocr_loss])
        mean_pl = red[2] if do_pl_reg else torch.zeros((), device=red[0].device)
        return (red[0], red[1], mean_pl), (red[3], red[4], red[5]), red[6]

    # ------------------------------------------------------------------------------------------
    def _eager_reduced_step(self, real_images, ocr_images, input_words, ocr_labels, do_r1_reg, do_pl_reg,
                            ocr_loss_weight):
        gen_losses, disc_losses, ocr_loss = self._train_step(real_images, ocr_images, input_words, ocr_labels,
                                                             do_r1_reg, do_pl_reg, ocr_loss_weight, None)
        vals = list(gen_losses) + list(disc_losses) + [ocr_loss]
        strategy = self.cfg.strategy
        if strategy is not None:
            vals = strategy.reduce_many(vals)
        return torch.stack([v.float().reshape(()) for v in vals])

    def _graphed_step(self, real_images, ocr_images, input_words, ocr_labels, do_r1_reg, do_pl_reg,
                      ocr_loss_weight):
        """Replay the whole iteration (forward, three backward passes, gradient all-reduces, three
        Adam updates) from a CUDA graph: inputs are copied into static buffers, the step-dependent
        scalars (Adam bias correction, OCR loss weight) live in device memory."""
        dev = self.generator.device
        opts = (self.g_optimizer, self.ocr_optimizer, self.d_optimizer)
        if self._static is None:
            self._static = {
                "real": torch.empty_like(real_images, device=dev),
                "words": torch.empty_like(input_words, device=dev),
                "labels": torch.empty_like(ocr_labels, device=dev),
                "ocr_images": ocr_images.to(dev).clone() if torch.is_tensor(ocr_images)
                else torch.zeros((), device=dev),
                "ocr_w": torch.zeros((), device=dev),
            }
            for o in opts:
                o.use_device_lr(dev)
        st = self._static
        st["real"].copy_(real_images, non_blocking=True)
        st["words"].copy_(input_words, non_blocking=True)
        st["labels"].copy_(ocr_labels, non_blocking=True)
        if torch.is_tensor(ocr_images) and ocr_images.dim() > 0:
            st["ocr_images"].copy_(ocr_images, non_blocking=True)
        st["ocr_w"].fill_(float(ocr_loss_weight))
        for o in opts:
            o.refresh_device_lr()
        key = (bool(do_r1_reg), bool(do_pl_reg))
        entry = self._graphs.get(key)
        if entry is None:
            # first use of this variant: run it eagerly once (lazy allocations, cuFuncSetAttribute,
            # optimiser slots), then capture.  The warm-up runs on the stream the capture will use: autograd pins every
            # node -- including the parameters' gradient accumulators, which outlive the step for as long as a dead
            # graph awaits garbage collection -- to the stream it was created on, and a node left on the legacy default
            # stream makes the captured backward wait for it (cudaErrorStreamCaptureImplicit).
            cs = self._capture_stream(dev)
            cur = torch.cuda.current_stream(dev)
            cs.wait_stream(cur)
            with torch.cuda.stream(cs):
                out = self._eager_reduced_step(st["real"], st["ocr_images"], st["words"], st["labels"], key[0], key[1],
                                               st["ocr_w"])
            cur.wait_stream(cs)
            out.record_stream(cur)
            self._graphs[key] = {"graph": None, "out": out, "warm": 1}
            res = out
        elif entry["graph"] is None:
            for o in opts:
                o.defer_iteration = True
            from . import lib as _lib

            import gc

            gc.collect()                       # dead autograd graphs of earlier eager steps (see the warm-up note above)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            n0 = _lib.load().tbg_launch_count()
            with torch.cuda.graph(graph, stream=self._capture_stream(dev)):
                out = self._eager_reduced_step(st["real"], st["ocr_images"], st["words"], st["labels"], key[0],
                                               key[1], st["ocr_w"])
            # launches of this repo's kernels recorded in the graph (replayed on every step)
            entry["tbg_launches"] = int(_lib.load().tbg_launch_count() - n0)
            for o in opts:
                o.defer_iteration = False
            entry["graph"], entry["out"] = graph, out
            graph.replay()
            self._bump_iterations()
            res = out
        else:
            entry["graph"].replay()
            self._bump_iterations()
            res = entry["out"]
        r = res.unbind(0)
        return (r[0], r[1], r[2]), (r[3], r[4], r[5]), r[6]

    def _capture_stream(self, dev):
        if self._capture is None:
            self._capture = torch.cuda.Stream(device=dev)
        return self._capture

    def _side_stream(self, dev):
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    def _reg_stream(self, dev, name: str):
        if name not in self._reg_streams:
            self._reg_streams[name] = torch.cuda.Stream(device=dev)
        return self._reg_streams[name]

    def graph_launches(self, do_r1_reg: bool = False, do_pl_reg: bool = False) -> int:
        """Number of this repo's kernel launches inside the captured graph of a step variant."""
        e = self._graphs.get((bool(do_r1_reg), bool(do_pl_reg)))
        return int(e.get("tbg_launches", 0)) if e else 0

    def _bump_iterations(self):
        self.g_optimizer.iterations.value += 1
        self.d_optimizer.iterations.value += 1
        if self.aster_ocr is not None:
            self.ocr_optimizer.iterations.value += 1

    # ------------------------------------------------------------------------------------------
    def _train_step(self, real_images, ocr_images, input_words, ocr_labels, do_r1_reg: bool, do_pl_reg: bool,
                    ocr_loss_weight: float, draws: Optional[dict] = None):
        """training_step.py:138-222"""
        from . import fused as _fused

        draws = draws or {}
        if self._step_weights is None:
            self._step_weights = _fused.StepWeights()
        self._step_weights.begin_step()                # all weight preparations of the iteration: one grouped launch
        G, D = self.generator, self.discriminator
        dev = G.device
        z = draws["z"].to(dev) if "z" in draws else torch.randn(self.batch_size_per_gpu, self.z_dim, device=dev)
        # :178 + :180 — mask_text_box(fake_images, input_words) runs inside the generator's last ToRGB launch
        fake_images = G((input_words, z), training=True, draws=draws, mask_output=True)

        # The OCR branch (convert_inputs -> ASTER -> loss, :375-402) only shares ``fake_images`` with the
        # discriminator branch: it is issued on a second stream so that its latency-bound kernels (whole-sequence
        # LSTM, attention decoder, small ResNet convolutions) overlap the discriminator's; inside the captured
        # CUDA graph the two branches become parallel sub-graphs.
        main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
        side = self._side_stream(dev) if (main is not None and self.overlap_ocr and self.aster_ocr is not None) else None
        ocr_loss = None
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ocr_loss = self._get_ocr_loss(fake_images, ocr_labels, ocr_images)
                ocr_loss = ocr_loss_weight * ocr_loss                                        # :191-192

        # The two regulariser branches (path length :300-347 on its own generator pass, R1 :349-373 on D(real)) share only
        # weights with the adversarial branch.  Their double-backward graphs are thousands of small launches at a reduced
        # batch, so each is issued on its own stream: forward here, and — because autograd runs a node's backward on the
        # stream of its forward — the second-order backward passes as well, next to the large kernels of the main branch.
        reg_pre = {}
        if main is not None and self.overlap_reg and (do_pl_reg or do_r1_reg):
            # parameters are leaves created on the main stream and differentiated from several streams on purpose (all
            # gradients are taken with torch.autograd.grad, which joins the streams before it returns)
            if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            if do_pl_reg:
                st_pl = self._reg_stream(dev, "pl")
                st_pl.wait_stream(main)
                with torch.cuda.stream(st_pl):
                    reg_pre["pl"] = self._path_length_reg(input_words, draws)
            if do_r1_reg:
                st_r1 = self._reg_stream(dev, "r1")
                st_r1.wait_stream(main)
                with torch.cuda.stream(st_r1):
                    reg_pre["r1"] = self._r1_reg(real_images)

        # D(fake) and D(real) (training_step.py:260,288) as ONE concatenated pass when no R1 penalty is due:
        # every discriminator layer is per-sample except the minibatch statistic, taken per call.
        self._batched_d = bool(self.batch_d_calls and L.use_fused() and not do_r1_reg)
        real_scores = None
        if self._batched_d:
            nb = fake_images.shape[0]
            scores = D(torch.cat([fake_images, real_images.detach().to(fake_images.dtype)], dim=0), n_calls=2)
            fake_scores, real_scores = scores[:nb], scores[nb:]
        else:
            fake_scores = None
        if not self._batched_d and fake_scores is None and reg_pre:
            fake_scores = self.discriminator(fake_images)          # issued before the joins below
        for name, val in reg_pre.items():
            main.wait_stream(self._reg_streams[name])
            for t in (val if isinstance(val, tuple) else (val,)):
                t.record_stream(main)
        fake_scores, reg_g_loss, g_loss, pl_penalty = self._get_generator_losses(fake_images, do_pl_reg,
                                                                                 input_words, draws, fake_scores,
                                                                                 reg_pre.get("pl"))
        reg_d_loss, d_loss, r1_penalty = self._get_discriminator_losses(fake_scores, real_images, do_r1_reg,
                                                                        real_scores, reg_pre.get("r1"))
        if self.aster_ocr is not None and side is None:
            ocr_loss = self._get_ocr_loss(fake_images, ocr_labels, ocr_images)
            ocr_loss = ocr_loss_weight * ocr_loss                                            # :191-192

        g_vars = [G.params[n] for n in self._g_names]
        o_vars = [G.params[n] for n in self._ocr_names]
        d_vars = [D.params[n] for n in self._d_names]
        # three tape.gradient calls on one persistent tape (:194-213): all at pre-update weights
        # The discriminator group first: its gradient all-reduce (the largest of the three, ~62 MB) is started right away
        # and overlaps the two generator backward passes below (all three are taken at the pre-update weights, so the
        # order of the tape.gradient calls is free).
        updates = not draws.get("skip_updates")
        d_grads = torch.autograd.grad(reg_d_loss, d_vars, retain_graph=True, allow_unused=True)
        if updates and self.overlap_comm:
            self.d_optimizer.begin_apply(D, self._d_names, list(d_grads))
        g_fake_ocr = None
        if side is not None:
            # first half of the OCR pass (chain rule split at fake_images): back through the frozen recogniser on
            # the second stream, concurrently with the generator pass below
            with torch.cuda.stream(side):
                (g_fake_ocr,) = torch.autograd.grad(ocr_loss, [fake_images], retain_graph=True)
        with _fused.skip_weight_grads("dconv"), \
                _fused.backward_batch_limit("dconv", fake_images.shape[0] if self._batched_d else 1 << 30):
            # only generator variables are wanted from this pass
            g_grads = torch.autograd.grad(reg_g_loss, g_vars, retain_graph=True, allow_unused=True)
        if updates and self.overlap_comm:
            self.g_optimizer.begin_apply(G, self._g_names, list(g_grads))
        if g_fake_ocr is not None:
            main.wait_stream(side)
            g_fake_ocr.record_stream(main)
            o_grads = torch.autograd.grad(fake_images, o_vars, grad_outputs=g_fake_ocr, allow_unused=True)
        else:
            o_grads = torch.autograd.grad(ocr_loss, o_vars, allow_unused=True) if ocr_loss is not None else None
        self.last_grads = (g_grads, o_grads, d_grads) if draws.get("keep_grads") else None

        if updates:
            # reference order of the three updates (:194-213): generator group, OCR group (synthesis is updated twice), D
            if not self.overlap_comm:
                # all three cross-replica sums after the last backward pass, back to back on the communication stream:
                # they then overlap only the (many-CTA, bandwidth-bound) Adam kernels, never the persistent one-CTA-per-SM
                # tensor-core kernels, which wait for a whole wave when NCCL holds a few SMs
                self.g_optimizer.begin_apply(G, self._g_names, list(g_grads))
                if o_grads is not None:
                    self.ocr_optimizer.begin_apply(G, self._ocr_names, list(o_grads))
                self.d_optimizer.begin_apply(D, self._d_names, list(d_grads))
                self.g_optimizer.finish_apply()
                if o_grads is not None:
                    self.ocr_optimizer.finish_apply()
                self.d_optimizer.finish_apply()
            else:
                self.g_optimizer.finish_apply()
                if o_grads is not None:
                    self.ocr_optimizer.begin_apply(G, self._ocr_names, list(o_grads))
                    self.ocr_optimizer.finish_apply()
                self.d_optimizer.finish_apply()

        self._step_weights.end_step()
        gen_losses = (reg_g_loss.detach(), g_loss.detach(), pl_penalty.detach())
        disc_losses = (reg_d_loss.detach(), d_loss.detach(), r1_penalty.detach())
        ocr_out = (ocr_loss / ocr_loss_weight).detach() if ocr_loss is not None else torch.zeros((), device=dev)
        return gen_losses, disc_losses, ocr_out

    # ------------------------------------------------------------------------------------------
    def _get_discriminator_losses(self, fake_scores, real_images, do_r1_reg: bool, real_scores=None, r1_pre=None):
        """training_step.py:237-266"""
        if do_r1_reg:
            real_scores, r1_penalty = r1_pre if r1_pre is not None else self._r1_reg(real_images)
        else:
            if real_scores is None:
                real_scores = self.discriminator(real_images)
            r1_penalty = torch.zeros((), device=real_scores.device)
        d_loss = discriminator_loss(fake_scores, real_scores, self.batch_size)
        reg_d_loss = d_loss + r1_penalty
        return reg_d_loss, d_loss, r1_penalty

    def _get_generator_losses(self, fake_images, do_pl_reg: bool, input_words, draws: dict, fake_scores=None,
                              pl_pre=None):
        """training_step.py:268-298"""
        if fake_scores is None:
            fake_scores = self.discriminator(fake_images)
        g_loss = generator_loss(fake_scores, self.batch_size)
        if pl_pre is not None:
            pl_penalty = pl_pre
        else:
            pl_penalty = self._path_length_reg(input_words, draws) if do_pl_reg \
                else torch.zeros((), device=fake_scores.device)
        reg_g_loss = g_loss + pl_penalty
        return fake_scores, reg_g_loss, g_loss, pl_penalty

    def _path_length_reg(self, input_words, draws: dict):
        """training_step.py:300-347"""
        G = self.generator
        dev = G.device
        pl_minibatch = max(1, self.batch_size_per_gpu // self.pl_minibatch_shrink)
        pl_z = draws["pl_z"].to(dev) if "pl_z" in draws else torch.randn(pl_minibatch, self.z_dim, device=dev)
        pl_draws = {"noises": draws["pl_noises"]} if "pl_noises" in draws else {}
        # generator(...) with the default training=False (:325-329)
        with L.double_backward():
            pl_fake_images, pl_style = G((input_words[:pl_minibatch], pl_z), batch_size=pl_minibatch, ret_style=True,
                                         draws=pl_draws)
        noise = draws["pl_image_noise"].to(dev) if "pl_image_noise" in draws else torch.randn_like(pl_fake_images)
        pl_noise = noise * self.pl_noise_scaler
        pl_noise_applied = (pl_fake_images * pl_noise).sum()
        (pl_grads,) = torch.autograd.grad(pl_noise_applied, pl_style, create_graph=True)    # :333
        pl_lengths = torch.sqrt((pl_grads ** 2).sum(dim=2).mean(dim=1))                      # :334-336
        with torch.no_grad():                                                                # :338-341
            self.pl_mean.copy_(self.pl_mean + self.pl_decay * (pl_lengths.mean() - self.pl_mean))
        pl_penalty = (pl_lengths - self.pl_mean) ** 2                                        # :344
        pl_penalty = pl_penalty * self.pl_minibatch_shrink * self.g_reg_interval             # :346
        return pl_penalty.sum() / self.batch_size                                            # :347

    def _r1_reg(self, real_images):
        """training_step.py:349-373"""
        real_images = real_images.detach().requires_grad_(True)
        with L.double_backward():
            real_scores = self.discriminator(real_images)
        real_loss = real_scores.sum()
        (real_grads,) = torch.autograd.grad(real_loss, real_images, create_graph=True)
        if real_grads.is_cuda:
            from .fused import R1SqNorm

            r1_penalty = R1SqNorm.apply(real_grads)[:, None]                                 # one reduction launch
        else:
            r1_penalty = (real_grads ** 2).sum(dim=(1, 2, 3))[:, None]
        r1_penalty = r1_penalty * (0.5 * self.r1_gamma) * self.d_reg_interval
        r1_penalty = r1_penalty.sum() / self.batch_size
        return real_scores, r1_penalty

    def _get_ocr_loss(self, fake_images, ocr_labels, ocr_images):
        """training_step.py:375-402"""
        fake_images_ocr_format = self.aster_ocr.convert_inputs(fake_images, ocr_labels, blank_label=1, cfg=self.cfg)
        logits = self.aster_ocr(fake_images_ocr_format)
        if self.ocr_loss_type == "mse":
            real_logits = self.aster_ocr(ocr_images)
            return mean_squared_loss(real_logits, logits, self.batch_size)
        elif self.ocr_loss_type == "softmax_crossentropy":
            return softmax_cross_entropy_loss(logits, ocr_labels, self.batch_size)
