"""Trainer — the reference's training loop around the B200 training step (mirror of train.py:22-261; SURVEY.md §8f
row f1).  Same attribute names, same schedule:

* lazy regularisation: ``do_r1_reg`` / ``do_pl_reg`` on every ``reg_interval``-th step (train.py:182-183),
* OCR-loss warm-up: weight ``1e-8`` until step 5000, ``cfg.ocr_loss_weight`` afterwards (:185-192),
* EMA clone update after every step (:208), ``LossTracker`` s per print frequency (:165-171,248-255),
* checkpoint every ``save_step_frequency`` steps and at the end with keep-N (:228-229,259-261),
* validation every ``validation_step_frequency`` steps (:236-247).

Differences that follow from the B200 design: one process per GPU (launch with torchrun; ``cfg.strategy`` is the
shim of strategy.py, which hands every rank its shard of each global batch), datasets are any iterable of GLOBAL-batch
``(real_images, ocr_image, input_words, ocr_labels)`` tensor tuples with the loader contract of
dataset_utils/training_data_loader.py:56-97 (the cv2 / lmdb input pipeline is row f2, not built), TensorBoard logging is optional (``scalar_writer`` callable) and image summaries are not built.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, Optional

import torch

from .aster_inferer import AsterInferer
from .config import Config, cfg as default_cfg
from .loss_tracker import LossTracker
from .model_loader import ModelLoader
from .optimizers import Adam, update_optimizer_params
from .strategy import Strategy
from .training_step import TrainingStep
from .validation_step import ValidationStep

TRAIN_LOSSES = ["reg_g_loss", "g_loss", "pl_penalty", "ocr_loss", "reg_d_loss", "d_loss", "r1_penalty"]   # train.py:153-161


class Trainer:
    """Train the model. The different configs can be tuned in config."""

    # train.py:185-192: the OCR loss is (nearly) switched off while the generator cannot write yet
    ocr_warmup_steps = 5000
    ocr_warmup_weight = 1e-8

    def __init__(self, cfg: Optional[Config] = None, device="cuda", *, train_dataset: Optional[Iterable] = None,
                 validation_dataset: Optional[Iterable] = None, ckpt_dir: Optional[str] = None,
                 scalar_writer: Optional[Callable[[dict, int], None]] = None, use_cuda_graph: bool = True,
                 printer=print):
        cfg = cfg if cfg is not None else default_cfg
        self.cfg = cfg
        if cfg.strategy is None:
            cfg.attach_strategy(Strategy())
        self.batch_size = cfg.batch_size
        self.strategy = cfg.strategy
        self.max_steps = cfg.max_steps
        self.summary_steps_frequency = cfg.summary_steps_frequency
        self.save_step_frequency = cfg.save_step_frequency
        self.validation_step_frequency = cfg.validation_step_frequency
        self._print = printer
        self.scalar_writer = scalar_writer
        # set optimizer params (train.py:38-39,110-129)
        self.g_opt = self.update_optimizer_params(cfg.g_opt)
        self.d_opt = self.update_optimizer_params(cfg.d_opt)
        self.pl_mean = torch.zeros((), device=device)                                            # train.py:40-46
        self.training_dataset = train_dataset
        self.validation_dataset = validation_dataset
        self.model_loader = ModelLoader(cfg, device=device)
        self.discriminator, self.generator, self.g_clone = self.model_loader.initiate_models()   # train.py:49-55
        mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
        self.d_optimizer, self.g_optimizer, self.ocr_optimizer = mk(self.d_opt), mk(self.g_opt), mk(self.g_opt)
        self.ocr_loss_weight = cfg.ocr_loss_weight
        self.aster_ocr = AsterInferer(cfg, device=device, synthetic_weights=bool(cfg.aster_synthetic_weights))
        self.training_step = TrainingStep(self.generator, self.discriminator, self.aster_ocr, self.g_optimizer,
                                          self.ocr_optimizer, self.d_optimizer, self.g_opt["reg_interval"],
                                          self.d_opt["reg_interval"], self.pl_mean, cfg)           # train.py:80-90
        self.training_step.use_cuda_graph = bool(use_cuda_graph) and torch.device(device).type == "cuda"
        self.validation_step = ValidationStep(self.g_clone, self.aster_ocr, cfg)
        self.manager = None
        if ckpt_dir is not None:
            self.manager = self.model_loader.load_checkpoint(                                     # train.py:94-108
                ckpt_kwargs={"d_optimizer": self.d_optimizer, "g_optimizer": self.g_optimizer,
                             "ocr_optimizer": self.ocr_optimizer, "discriminator": self.discriminator,
                             "generator": self.generator, "g_clone": self.g_clone, "pl_mean": self.pl_mean},
                model_description="Full model", expect_partial=False, ckpt_dir=ckpt_dir,
                max_to_keep=cfg.num_ckpts_to_keep)

    update_optimizer_params = staticmethod(update_optimizer_params)

    # -- the schedule of train.py:178-192 as pure functions of the step counter ------------------------------------
    def regularisation_flags(self, step: int):
        do_r1_reg = (step + 1) % self.d_opt["reg_interval"] == 0
        do_pl_reg = (step + 1) % self.g_opt["reg_interval"] == 0
        return do_r1_reg, do_pl_reg

    def ocr_weight(self, step: int) -> float:
        return self.ocr_loss_weight if step > self.ocr_warmup_steps else self.ocr_warmup_weight

    def _save(self, step: int) -> None:
        if self.manager is not None and self.strategy.rank == 0:
            self.manager.save(checkpoint_number=step)

    # -- pieces of the loop body (train.py:178-256) ---------------------------------------------------------------
    def _one_step(self, batch) -> dict:
        """One generator/discriminator/OCR update + EMA; returns the seven tracked losses by name."""
        real_images, ocr_image, input_words, ocr_labels = batch
        step = self.g_optimizer.iterations.numpy()
        r1, pl = self.regularisation_flags(step)
        (reg_g, g, pl_pen), (reg_d, d, r1_pen), ocr = self.training_step.dist_train_step(
            real_images, ocr_image, input_words, ocr_labels, r1, pl, self.ocr_weight(step))
        self.g_clone.set_as_moving_average_of(self.generator)                                      # train.py:208
        # ONE device -> host read of the seven scalars per iteration (every tracker then works on Python floats)
        vals = torch.stack([torch.as_tensor(v, dtype=torch.float32, device=reg_g.device).reshape(())
                            for v in (reg_g, g, pl_pen, ocr, reg_d, d, r1_pen)]).tolist()
        return dict(zip(TRAIN_LOSSES, vals))

    def _validate(self, tracker: LossTracker, step: int) -> None:
        for words, labels in self.strategy.experimental_distribute_dataset(self.validation_dataset):
            tracker.increment_losses({"validation_ocr_loss": self.validation_step.dist_validation_step(words, labels)})
        self._log(tracker, step)
        tracker.print_losses(step)
        tracker.reinitialize_tracker()

    def _log(self, tracker: LossTracker, step: int) -> None:
        if self.scalar_writer is not None:
            self.scalar_writer({name: mean.result() for name, mean in tracker.losses.items()}, step)

    def train(self) -> int:
        """Main training loop (train.py:131-261).  Returns the number of generator updates done."""
        if self.training_dataset is None:
            raise ValueError("Trainer.train() needs a train_dataset iterable")
        self._print("Start Training")
        replicas = self.strategy.num_replicas_in_sync
        freq = self.summary_steps_frequency
        trackers = [LossTracker(TRAIN_LOSSES, every, log, num_replicas=replicas, printer=self._print)
                    for every, log in zip(freq["print_steps"], freq["log_losses"])]
        val_tracker = LossTracker(["validation_ocr_loss"], num_replicas=replicas, printer=self._print)
        from .prefetch import DevicePrefetcher

        # this rank's shards, copied to the device one batch ahead of the iteration that consumes them
        shards = DevicePrefetcher(self.strategy.experimental_distribute_dataset(self.training_dataset),
                                  self.generator.device)
        for batch in shards:
            losses = self._one_step(batch)
            step = self.g_optimizer.iterations.numpy()                     # counts generator updates (train.py:211)
            for tracker in trackers:
                tracker.increment_losses(losses)
            if step % self.save_step_frequency == 0:
                self._save(step)
            if self.validation_dataset is not None and step % self.validation_step_frequency == 0:
                self._validate(val_tracker, step)
            for tracker in trackers:
                if step % tracker.print_step == 0:
                    tracker.print_losses(step)
                    if tracker.log_losses:
                        self._log(tracker, step)
                    tracker.reinitialize_tracker()
            if step == self.max_steps:
                break
        last = self.g_optimizer.iterations.numpy()
        self._save(last)                                                   # train.py:259-261
        return last


def synthetic_dataset(cfg: Config, n_batches: int, device="cuda", seed: Optional[int] = None):
    """Seeded stand-in for TrainingDataLoader.load_dataset (dataset_utils/training_data_loader.py:56-97): random words of
    length 1..max_char_number, their ASTER labels, uniform-noise "real" images zeroed right of the word, and the scalar
    0.0 ocr image of softmax-cross-entropy mode.  Global batches (``cfg.batch_size``); the strategy shards them."""
    from .char_tokens import main_to_aster_ids
    from .utils import mask_text_box

    g = torch.Generator().manual_seed(cfg.shuffle_seed if seed is None else seed)
    B, mcn = cfg.batch_size, cfg.max_char_number
    zero = torch.zeros((), device=device)
    for _ in range(n_batches):
        lens = torch.randint(1, mcn + 1, (B,), generator=g)
        chars = torch.randint(1, 70, (B, mcn), generator=g)
        words = torch.where(torch.arange(mcn)[None, :] < lens[:, None], chars, torch.zeros_like(chars)).to(torch.int32)
        labels = torch.from_numpy(main_to_aster_ids(words.numpy())).to(torch.int32)
        real = mask_text_box(torch.rand(B, 3, cfg.char_height, cfg.image_width, generator=g) * 2 - 1, words, cfg.char_width)
        yield real.to(device), zero, words.to(device), labels.to(device)


if __name__ == "__main__":   # synthetic-data smoke run: python -m textboxgan_b200.train [steps]
    import sys

    from .config import baseline_config

    c = baseline_config(1)
    c.max_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    t = Trainer(c, train_dataset=synthetic_dataset(c, c.max_steps), ckpt_dir=os.environ.get("TBG_CKPT_DIR"))
    t.train()
