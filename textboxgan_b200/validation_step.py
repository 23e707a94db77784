"""ValidationStep — OCR loss of the EMA generator on validation words (mirror of
validation_step.py:10-90)."""
from __future__ import annotations

from typing import Optional

import torch

from .aster_inferer import AsterInferer
from .config import Config
from .generator import Generator
from .losses import softmax_cross_entropy_loss
from .utils import mask_text_box


class ValidationStep:
    def __init__(self, generator: Generator, aster_ocr: AsterInferer, cfg: Optional[Config] = None):
        self.cfg = cfg if cfg is not None else generator.cfg
        self.generator = generator
        self.aster_ocr = aster_ocr
        self.batch_size = self.cfg.batch_size
        self.batch_size_per_gpu = self.cfg.batch_size_per_gpu
        self.z_dim = self.cfg.z_dim
        self.char_width = self.cfg.char_width

    def dist_validation_step(self, input_words, ocr_labels):
        """validation_step.py:24-55"""
        strategy = self.cfg.strategy
        if strategy is None:
            return self._validation_step(input_words, ocr_labels)
        ocr_loss = strategy.run(self._validation_step, args=(input_words, ocr_labels))
        return strategy.reduce("SUM", ocr_loss, axis=None)

    @torch.no_grad()
    def _validation_step(self, input_words, ocr_labels, z: Optional[torch.Tensor] = None, draws: Optional[dict] = None):
        """validation_step.py:57-90 (``z`` / ``draws`` inject the latent and the per-layer noise for parity tests)."""
        dev = self.generator.device
        if z is None:
            z = torch.randn(input_words.shape[0], self.z_dim, device=dev)
        fake_images = self.generator((input_words, z), training=False, draws=draws)
        fake_images = mask_text_box(fake_images, input_words, self.char_width)
        ocr_input_image = self.aster_ocr.convert_inputs(fake_images, ocr_labels, blank_label=1, cfg=self.cfg)
        logits = self.aster_ocr(ocr_input_image)
        return softmax_cross_entropy_loss(logits, ocr_labels, self.batch_size)
