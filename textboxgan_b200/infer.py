"""Infer — generate text boxes for chosen words / score a test corpus with the EMA generator (mirror of
infer.py:26-134).  Same method names (including the reference's spelling
``genererate_chosen_words``), same behaviour: one latent shared by all words of a call, optional fixed style
vector (e.g. from the projector), truncation psi, images cropped to ``char_width * len(word)`` (NOT masked), written
with cv2 as BGR PNGs."""
from __future__ import annotations

import os
from fractions import Fraction
from typing import List, Optional

import numpy as np
import torch

from .aster_inferer import AsterInferer
from .config import Config, cfg as default_cfg
from .data_loader import ValidationDataLoader
from .loss_tracker import LossTracker
from .model_loader import ModelLoader
from .utils import generator_output_to_uint8, string_to_main_int_sequence
from .validation_step import ValidationStep


class Infer:
    """Infer the trained model"""

    def __init__(self, cfg: Optional[Config] = None, device="cuda", ckpt_dir: Optional[str] = None, generator=None,
                 printer=print):
        self.cfg = cfg if cfg is not None else default_cfg
        self.device = device
        self.generator = generator if generator is not None else ModelLoader(self.cfg, device=device).load_generator(
            is_g_clone=True, ckpt_dir=ckpt_dir)                                          # infer.py:30-32
        self.aster_ocr = AsterInferer(self.cfg, device=device, synthetic_weights=bool(self.cfg.aster_synthetic_weights))
        self.test_step = ValidationStep(self.generator, self.aster_ocr, self.cfg)
        self.strategy = self.cfg.strategy
        self._print = printer

    def _crop(self, n_chars: int) -> int:
        return int(Fraction(self.cfg.char_width) * n_chars)

    @torch.no_grad()
    def generate(self, words_list: List[str], w_latents: Optional[torch.Tensor] = None, truncation_psi: float = 1.0,
                 z: Optional[torch.Tensor] = None) -> np.ndarray:
        """uint8 HWC images ``[N, H, W, 3]`` for ``words_list`` (infer.py:62-86)."""
        cfg, G = self.cfg, self.generator
        n = len(words_list)
        words = torch.from_numpy(string_to_main_int_sequence(words_list, cfg.max_char_number)).to(self.device)
        if w_latents is not None:
            x = G._word_encoder(words, n, None)
            style = w_latents.to(self.device).float().reshape(1, 1, -1).expand(n, G.n_style, -1).contiguous()
            fake_images = G._synthesis(x, style, G._noises(n, {}), fused_epilogue=True)
        else:
            if z is None:
                z = torch.randn(1, cfg.z_dim, device=self.device)
            fake_images = G((words, z.to(self.device).expand(n, -1).contiguous()), training=False,
                            truncation_psi=truncation_psi, batch_size=n)
        return generator_output_to_uint8(fake_images).cpu().numpy()

    def genererate_chosen_words(self, words_list: List[str], prefix: str, output_dir: str, do_sentence: bool,
                                w_latents=None, truncation_psi: float = 1.0) -> List[str]:
        """infer.py:36-104.  Returns the paths written."""
        import cv2

        fake_images = self.generate(words_list, w_latents, truncation_psi)
        os.makedirs(output_dir, exist_ok=True)
        written = []
        if do_sentence:
            sentence_image = np.concatenate([img[:, : self._crop(len(w))] for img, w in zip(fake_images, words_list)], axis=1)
            path = os.path.join(output_dir, f"{prefix}_sentence_image.png")
            cv2.imwrite(path, sentence_image)
            written.append(path)
        else:
            for image, word in zip(fake_images, words_list):
                path = os.path.join(output_dir, f"{prefix}_{word}_image.png")
                cv2.imwrite(path, image[:, : self._crop(len(word))])
                written.append(path)
        return written

    def infer_test_set(self, num_test_set_runs: int, text_corpus_dir: str, file_name: str = "test_corpus.txt") -> float:
        """infer.py:106-134: mean OCR loss of the test corpus over several latent draws."""
        loader = ValidationDataLoader(self.cfg, text_corpus_dir, file_name, device=self.device)
        global_tracker = LossTracker(["test_ocr_loss"], printer=self._print)
        for _ in range(num_test_set_runs):
            tracker = LossTracker(["test_ocr_loss"], printer=self._print)
            step = 0
            ds = loader.load_dataset(batch_size=self.cfg.batch_size)
            if self.strategy is not None:
                ds = self.strategy.experimental_distribute_dataset(ds)
            for step, (input_words, ocr_labels) in enumerate(ds):
                tracker.increment_losses({"test_ocr_loss": self.test_step.dist_validation_step(input_words, ocr_labels)})
            tracker.print_losses(step)
            global_tracker.increment_losses({"test_ocr_loss": tracker.losses["test_ocr_loss"].result()})
        self._print("_________AVERAGE TEST LOSS___________")
        global_tracker.print_losses(step=num_test_set_runs)
        return global_tracker.losses["test_ocr_loss"].result()
