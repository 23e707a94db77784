"""Training / validation data loaders (mirror of dataset_utils/training_data_loader.py:13-97 and
dataset_utils/validation_data_loader.py:11-51) producing the GLOBAL-batch tensor tuples
``Trainer`` consumes.  The per-sample transform is the reference's, line for line:

    image = cv2.imread(...)                                   BGR uint8
    main  = cv2.resize(image, (char_width*len(word), char_height)) / 127.5 - 1        (:64-65)
    pad right with zeros to max_char_number*char_width, HWC -> CHW                    (:75-86)
    with probability 0.25 (softmax-cross-entropy mode) the WORD (not the image) is replaced by the next corpus
    word (:88-92); words -> main ids (pad 0) and ASTER ids (pad 1)                   (:94-95)

What differs: the reference streams through ``tf.data`` (``.repeat().shuffle(seed).batch(drop_remainder)``); here the
same pipeline is a Python generator over a seeded shuffle buffer — the sample ORDER is therefore not TensorFlow's
(its shuffle RNG is not reproducible outside TF), the sample CONTENT is.  For a fractional ``char_width`` (BASELINE
configs 2/3, see config.py) widths are rounded down: ``floor(char_width * len(word))``.
"""
from __future__ import annotations

import os
import random
from fractions import Fraction
from typing import Iterator, List, Optional, Tuple

import numpy as np
import torch

from .config import Config
from .utils import string_to_aster_int_sequence, string_to_main_int_sequence


class TrainingDataLoader:
    """Loads the dataset which is used for training."""

    def __init__(self, cfg: Config, text_boxes_dir: str, text_corpus_dir: Optional[str] = None, device="cpu",
                 seed: Optional[int] = None, device_transform: bool = False):
        """``device_transform``: decode on the host, then resize / normalise / pad / transpose the WHOLE batch in one
        kernel launch on ``device`` (``tbg_batch_resize_normalize``) instead of per-sample cv2 calls — at 10-20 K images/s
        the per-sample host transform cannot feed the training step."""
        self.cfg = cfg
        self.text_boxes_dir = text_boxes_dir
        self.device = device
        self.device_transform = bool(device_transform)
        self.return_ocr_image = cfg.ocr_loss_type == "mse"                         # :17
        self.use_corpus_word = cfg.ocr_loss_type == "softmax_crossentropy"         # :18
        self.corpus_words: List[str] = []
        if text_corpus_dir is not None:
            with open(os.path.join(text_corpus_dir, "train_corpus.txt"), "r") as f:
                self.corpus_words = f.readlines()
        self.corpus_words_generator = iter(self.corpus_words)
        self.corpus_word_ratio = 0.25                                              # :24
        self._rng = random.Random(cfg.shuffle_seed if seed is None else seed)

    # -- one sample (dataset_utils/training_data_loader.py:56-97) ---------------------------------------------------
    def _word_width(self, n_chars: int) -> int:
        return int(Fraction(self.cfg.char_width) * n_chars)

    def _data_getter(self, data: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        import cv2

        cfg = self.cfg
        image_name, word = data.split(",", 1)
        word = word.strip("\n")
        image = cv2.imread(os.path.join(self.text_boxes_dir, image_name))
        if image is None:
            raise FileNotFoundError(os.path.join(self.text_boxes_dir, image_name))
        raw_image, raw_width = image, self._word_width(len(word))
        if self.device_transform:
            main_image = None                      # resized on the device, batch at a time (_collate)
        else:
            main_image = cv2.resize(image, (raw_width, cfg.char_height))
            main_image = main_image.astype(np.float32) / 127.5 - 1.0
        if self.return_ocr_image:
            ocr_image = cv2.resize(image, (cfg.aster_image_dims[1], cfg.aster_image_dims[0]))
            ocr_image = ocr_image.astype(np.float32) / 127.5 - 1.0
        else:
            ocr_image = np.float32(0.0)
        if self.device_transform:
            padded_image = (np.ascontiguousarray(raw_image), raw_width)
        else:
            padding_length = cfg.image_width - main_image.shape[1]     # == (max_char_number - len(word)) * char_width
            padded_image = cv2.copyMakeBorder(src=main_image, top=0, bottom=0, left=0, right=padding_length,
                                              borderType=cv2.BORDER_CONSTANT)
            padded_image = np.transpose(padded_image, (2, 0, 1))       # H,W,C to C,H,W
        if self.use_corpus_word and self.corpus_words and self._rng.random() > 1 - self.corpus_word_ratio:
            word = next(self.corpus_words_generator, None)
            if word is None:
                self.corpus_words_generator = iter(self.corpus_words)
                word = next(self.corpus_words_generator)
            # NOTE: as in the reference the corpus line keeps its trailing newline, which tokenises to OOV (= 0 / 1)
        input_word_array = string_to_main_int_sequence([word], cfg.max_char_number)[0]
        ocr_label_array = string_to_aster_int_sequence([word], cfg.max_char_number)[0]
        return padded_image, ocr_image, input_word_array, ocr_label_array

    # -- the stream -----------------------------------------------------------------------------------------------
    def load_dataset(self, batch_size: int, buffer_size: int = -1, repeat: bool = True) -> Iterator[tuple]:
        """``.repeat().shuffle(buffer).batch(batch_size, drop_remainder=True)`` as a generator of tensor tuples
        ``(real_images [B,3,H,W] f32, ocr_images [B,64,256,3] f32 | scalar 0.0, input_words [B,mcn] i32,
        ocr_labels [B,mcn] i32)``."""
        with open(os.path.join(self.text_boxes_dir, "annotations_filtered.txt"), "r") as f:
            lines = f.readlines()
        buf_n = len(lines) if buffer_size == -1 else buffer_size

        def samples():
            while True:
                for ln in lines:
                    yield ln
                if not repeat:
                    return

        buf: List[str] = []
        batch: List[tuple] = []
        src = samples()
        exhausted = False
        while True:
            while not exhausted and len(buf) < max(1, buf_n):
                try:
                    buf.append(next(src))
                except StopIteration:
                    exhausted = True
            if not buf:
                return
            batch.append(self._data_getter(buf.pop(self._rng.randrange(len(buf)))))
            if len(batch) == batch_size:
                yield self._collate(batch)
                batch = []

    def _device_images(self, items: List[tuple]) -> torch.Tensor:
        """[(uint8 HWC BGR image, target width)] -> fp32 [B,3,H,W] on the device: one packed host buffer, one
        host-to-device copy, one kernel launch."""
        from . import kernels as K

        cfg, dev = self.cfg, self.device
        sizes = [im.size for im, _ in items]
        offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        packed = torch.empty(int(sum(sizes)), dtype=torch.uint8)
        if torch.device(dev).type == "cuda":
            packed = packed.pin_memory()
        flat = packed.numpy()
        for (im, _), off, n in zip(items, offsets, sizes):
            flat[off: off + n] = im.reshape(-1)
        meta = lambda v, dt: torch.as_tensor(np.asarray(v), dtype=dt).to(dev, non_blocking=True)
        return K.batch_resize_normalize(packed.to(dev, non_blocking=True), meta(offsets, torch.int64),
                                        meta([im.shape[0] for im, _ in items], torch.int32),
                                        meta([im.shape[1] for im, _ in items], torch.int32),
                                        meta([w for _, w in items], torch.int32), cfg.char_height, cfg.image_width)

    def _collate(self, batch: List[tuple]) -> tuple:
        dev = self.device
        if self.device_transform:
            real = self._device_images([b[0] for b in batch])
        else:
            real = torch.from_numpy(np.stack([b[0] for b in batch])).to(dev)
        ocr = torch.from_numpy(np.stack([b[1] for b in batch])).to(dev) if self.return_ocr_image \
            else torch.zeros((), device=dev)
        words = torch.from_numpy(np.stack([b[2] for b in batch]).astype(np.int32)).to(dev)
        labels = torch.from_numpy(np.stack([b[3] for b in batch]).astype(np.int32)).to(dev)
        return real, ocr, words, labels


class ValidationDataLoader:
    """Loads the dataset which is used for validation and testing (words only)."""

    def __init__(self, cfg: Config, text_corpus_dir: str, file_name: str, device="cpu"):
        self.cfg, self.path, self.device = cfg, os.path.join(text_corpus_dir, file_name), device

    def _data_getter(self, data: str) -> Tuple[np.ndarray, np.ndarray]:
        word = data.strip("\n")
        return (string_to_main_int_sequence([word], self.cfg.max_char_number)[0],
                string_to_aster_int_sequence([word], self.cfg.max_char_number)[0])

    def load_dataset(self, batch_size: int) -> Iterator[tuple]:
        with open(self.path, "r") as f:
            words = f.readlines()
        for i in range(0, len(words) - batch_size + 1, batch_size):            # drop_remainder=True
            pairs = [self._data_getter(w) for w in words[i: i + batch_size]]
            yield (torch.from_numpy(np.stack([p[0] for p in pairs]).astype(np.int32)).to(self.device),
                   torch.from_numpy(np.stack([p[1] for p in pairs]).astype(np.int32)).to(self.device))
