"""Host/device helpers of the training step (mirror of utils/utils.py)."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from .char_tokens import CharTokenizer, pad_sequences

_TOKENIZER = CharTokenizer()


def mask_text_box(fake_images: torch.Tensor, input_words: torch.Tensor, char_width: int) -> torch.Tensor:
    """utils/utils.py:11-45 — zero the image columns that belong to pad tokens.

    ``mask[b, x] = input_words[b, x // char_width] != 0`` for NCHW ``fake_images``."""
    keep = (input_words != 0).to(fake_images.dtype)
    if isinstance(char_width, int):
        mask = torch.repeat_interleave(keep, char_width, dim=1)[:, None, None, :]
    else:   # fractional char_width (config.baseline_config): column x belongs to character floor(x / char_width)
        mask = keep[:, column_char_index(fake_images.shape[3], char_width, fake_images.device)][:, None, None, :]
    return fake_images * mask


def column_char_index(width: int, char_width, device=None) -> torch.Tensor:
    """floor(x / char_width) for x in [0, width) in exact integer arithmetic (char_width int or Fraction)."""
    from fractions import Fraction

    cw = Fraction(char_width)
    x = torch.arange(width, device=device, dtype=torch.long)
    return (x * cw.denominator) // cw.numerator


def crop_width(first_blank: torch.Tensor, char_width) -> torch.Tensor:
    """first_blank * char_width rounded down (aster_inferer.py:182), exact for int or Fraction widths."""
    from fractions import Fraction

    cw = Fraction(char_width)
    return (first_blank.long() * cw.numerator) // cw.denominator


def generator_output_to_uint8(fake_images: torch.Tensor) -> torch.Tensor:
    """utils/utils.py:48-63"""
    x = (torch.clamp(fake_images, -1.0, 1.0) + 1.0) * 127.5
    return x.permute(0, 2, 3, 1).to(torch.uint8)


def string_to_main_int_sequence(words_list: List[str], max_char_number: int) -> np.ndarray:
    """utils/utils.py:66-85 — pad = OOV = 0, '0' = 1 ... '"' = 69."""
    seq = _TOKENIZER.main.texts_to_sequences(words_list)
    return pad_sequences(seq, maxlen=max_char_number, value=1) - 1


def string_to_aster_int_sequence(words_list: List[str], max_char_number: int) -> np.ndarray:
    """utils/utils.py:88-105 — pad = 1, '0' = 2 ... '~' = 95."""
    seq = _TOKENIZER.aster.texts_to_sequences(words_list)
    return pad_sequences(seq, maxlen=max_char_number, value=1)
