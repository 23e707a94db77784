"""Character tokenisation, bit-exact integer path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates config/char_tokens.py:4-17 and utils/utils.py:66-105 without Keras:
``Tokenizer(char_level=True, lower=False, oov_token="<OOV>").fit_on_texts(VECTOR)`` gives every
character count 1, so ``word_index`` follows first-occurrence order with ``<OOV>`` = 1 and the
characters from 2 (keras_preprocessing/text.py: ``sorted_voc = [oov] + words sorted by count,
stable``).  ``pad_sequences(..., maxlen, value=1, padding="post")`` uses the default
``truncating="pre"`` (keeps the LAST ``maxlen`` tokens).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

# config/char_tokens.py:4-6
MAIN_CHAR_VECTOR = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ-'.!?,\""
# config/char_tokens.py:9
ASTER_CHAR_VECTOR = (
    "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~"
)


def build_word_index(char_vector: str) -> Dict[str, int]:
    """config/char_tokens.py:14-17 — Keras char-level tokenizer fitted on the vector string."""
    index: Dict[str, int] = {"<OOV>": 1}
    for ch in char_vector:  # each char is one "text"; all counts equal -> insertion order
        if ch not in index:
            index[ch] = len(index) + 1
    return index


MAIN_WORD_INDEX = build_word_index(MAIN_CHAR_VECTOR)
ASTER_WORD_INDEX = build_word_index(ASTER_CHAR_VECTOR)


def texts_to_sequences(words: List[str], word_index: Dict[str, int]) -> List[List[int]]:
    """Keras ``Tokenizer.texts_to_sequences`` with an OOV token: unknown chars map to index 1."""
    return [[word_index.get(ch, 1) for ch in w] for w in words]


def pad_sequences(seqs: List[List[int]], maxlen: int, value: int) -> np.ndarray:
    """``pad_sequences(seqs, maxlen=maxlen, value=value, padding="post")`` (truncating="pre")."""
    out = np.full((len(seqs), maxlen), value, dtype=np.int32)
    for i, s in enumerate(seqs):
        s = s[-maxlen:] if maxlen > 0 else []
        out[i, : len(s)] = np.asarray(s, dtype=np.int32)
    return out


def string_to_main_int_sequence(words: List[str], max_char_number: int) -> np.ndarray:
    """utils/utils.py:66-85 — pad=0, OOV=0, '0'=1 ... '\"'=69."""
    seq = texts_to_sequences(words, MAIN_WORD_INDEX)
    return pad_sequences(seq, max_char_number, value=1) - 1


def string_to_aster_int_sequence(words: List[str], max_char_number: int) -> np.ndarray:
    """utils/utils.py:88-105 — pad=1 (also the OOV id), '0'=2 ... '~'=95."""
    seq = texts_to_sequences(words, ASTER_WORD_INDEX)
    return pad_sequences(seq, max_char_number, value=1)


def main_to_aster_ids(main_ids: np.ndarray) -> np.ndarray:
    """Map main-vocabulary ids (0 pad, 1..69) to ASTER ids (1 pad, 2..95) char by char."""
    lut = np.ones(len(MAIN_CHAR_VECTOR) + 1, dtype=np.int32)
    for i, ch in enumerate(MAIN_CHAR_VECTOR):
        lut[i + 1] = ASTER_WORD_INDEX[ch]
    return lut[main_ids]
