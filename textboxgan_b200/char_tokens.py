"""Character vocabularies and tokenisers (mirror of config/char_tokens.py:4-17 and the string
helpers of utils/utils.py:66-105) without the Keras dependency.

``Tokenizer(char_level=True, lower=False, oov_token="<OOV>").fit_on_texts(VECTOR)``: every
character occurs once, so indices follow string order with ``<OOV>`` = 1 and characters from 2.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

MAIN_CHAR_VECTOR = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ-'.!?,\""
ASTER_CHAR_VECTOR = (
    "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~"
)


class _CharLevelTokenizer:
    """The subset of ``keras_preprocessing.text.Tokenizer`` the reference uses."""

    def __init__(self, vector: str, oov_token: str = "<OOV>"):
        self.oov_token = oov_token
        self.word_index: Dict[str, int] = {oov_token: 1}
        for ch in vector:
            if ch not in self.word_index:
                self.word_index[ch] = len(self.word_index) + 1
        self.index_word = {v: k for k, v in self.word_index.items()}

    def texts_to_sequences(self, texts: List[str]) -> List[List[int]]:
        oov = self.word_index[self.oov_token]
        return [[self.word_index.get(ch, oov) for ch in t] for t in texts]

    def sequences_to_texts(self, seqs) -> List[str]:
        return [" ".join(self.index_word.get(int(i), self.oov_token) for i in s if int(i) in self.index_word) for s in seqs]


class CharTokenizer:
    """config/char_tokens.py:12-17"""

    def __init__(self):
        self.main = _CharLevelTokenizer(MAIN_CHAR_VECTOR)
        self.aster = _CharLevelTokenizer(ASTER_CHAR_VECTOR)


def main_to_aster_ids(main_ids: np.ndarray) -> np.ndarray:
    """Main-vocabulary ids (0 = pad/OOV, '0' = 1 ... '"' = 69; utils/utils.py:80-85) -> ASTER ids of the same
    characters (1 = pad, '0' = 2 ... '~' = 95; utils/utils.py:102-105)."""
    tok = CharTokenizer()
    lut = np.ones(len(MAIN_CHAR_VECTOR) + 1, dtype=np.int32)
    for i, ch in enumerate(MAIN_CHAR_VECTOR):
        lut[i + 1] = tok.aster.word_index[ch]
    return lut[np.asarray(main_ids, dtype=np.int64)]


def pad_sequences(seqs: List[List[int]], maxlen: int, value: int) -> np.ndarray:
    """keras ``pad_sequences(..., padding="post")`` with the default ``truncating="pre"``."""
    out = np.full((len(seqs), maxlen), value, dtype=np.int32)
    for i, s in enumerate(seqs):
        s = s[-maxlen:]
        out[i, : len(s)] = np.asarray(s, dtype=np.int32)
    return out
