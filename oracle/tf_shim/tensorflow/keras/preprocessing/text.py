"""keras_preprocessing.text.Tokenizer — the subset config/char_tokens.py:12-17 uses, restated from
the Keras source: ``fit_on_texts`` counts tokens (chars when char_level), ``word_index`` lists the
OOV token first and then tokens by descending count (stable, i.e. first-seen order on ties), from 1."""
from collections import OrderedDict


class Tokenizer:
    def __init__(self, num_words=None, filters="", lower=True, split=" ", char_level=False, oov_token=None, **_):
        self.char_level, self.lower, self.oov_token = char_level, lower, oov_token
        self.word_counts = OrderedDict()
        self.word_index = {}
        self.index_word = {}

    def fit_on_texts(self, texts):
        for text in texts:
            if self.lower:
                text = text.lower()
            seq = list(text) if self.char_level else text.split()
            for w in seq:
                self.word_counts[w] = self.word_counts.get(w, 0) + 1
        wcounts = sorted(self.word_counts.items(), key=lambda kv: kv[1], reverse=True)  # stable
        vocab = ([self.oov_token] if self.oov_token is not None else []) + [w for w, _ in wcounts]
        self.word_index = dict(zip(vocab, range(1, len(vocab) + 1)))
        self.index_word = {i: w for w, i in self.word_index.items()}

    def texts_to_sequences(self, texts):
        oov = self.word_index.get(self.oov_token)
        out = []
        for text in texts:
            if self.lower:
                text = text.lower()
            seq = list(text) if self.char_level else text.split()
            vect = []
            for w in seq:
                i = self.word_index.get(w)
                if i is not None:
                    vect.append(i)
                elif oov is not None:
                    vect.append(oov)
            out.append(vect)
        return out
