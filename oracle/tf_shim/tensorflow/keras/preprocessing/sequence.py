"""keras_preprocessing.sequence.pad_sequences restated (defaults: dtype int32, padding/truncating 'pre')."""
import numpy as np


def pad_sequences(sequences, maxlen=None, dtype="int32", padding="pre", truncating="pre", value=0.0):
    lengths = [len(s) for s in sequences]
    if maxlen is None:
        maxlen = max(lengths) if lengths else 0
    x = np.full((len(sequences), maxlen), value, dtype=dtype)
    for i, s in enumerate(sequences):
        if not len(s):
            continue
        trunc = s[-maxlen:] if truncating == "pre" else s[:maxlen]
        trunc = np.asarray(trunc, dtype=dtype)
        if padding == "post":
            x[i, : len(trunc)] = trunc
        else:
            x[i, -len(trunc):] = trunc
    return x
