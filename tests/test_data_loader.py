"""Row f2: the loader contract of dataset_utils/training_data_loader.py:56-97 on synthetic PNGs."""
import os

import numpy as np
import pytest
import torch

from common import fractional_cfg, small_cfg
from oracle import train_step as OT
from textboxgan_b200.data_loader import TrainingDataLoader, ValidationDataLoader

cv2 = pytest.importorskip("cv2")


def _make_dataset(tmp_path, words):
    boxes = tmp_path / "text_boxes"
    corpus = tmp_path / "text_corpus"
    boxes.mkdir()
    corpus.mkdir()
    rng = np.random.RandomState(0)
    lines = []
    for i, w in enumerate(words):
        img = rng.randint(0, 256, size=(20 + i, 30 + 7 * len(w), 3), dtype=np.uint8)
        cv2.imwrite(str(boxes / f"{i}.png"), img)
        lines.append(f"{i}.png,{w}\n")
    (boxes / "annotations_filtered.txt").write_text("".join(lines))
    (corpus / "train_corpus.txt").write_text("corpus\nwords\n")
    (corpus / "validation_corpus.txt").write_text("alpha\nbeta\ngamma\n")
    return str(boxes), str(corpus)


@pytest.mark.parametrize("frac", [False, True])
def test_training_sample_contract(tmp_path, frac):
    cfg = fractional_cfg(2) if frac else small_cfg(2)
    words = ["Hi", "a,b", "World!"]           # "a,b": only the first comma separates name and word (:60)
    boxes, corpus = _make_dataset(tmp_path, words)
    ld = TrainingDataLoader(cfg, boxes, None)                                    # no corpus: words are never swapped
    for i, w in enumerate(words):
        img, ocr, ids, lab = ld._data_getter(f"{i}.png,{w}\n")
        assert img.shape == (3, cfg.char_height, cfg.image_width) and img.dtype == np.float32
        wpx = int(cfg.char_width * len(w))
        assert float(np.abs(img[:, :, wpx:]).max()) == 0.0 and img.min() >= -1.0 and img.max() <= 1.0
        ref = cv2.resize(cv2.imread(os.path.join(boxes, f"{i}.png")), (wpx, cfg.char_height)).astype(np.float32) / 127.5 - 1.0
        assert np.array_equal(img[:, :, :wpx], ref.transpose(2, 0, 1))
        assert float(ocr) == 0.0                                                # softmax-cross-entropy mode (:73-74)
        assert ids.shape == (cfg.max_char_number,) and lab.shape == (cfg.max_char_number,)
        assert (ids[: len(w)] > 0).all() and (ids[len(w):] == 0).all() and (lab[len(w):] == 1).all()
        # the image is consistent with the training step's own mask of the same word
        masked = OT.mask_text_box(torch.from_numpy(img)[None], torch.from_numpy(ids)[None], cfg.char_width)[0]
        assert torch.equal(masked, torch.from_numpy(img))


def test_stream_batches_shuffle_and_corpus_swap(tmp_path):
    cfg = small_cfg(2)
    words = ["one", "two", "three", "four", "five"]
    boxes, corpus = _make_dataset(tmp_path, words)
    ld = TrainingDataLoader(cfg, boxes, corpus, seed=3)
    it = ld.load_dataset(batch_size=2)
    batches = [next(it) for _ in range(40)]
    real, ocr, ids, lab = batches[0]
    assert real.shape == (2, 3, cfg.char_height, cfg.image_width) and ocr.dim() == 0
    assert ids.dtype == torch.int32 and lab.dtype == torch.int32 and ids.shape == (2, cfg.max_char_number)
    # ~25 % of the words come from the corpus ("corpus\\n" / "words\\n": 6 or 5 characters + OOV newline)
    from textboxgan_b200.utils import string_to_main_int_sequence as enc
    corpus_ids = {tuple(enc([w], cfg.max_char_number)[0]) for w in ("corpus\n", "words\n")}
    swapped = sum(tuple(r.tolist()) in corpus_ids for b in batches for r in b[2])
    assert 8 <= swapped <= 35                                                  # 80 samples, p = 0.25
    # no repeat: the stream ends after the last full batch (drop_remainder)
    ld2 = TrainingDataLoader(cfg, boxes, None, seed=1)
    assert len(list(ld2.load_dataset(batch_size=2, repeat=False))) == 2
    v = ValidationDataLoader(cfg, corpus, "validation_corpus.txt")
    vb = list(v.load_dataset(batch_size=2))
    assert len(vb) == 1 and vb[0][0].shape == (2, cfg.max_char_number) and int(vb[0][1][0, 5]) == 1


def test_device_transform_matches_host_transform(tmp_path):
    """The batched transform (tbg_batch_resize_normalize semantics, emulated on the CPU here) against the host path
    ``cv2.resize(...)/127.5 - 1`` + right pad + HWC->CHW (training_data_loader.py:63-82).  cv2's fixed-point bilinear
    and the kernel's float bilinear may round a value to the neighbouring uint8 level: tolerance 1 LSB = 1/127.5."""
    from emu import emulated_kernels

    cfg = small_cfg(2)
    words = ["Hi", "a,b", "World!", "x"]
    boxes, corpus = _make_dataset(tmp_path, words)
    # one image at exactly 2x the target size: cv2 switches INTER_LINEAR to the 2x2 area mean there
    w2 = int(cfg.char_width * 2)
    rng = np.random.RandomState(5)
    cv2.imwrite(os.path.join(boxes, "4.png"), rng.randint(0, 256, size=(2 * cfg.char_height, 2 * w2, 3), dtype=np.uint8))
    with open(os.path.join(boxes, "annotations_filtered.txt"), "a") as f:
        f.write("4.png,ab\n")
    host = TrainingDataLoader(cfg, boxes, None, seed=7)
    devl = TrainingDataLoader(cfg, boxes, None, seed=7, device_transform=True)
    with emulated_kernels():
        got = list(devl.load_dataset(batch_size=5, repeat=False))
    ref = list(host.load_dataset(batch_size=5, repeat=False))
    assert len(got) == len(ref) == 1
    for a, b in zip(got[0][1:], ref[0][1:]):
        assert torch.equal(a, b)
    g, r = got[0][0], ref[0][0]
    assert g.shape == r.shape and g.dtype == torch.float32
    diff = (g - r).abs()
    assert float(diff.max()) <= 1.0 / 127.5 + 1e-6
    assert float((diff > 1e-6).float().mean()) < 0.35            # the two roundings agree on most pixels
    assert torch.equal(g == 0, r == 0) or float(g[r == 0].abs().max()) <= 1.0 / 127.5 + 1e-6
    # the 2:1 sample and the zero pad are exact
    lens = (ref[0][2] > 0).sum(1)
    exact = 0
    for i in range(5):
        wpx = int(cfg.char_width * int(lens[i]))
        assert float(g[i, :, :, wpx:].abs().max()) == 0.0
        exact += int(torch.equal(g[i], r[i]))
    assert exact >= 1                                            # at least the 2:1 sample
