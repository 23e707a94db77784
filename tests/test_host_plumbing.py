"""Host-side plumbing that no kernel test covers: flat-buffer layout of the variables (256-byte alignment,
contiguous optimiser groups), the multi-tensor gradient gather of the Adam path, and bench.py's reference arm."""
import json
import os
import subprocess
import sys

import torch

from common import small_cfg
from emu import emulated_kernels
from textboxgan_b200.discriminator import Discriminator
from textboxgan_b200.generator import Generator
from textboxgan_b200.model_base import Model
from textboxgan_b200.optimizers import Adam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flat_buffer_alignment_and_optimizer_groups():
    cfg = small_cfg(4)
    G = Generator(cfg, device="cpu", seed=0)
    D = Discriminator(cfg, device="cpu", seed=0)
    for m in (G, D):
        for name, (off, n) in m.segments.items():
            assert off % Model.ALIGN == 0, name                       # every variable starts on a 256-byte boundary
            assert m.params[name].data_ptr() == m.flat.data_ptr() + 4 * off and m.params[name].numel() == n
        # padding between variables is zero and stays out of every view
        covered = torch.zeros_like(m.flat, dtype=torch.bool)
        for off, n in m.segments.values():
            covered[off: off + n] = True
        assert float(m.flat[~covered].abs().sum()) == 0.0
    # the three optimiser groups of training_step.py:196,203,210 are contiguous ranges
    g_names = G.trainable_names(("synthesis/", "latent_encoder/"))
    o_names = G.trainable_names(("word_encoder/", "synthesis/"))
    for m, names in ((G, g_names), (G, o_names), (D, D.trainable_names())):
        start, end = m.flat_range(names)
        assert start % Model.ALIGN == 0 and end % Model.ALIGN == 0
        assert sum(m.segments[n][1] for n in names) <= end - start


def test_flat_adam_matches_per_variable_update_and_ignores_padding():
    torch.manual_seed(0)
    cfg = small_cfg(4)
    D = Discriminator(cfg, device="cpu", seed=0)
    names = D.trainable_names()
    grads = [torch.randn_like(D.params[n]) for n in names]
    before = {n: D.params[n].detach().clone() for n in names}
    opt = Adam(0.002, beta_1=0.0, beta_2=0.99, epsilon=1e-8)
    with emulated_kernels():
        # a variable without a gradient is a wiring error on the flat path (Keras would skip its slots; a zero gradient
        # through Adam would decay v instead): loud, and nothing is updated
        import pytest

        with pytest.raises(RuntimeError, match="has no gradient"):
            opt.apply_gradients(zip(grads[:3] + [None] + grads[4:], [D.params[n] for n in names]), model=D, names=names)
        assert all(torch.equal(D.params[n], before[n]) for n in names) and opt.iterations.numpy() == 0
        opt.apply_gradients(zip(grads, [D.params[n] for n in names]), model=D, names=names)
    lr_t = 0.002 * (1 - 0.99) ** 0.5
    for n, g in zip(names, grads):
        if g is None:
            assert torch.equal(D.params[n], before[n])
            continue
        v = (1 - 0.99) * g * g
        want = before[n] - lr_t * g / (v.sqrt() + 1e-8)
        assert torch.allclose(D.params[n].detach(), want, rtol=1e-5, atol=1e-7), n
    covered = torch.zeros_like(D.flat, dtype=torch.bool)
    for off, cnt in D.segments.values():
        covered[off: off + cnt] = True
    assert float(D.flat[~covered].abs().sum()) == 0.0                   # Adam never moves the padding
    assert opt.iterations.numpy() == 1


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port of the reference's CPU path, no GPU involved)."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "8",
                          "--warmup", "1", "--config", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "train_step_images_per_sec" and line["value"] > 0
    assert line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the line states what actually ran: 8 iterations of the schedule (the 8th is a path-length step) at batch 4
    assert line["steps"] == 8 and line["warmup"] == 1 and line["config"]["sample_batch"] == 4
    assert "7 plain, 1 path-length, 0 path-length+R1" in line["cpu_baseline"]["sample"]
    assert abs(line["ms_per_step"] * 1e-3 * line["value"] - 4.0) < 1e-6


def test_bench_synthetic_inputs_follow_the_loader_contract():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import train_step as OT
    from textboxgan_b200.config import baseline_config

    for idx in (1, 2):
        cfg = baseline_config(idx)
        real, words, labels = bench.synthetic_inputs(cfg, 8, 4444)
        assert real.shape == (8, 3, cfg.char_height, cfg.image_width) and real.dtype == torch.float32 and real.is_contiguous()
        assert words.shape == (8, cfg.max_char_number) and words.dtype == torch.int32 and labels.dtype == torch.int32
        lens = (words != 0).sum(1)
        assert int(lens.min()) >= 1 and int(words.max()) <= 69
        for b in range(8):
            n = int(lens[b])
            assert (words[b, :n] > 0).all() and (words[b, n:] == 0).all() and (labels[b, n:] == 1).all()
            assert (labels[b, :n] >= 2).all() and (labels[b, :n] <= 95).all()
        # zero right of the word, exactly like the training step's own mask (and the oracle's)
        assert torch.equal(real, OT.mask_text_box(real, words, cfg.char_width))
        assert float(real.abs().max()) <= 1.0
    # the measured arm of bench.py does not touch the oracle
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def run_ours"):src.index("def main")]
    assert "from oracle" not in body and "import oracle" not in body        # only the cpu_baseline leg calls it


def test_device_prefetcher_preserves_order_and_ends():
    """DevicePrefetcher (the Trainer's input stage): same batches in the same order, scalars passed through, StopIteration
    at the end; on a CPU device it is a pass-through."""
    import torch
    from textboxgan_b200.prefetch import DevicePrefetcher

    batches = [(torch.full((2, 3), float(i)), torch.zeros(()), torch.arange(4) + i, 0.5) for i in range(5)]
    got = list(DevicePrefetcher(iter(batches), "cpu"))
    assert len(got) == 5
    for a, b in zip(got, batches):
        assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and a[3] == 0.5
    assert list(DevicePrefetcher(iter([]), "cpu")) == []
