"""Row f1: the training loop around the step — schedule, OCR warm-up, LossTracker semantics,
checkpoint keep-N and resume — on the CPU with the kernels emulated (tests/emu.py)."""
import torch

from common import small_cfg
from emu import emulated_kernels
from oracle import tokens as OTK
from textboxgan_b200 import char_tokens as CT
from textboxgan_b200.loss_tracker import LossTracker
from textboxgan_b200.train import Trainer, synthetic_dataset


def test_loss_tracker_skips_non_positive_values_like_the_reference():
    lines = []
    t = LossTracker(["a", "pen"], print_step=2, log_losses=True, printer=lines.append)
    t.increment_losses({"a": torch.tensor(2.0), "pen": torch.tensor(0.0)})     # penalty not computed on this step
    t.increment_losses({"a": torch.tensor(4.0), "pen": torch.tensor(6.0)})
    assert t.losses["a"].result() == 3.0 and t.losses["pen"].result() == 6.0    # utils/loss_tracker.py:41-43
    line = t.print_losses(10)
    assert line.startswith("Step: 10. Avg over the last 2 steps.") and "- a: 3.0000, - pen: 6.0000" in line
    t.reinitialize_tracker()
    assert t.losses["a"].result() == 0.0 and t.timer.count == 0


def test_main_to_aster_ids_matches_the_oracle_table():
    import numpy as np

    ids = np.arange(0, 70).reshape(2, 35)
    assert np.array_equal(CT.main_to_aster_ids(ids), OTK.main_to_aster_ids(ids))


def test_trainer_schedule_checkpoints_and_resume(tmp_path):
    cfg = small_cfg(4)
    cfg.g_opt = dict(cfg.g_opt, reg_interval=2)          # PL on steps 1, 3 (0-based), train.py:182-183
    cfg.d_opt = dict(cfg.d_opt, reg_interval=3)          # R1 on step 2
    cfg.max_steps = 4
    cfg.save_step_frequency = 2
    cfg.num_ckpts_to_keep = 2
    cfg.summary_steps_frequency = {"print_steps": [2], "log_losses": [True]}
    lines, scalars, calls = [], [], []
    with emulated_kernels():
        tr = Trainer(cfg, device="cpu", train_dataset=synthetic_dataset(cfg, 10, device="cpu"), ckpt_dir=str(tmp_path),
                     scalar_writer=lambda d, s: scalars.append((s, d)), printer=lines.append)
        assert tr.regularisation_flags(0) == (False, False) and tr.regularisation_flags(1) == (False, True)
        assert tr.regularisation_flags(2) == (True, False) and tr.regularisation_flags(5) == (True, True)
        assert tr.ocr_weight(5000) == 1e-8 and tr.ocr_weight(5001) == cfg.ocr_loss_weight      # train.py:185-192
        real_step = tr.training_step.dist_train_step

        def spy(*a):
            calls.append((a[4], a[5], a[6]))
            return real_step(*a)

        tr.training_step.dist_train_step = spy
        done = tr.train()
        assert done == 4 and tr.g_optimizer.iterations.numpy() == 4 and tr.d_optimizer.iterations.numpy() == 4
        assert calls == [(False, False, 1e-8), (False, True, 1e-8), (True, False, 1e-8), (False, True, 1e-8)]
        # EMA clone moved away from its initial copy but not onto the generator
        assert not torch.equal(tr.g_clone.flat, tr.generator.flat)
        # keep-N: saved at steps 2, 4 and the final save (step 4 again) -> two files
        names = sorted(p.name for p in tmp_path.iterdir())
        assert names == ["ckpt-2.pt", "ckpt-4.pt"]
        # printed every 2 steps, scalars logged, penalties averaged over the regularised steps only
        assert sum(l.startswith("Step: ") for l in lines) == 2 and [s for s, _ in scalars] == [2, 4]
        assert scalars[0][1]["pl_penalty"] > 0 and scalars[0][1]["r1_penalty"] == 0.0 and scalars[1][1]["r1_penalty"] > 0
        # resume: a fresh Trainer restores weights, optimiser step counters and pl_mean
        tr2 = Trainer(cfg, device="cpu", train_dataset=synthetic_dataset(cfg, 1, device="cpu"), ckpt_dir=str(tmp_path),
                      printer=lines.append)
        assert tr2.g_optimizer.iterations.numpy() == 4
        assert torch.equal(tr2.generator.flat, tr.generator.flat) and torch.equal(tr2.g_clone.flat, tr.g_clone.flat)
        assert torch.equal(tr2.discriminator.flat, tr.discriminator.flat)
        assert float(tr2.pl_mean) == float(tr.pl_mean) != 0.0


def test_infer_generates_cropped_pngs_and_scores_a_corpus(tmp_path):
    """Row f3 (infer.py:26-134) on the CPU with emulated kernels."""
    cv2 = __import__("pytest").importorskip("cv2")
    from textboxgan_b200.generator import Generator
    from textboxgan_b200.infer import Infer

    cfg = small_cfg(2)
    lines = []
    with emulated_kernels():
        G = Generator(cfg, device="cpu", seed=0)
        inf = Infer(cfg, device="cpu", generator=G, printer=lines.append)
        z = torch.randn(1, cfg.z_dim, generator=torch.Generator().manual_seed(0))
        a = inf.generate(["ab", "hello"], z=z)
        b = inf.generate(["ab", "hello"], z=z, truncation_psi=0.5)
        assert a.shape == (2, cfg.char_height, cfg.image_width, 3) and a.dtype.name == "uint8"
        assert (a != b).any()                                                # truncation moves the style
        paths = inf.genererate_chosen_words(["ab", "hello"], "t", str(tmp_path), do_sentence=False)
        im = cv2.imread(paths[1])
        assert im.shape == (cfg.char_height, cfg.char_width * 5, 3)          # cropped to len(word) characters
        (sp,) = inf.genererate_chosen_words(["ab", "hello"], "t", str(tmp_path), do_sentence=True)
        assert cv2.imread(sp).shape == (cfg.char_height, cfg.char_width * 7, 3)
        w = torch.randn(cfg.style_dim)
        c = inf.generate(["ab"], w_latents=w)
        assert c.shape == (1, cfg.char_height, cfg.image_width, 3)
        (tmp_path / "test_corpus.txt").write_text("one\ntwo\nthree\nfour\n")
        loss = inf.infer_test_set(2, str(tmp_path))
        assert loss > 0 and any("AVERAGE TEST LOSS" in l for l in lines)


def test_tensorboard_writer_logs_scalars_and_config(tmp_path):
    import os

    from textboxgan_b200.tensorboard_writer import TensorboardWriter

    w = TensorboardWriter(str(tmp_path), small_cfg(2))
    t = LossTracker(["a"])
    t.increment_losses({"a": 2.0})
    w.log_scalars(t.losses, 3)
    w.log_scalars({"b": 1.5}, 4)
    w.log_config_file(0)
    files = [f for f in os.listdir(tmp_path) if "tfevents" in f]
    assert files and os.path.getsize(os.path.join(tmp_path, files[0])) > 0
