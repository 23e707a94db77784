"""The oracle against golden vectors produced by the REFERENCE'S OWN SOURCES
(tests/golden/make_golden.py executes the unmodified files under /root/reference on the CPU through
the TensorFlow-API shim in oracle/tf_shim and commits tests/golden/reference_tiny.npz).

This is what pins the StyleGAN2 / word-encoder / loss / training-step arithmetic of ``oracle/`` to
the reference (the reference ships no golden vectors of its own)."""
import ast
import copy
import os

import numpy as np
import pytest
import torch

from common import rel_err
from oracle import aster as OA
from oracle import stylegan as OS
from oracle import tokens as TK
from oracle import train_step as OT
from textboxgan_b200.config import Config

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "reference_tiny.npz")


@pytest.fixture(scope="module")
def fx():
    z = np.load(FIX, allow_pickle=False)
    tiny = ast.literal_eval(str(z["cfg_json"]))
    cfg = Config(char_height=tiny["char_height"], char_width=tiny["char_width"], max_char_number=tiny["max_char_number"],
                 embedding_out_dim=tiny["embedding_out_dim"], word_encoder_dense_dim=tiny["word_encoder_dense_dim"],
                 generator_resolutions=[tuple(r) for r in tiny["generator_resolutions"]],
                 generator_feat_maps=[16] + tiny["generator_feat_maps"][1:],
                 discrim_resolutions=[tuple(r) for r in tiny["discrim_resolutions"]],
                 discrim_feat_maps=tiny["discrim_feat_maps"], z_dim=tiny["z_dim"], style_dim=tiny["style_dim"],
                 n_mapping=tiny["n_mapping"], batch_size_per_gpu=tiny["batch_size_per_gpu"])
    t = lambda a: torch.from_numpy(np.array(a))
    G = {k[2:]: t(z[k]) for k in z.files if k.startswith("G/")}
    D = {k[2:]: t(z[k]) for k in z.files if k.startswith("D/")}
    G2 = {k[3:]: t(z[k]) for k in z.files if k.startswith("G2/")}
    D2 = {k[3:]: t(z[k]) for k in z.files if k.startswith("D2/")}
    draws = {}
    for k in z.files:
        if k.startswith("draw/") and "noises/" not in k:
            v = z[k]
            draws[k[5:]] = t(v) if v.ndim > 0 else (float(v) if "coin" in k else int(v))
    n_noise = len([k for k in z.files if k.startswith("draw/noises/")])
    draws["noises"] = [t(z[f"draw/noises/{i}"]) for i in range(n_noise)]
    draws["pl_noises"] = [t(z[f"draw/pl_noises/{i}"]) for i in range(n_noise)]
    out = {k[4:]: z[k] for k in z.files if k.startswith("out/")}
    return dict(cfg=cfg, G=G, D=D, G2=G2, D2=D2, draws=draws, out=out, real=t(z["real"]), words=t(z["words"]),
                labels=t(z["labels"]))


def test_tokenisers_match_reference_functions(fx):
    words = [str(w) for w in fx["out"]["tok_words"]]
    assert (TK.string_to_main_int_sequence(words, 8) == fx["out"]["tok_main"]).all()
    assert (TK.string_to_aster_int_sequence(words, 8) == fx["out"]["tok_aster"]).all()


def test_generator_discriminator_forward_match_reference(fx):
    cfg, G, D, d = fx["cfg"], fx["G"], fx["D"], fx["draws"]
    state = {}
    for fused in (True, False):
        fake = OS.generator(fx["words"], d["z"], G, cfg, training=True, draws=d, fused=fused, state_out=state)
        assert rel_err(fake, fx["out"]["fake_train"]) < 2e-5
    assert rel_err(state["w_avg"], fx["out"]["w_avg_after_fwd"]) < 1e-5
    ev, st = OS.generator(fx["words"], d["z"], G, cfg, training=False, draws=d, truncation_psi=0.7, ret_style=True)
    assert rel_err(ev, fx["out"]["fake_eval"]) < 2e-5 and rel_err(st, fx["out"]["style_eval"]) < 1e-5
    masked = OT.mask_text_box(torch.from_numpy(fx["out"]["fake_train"]), fx["words"], cfg.char_width)
    assert torch.equal(masked, torch.from_numpy(fx["out"]["masked"]))
    sf = OS.discriminator(masked, D, cfg)
    sr = OS.discriminator(fx["real"], D, cfg)
    assert rel_err(sf, fx["out"]["scores_fake"]) < 2e-5 and rel_err(sr, fx["out"]["scores_real"]) < 2e-5
    assert abs(float(OT.generator_loss(sf, cfg.batch_size)) - float(fx["out"]["g_loss"])) < 1e-5
    assert abs(float(OT.discriminator_loss(sf, sr, cfg.batch_size)) - float(fx["out"]["d_loss"])) < 1e-5


def test_convert_inputs_and_ocr_loss_match_reference(fx):
    cfg = fx["cfg"]
    masked = torch.from_numpy(fx["out"]["masked"])
    ci = OA.convert_inputs(masked, fx["labels"], 1, cfg)
    assert rel_err(ci, fx["out"]["convert_inputs"]) < 1e-6
    logits = OA.aster_inferer_call(ci, OA.init_aster_params(), cfg)
    assert abs(float(OA.softmax_cross_entropy_loss(logits, fx["labels"], cfg.batch_size)) - float(fx["out"]["ocr_sce"])) < 1e-3


def test_full_training_step_matches_reference(fx):
    """TrainingStep._train_step of the reference (R1 + path length, three Adam updates) vs the
    oracle: seven losses, pl_mean, and every updated variable."""
    cfg, d = fx["cfg"], fx["draws"]
    st = OT.StepState(copy.deepcopy(fx["G"]), copy.deepcopy(fx["D"]), OA.init_aster_params(), OT.make_adam(cfg.g_opt),
                      OT.make_adam(cfg.g_opt), OT.make_adam(cfg.d_opt), torch.zeros(()))
    out = OT.train_step(st, cfg, fx["real"], torch.zeros(()), fx["words"], fx["labels"], True, True, 1e-4, d, fused=False)
    got = [float(v) for v in (*out[0], *out[1], out[2])]
    want = [float(v) for v in fx["out"]["step_losses"]]
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (got, want)
    assert abs(float(st.pl_mean) - float(fx["out"]["pl_mean"])) < 1e-6
    for name, ref in fx["G2"].items():
        # Adam with beta1 = 0 normalises tiny gradients to +-lr: compare through the step size
        assert (st.G[name] - ref).abs().max() < 2e-4, name
    for name, ref in fx["D2"].items():
        assert (st.D[name] - ref).abs().max() < 2e-4, name
