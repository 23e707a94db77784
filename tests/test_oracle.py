"""Oracle self-checks on the CPU: hand-derivable known answers, the reference's
internal cross-check (upfirdn_2d_ref vs closed form), fused vs non-fused modulated conv, parameter
counts, gradcheck in fp64."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import small_cfg
from oracle import stylegan as S
from oracle import tokens as TK
from oracle import train_step as T
from textboxgan_b200.config import Config, baseline_config, reference_default


# ---------------------------------------------------------------------------------------------
# tokenisation (bit-exact integer path) — SURVEY.md A.1
# ---------------------------------------------------------------------------------------------
def test_main_tokenizer_known_answers():
    idx = TK.MAIN_WORD_INDEX
    assert idx["<OOV>"] == 1 and idx["0"] == 2 and idx["9"] == 11 and idx["a"] == 12 and idx["z"] == 37
    assert idx["A"] == 38 and idx["Z"] == 63 and idx["-"] == 64 and idx["'"] == 65 and idx["."] == 66
    assert idx["!"] == 67 and idx["?"] == 68 and idx[","] == 69 and idx['"'] == 70
    assert len(idx) == 70  # embedding rows (word_encoder.py:17)
    seq = TK.string_to_main_int_sequence(["0", "aZ\"", "a b"], 8)
    assert seq.dtype == np.int32
    assert seq[0].tolist() == [1, 0, 0, 0, 0, 0, 0, 0]
    assert seq[1].tolist() == [11, 62, 69, 0, 0, 0, 0, 0]
    assert seq[2].tolist() == [11, 0, 12, 0, 0, 0, 0, 0]  # OOV (space) shares id 0 with the pad


def test_aster_tokenizer_known_answers():
    idx = TK.ASTER_WORD_INDEX
    assert idx["0"] == 2 and idx["Z"] == 63 and idx["!"] == 64 and idx["~"] == 95 and len(idx) == 95
    seq = TK.string_to_aster_int_sequence(["0~", ""], 4)
    assert seq.tolist() == [[2, 95, 1, 1], [1, 1, 1, 1]]


def test_truncation_keeps_last_chars_and_empty_words():
    seq = TK.string_to_main_int_sequence(["0123456789"], 8)
    assert seq[0].tolist() == [3, 4, 5, 6, 7, 8, 9, 10]  # truncating="pre" drops LEADING chars
    assert TK.string_to_main_int_sequence([""], 3).tolist() == [[0, 0, 0]]


def test_product_tokenizer_is_bit_exact_with_oracle():
    from textboxgan_b200 import utils as PU

    words = ["Hello", "w0rld!", "", "a-b'c.d", "ThisIsAVeryLongWord", "né", '"?,']
    for mcn in (4, 8, 16):
        assert (PU.string_to_main_int_sequence(words, mcn) == TK.string_to_main_int_sequence(words, mcn)).all()
        assert (PU.string_to_aster_int_sequence(words, mcn) == TK.string_to_aster_int_sequence(words, mcn)).all()
    m = TK.string_to_main_int_sequence(words, 8)
    a = TK.string_to_aster_int_sequence(words, 8)
    # main -> aster id map is consistent wherever the char is in both vocabularies
    conv = TK.main_to_aster_ids(m)
    known = m > 0
    assert (conv[known] == a[known]).all()


# ---------------------------------------------------------------------------------------------
# resampling constants — SURVEY.md A.4
# ---------------------------------------------------------------------------------------------
def test_compute_paddings_known_answers():
    k, p0, p1 = S.compute_paddings([1, 3, 3, 1], True, False, is_conv=True)      # modconv up
    assert (p0, p1) == (1, 1) and abs(k.sum() - 4.0) < 1e-6
    k, p0, p1 = S.compute_paddings([1, 3, 3, 1], True, False, is_conv=False)     # RGB skip
    assert (p0, p1) == (2, 1) and abs(k.sum() - 4.0) < 1e-6
    k, p0, p1 = S.compute_paddings([1, 3, 3, 1], False, True, is_conv=True, convW=3)   # D conv_1
    assert (p0, p1) == (2, 3) and abs(k.sum() - 1.0) < 1e-6
    k, p0, p1 = S.compute_paddings([1, 3, 3, 1], False, True, is_conv=True, convW=1)   # D skip
    assert (p0, p1) == (1, 2)
    assert np.allclose(k, np.outer([1, 3, 3, 1], [1, 3, 3, 1]) / 64.0)


@pytest.mark.parametrize("up,down,pad", [(1, 1, (1, 1)), (2, 1, (2, 1)), (1, 2, (0, 0)), (1, 1, (2, 3)), (2, 2, (1, 0))])
def test_upfirdn_ref_matches_closed_form(up, down, pad):
    """The literal restatement of upfirdn_2d_ref agrees with a direct sum over taps."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 7, 2, generator=g, dtype=torch.float64)
    k = torch.randn(4, 4, generator=g, dtype=torch.float64).numpy()
    y = S.upfirdn_2d_ref(x, k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    inH, inW = 5, 7
    outH = (inH * up + pad[0] + pad[1] - 4) // down + 1
    outW = (inW * up + pad[0] + pad[1] - 4) // down + 1
    assert y.shape == (3, outH, outW, 2)
    z = torch.zeros(3, inH * up + pad[0] + pad[1], inW * up + pad[0] + pad[1], 2, dtype=torch.float64)
    z[:, pad[0]: pad[0] + inH * up: up, pad[0]: pad[0] + inW * up: up, :] = x
    kf = torch.from_numpy(np.ascontiguousarray(k[::-1, ::-1]))
    ref = torch.zeros(3, outH, outW, 2, dtype=torch.float64)
    for oy in range(outH):
        for ox in range(outW):
            patch = z[:, oy * down: oy * down + 4, ox * down: ox * down + 4, :]
            ref[:, oy, ox, :] = (patch * kf[None, :, :, None]).sum(dim=(1, 2))
    assert torch.allclose(y, ref, atol=1e-12)


# ---------------------------------------------------------------------------------------------
# configs / ladders / parameter counts — SURVEY.md A.10, Appendix B
# ---------------------------------------------------------------------------------------------
def test_reference_default_config_and_param_counts():
    cfg = reference_default()
    assert cfg.image_width == 256 and cfg.generator_feat_maps == [128, 512, 256, 256, 128, 128]
    assert cfg.discrim_feat_maps == [64, 128, 128, 256, 256, 512, 512]
    assert cfg.discrim_resolutions == [(64, 256), (32, 128), (16, 64), (8, 32), (8, 16), (4, 8), (4, 4)]
    assert cfg.n_style == 15
    g = torch.Generator().manual_seed(0)
    GP = S.init_generator_params(cfg, g)
    DP = S.init_discriminator_params(cfg, g)
    n_g = sum(v.numel() for k, v in GP.items())
    we = sum(v.numel() for k, v in GP.items() if k.startswith("word_encoder/"))
    mp = sum(v.numel() for k, v in GP.items() if k.startswith("latent_encoder/g_mapping"))
    sy = sum(v.numel() for k, v in GP.items() if k.startswith("synthesis/"))
    assert we - 32 == 10656  # 2208 embedding + 8192 + 256 trainable (+ the frozen 32-wide zero row)
    assert mp == 1313280
    assert sy == 8677916
    assert n_g - GP["latent_encoder/w_avg"].numel() - GP["word_encoder/w0_embedding"].numel() == 10001852
    assert sum(v.numel() for v in DP.values()) == 15594817


def test_baseline_ladders():
    c0 = baseline_config(0)
    assert c0.generator_feat_maps == [128, 512, 256, 256] and c0.discrim_feat_maps == [128, 256, 256, 512, 512]
    c1 = baseline_config(1)
    assert c1.generator_resolutions[-1] == (32, 128) and c1.n_style == 12 and c1.char_width == 16
    assert c1.discrim_feat_maps == [128, 128, 256, 256, 512, 512]
    c4 = baseline_config(4, n_gpus=8)
    assert c4.batch_size_per_gpu == 64 and c4.generator_feat_maps[0] == 256 and c4.n_style == 18


# ---------------------------------------------------------------------------------------------
# modulated conv: the reference's two formulations agree (modulated_conv2d.py:85-93 vs :95-96)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("up", [False, True])
def test_fused_and_nonfused_modconv_agree(up):
    g = torch.Generator().manual_seed(0)
    P = {"c/w": torch.randn(3, 3, 6, 5, generator=g, dtype=torch.float64),
         "c/mod_dense/w": torch.randn(7, 6, generator=g, dtype=torch.float64),
         "c/mod_bias/b": torch.randn(6, generator=g, dtype=torch.float64)}
    x = torch.randn(3, 6, 4, 8, generator=g, dtype=torch.float64)
    y = torch.randn(3, 7, generator=g, dtype=torch.float64)
    a = S.modulated_conv2d(x, y, P, "c", up=up, demodulate=True, fused=True, in_h_res=4, in_w_res=8)
    b = S.modulated_conv2d(x, y, P, "c", up=up, demodulate=True, fused=False, in_h_res=4, in_w_res=8)
    assert a.shape == (3, 5, 8 if up else 4, 16 if up else 8)
    assert torch.allclose(a, b, atol=1e-10)


def test_mask_text_box_and_zero_init_invariants():
    cfg = small_cfg(4)
    g = torch.Generator().manual_seed(0)
    GP = S.init_generator_params(cfg, g)
    _, words, _ = T.synthetic_batch(cfg, 4, g)
    d1 = T.make_draws(cfg, 4, g)
    d2 = dict(d1)
    d2["noises"] = [torch.randn_like(n) for n in d1["noises"]]
    a = S.generator(words, d1["z"], GP, cfg, training=True, draws=d1)
    b = S.generator(words, d1["z"], GP, cfg, training=True, draws=d2)
    assert torch.equal(a, b)  # noise strengths initialise to 0 (noise.py:9-10)
    m = T.mask_text_box(a, words, cfg.char_width)
    for bi in range(4):
        n = int((words[bi] != 0).sum())
        assert float(m[bi, :, :, n * cfg.char_width:].abs().max()) == 0.0 if n < cfg.max_char_number else True
        assert torch.equal(m[bi, :, :, : n * cfg.char_width], a[bi, :, :, : n * cfg.char_width])


def test_minibatch_std_grouping():
    x = torch.randn(8, 3, 2, 2, dtype=torch.float64)
    y = S.minibatch_std(x, 4, 1)
    assert y.shape == (8, 4, 2, 2)
    # sample n belongs to group {m, m + B/G, ...} with m = n mod (B/G)
    for m in range(2):
        grp = x[m::2]
        ref = torch.sqrt(grp.var(dim=0, unbiased=False) + 1e-8).mean()
        assert torch.allclose(y[m::2, 3], ref.expand(4, 2, 2))


def test_gradcheck_modconv_fp64():
    g = torch.Generator().manual_seed(0)
    P = {"c/w": torch.randn(3, 3, 3, 2, generator=g, dtype=torch.float64),
         "c/mod_dense/w": torch.randn(4, 3, generator=g, dtype=torch.float64),
         "c/mod_bias/b": torch.randn(3, generator=g, dtype=torch.float64)}
    x = torch.randn(2, 3, 2, 4, generator=g, dtype=torch.float64, requires_grad=True)
    y = torch.randn(2, 4, generator=g, dtype=torch.float64, requires_grad=True)
    w = P["c/w"].clone().requires_grad_(True)

    def f(x_, y_, w_):
        return S.modulated_conv2d(x_, y_, {**P, "c/w": w_}, "c", up=True, demodulate=True, fused=False,
                                  in_h_res=2, in_w_res=4)

    assert torch.autograd.gradcheck(f, (x, y, w), eps=1e-6, atol=1e-5)
    assert torch.autograd.gradgradcheck(f, (x, y, w), eps=1e-6, atol=1e-4)


def test_adam_state_tf_semantics():
    P = {"a": torch.tensor([1.0, 2.0])}
    st = T.AdamState(lr=0.1, beta1=0.5, beta2=0.9, eps=1e-8)
    g = torch.tensor([0.3, -0.4])
    st.apply(P, {"a": g})
    m = 0.5 * g
    v = 0.1 * g * g
    lr_t = 0.1 * math.sqrt(1 - 0.9) / (1 - 0.5)
    assert torch.allclose(P["a"], torch.tensor([1.0, 2.0]) - lr_t * m / (v.sqrt() + 1e-8))
    p = T.update_optimizer_params({"learning_rate": 0.002, "beta1": 0.0, "beta2": 0.99, "epsilon": 1e-8,
                                   "reg_interval": 8})
    assert abs(p["learning_rate"] - 0.002 * 8 / 9) < 1e-12 and p["beta1"] == 0.0
    assert abs(p["beta2"] - 0.99 ** (8 / 9)) < 1e-12


def test_fractional_char_width_extension_reduces_to_reference_for_integer_widths():
    """mask / crop for char_width = W/mcn (BASELINE configs[2], [3]): floor(x / char_width) in exact integer
    arithmetic; for integer widths it is the reference's repeat (utils/utils.py:30-36) and crop
    (aster_inferer.py:182) bit for bit."""
    from fractions import Fraction

    from textboxgan_b200 import utils as U
    from textboxgan_b200.config import baseline_config

    words = torch.tensor([[5, 3, 0, 0, 0, 0, 0, 0], [1, 2, 3, 4, 5, 6, 7, 8], [9, 0, 0, 0, 0, 0, 0, 0]], dtype=torch.int32)
    img = torch.rand(3, 3, 4, 128)
    a = T.mask_text_box(img, words, 16)
    b = T.mask_text_box(img, words, Fraction(128, 8))
    c = U.mask_text_box(img, words, Fraction(16))
    assert torch.equal(a, b) and torch.equal(a, c)
    assert torch.equal(U.crop_width(torch.tensor([0, 1, 5, 8]), 16), torch.tensor([0, 16, 80, 128]))
    # 256 columns, 12 characters: boundaries at floor(n * 64/3); every character gets 21 or 22 columns
    idx = U.column_char_index(256, Fraction(256, 12))
    counts = torch.bincount(idx, minlength=12)
    assert idx.min() == 0 and idx.max() == 11 and set(counts.tolist()) == {21, 22} and int(counts.sum()) == 256
    assert torch.equal(U.crop_width(torch.tensor([1, 3, 12]), Fraction(256, 12)), torch.tensor([21, 64, 256]))
    w12 = torch.zeros(2, 12, dtype=torch.int32)
    w12[0, :3] = 7
    w12[1, :] = 7
    im = torch.ones(2, 3, 2, 256)
    m = U.mask_text_box(im, w12, Fraction(256, 12))
    assert torch.equal(m, T.mask_text_box(im, w12, Fraction(256, 12)))
    assert float(m[0, 0, 0].sum()) == 64 and float(m[1, 0, 0].sum()) == 256
    c2 = baseline_config(2)
    assert c2.image_width == 256 and c2.max_char_number == 12 and c2.generator_feat_maps[0] == 192 and c2.n_style == 15
