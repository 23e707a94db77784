"""Generate golden fixtures by executing the REFERENCE'S OWN SOURCES (unmodified, from
/root/reference) on the CPU under the TensorFlow-API shim of oracle/tf_shim.

    python tests/golden/make_golden.py            # writes tests/golden/reference_tiny.npz

The reference cannot travel to the GPU box and TensorFlow is not installable offline, so the
fixture is committed: parameters, inputs, every random draw, and the outputs of the reference code
(generator forward in training and inference mode, discriminator forward, mask_text_box,
convert_inputs, GAN/OCR losses, the full ``TrainingStep._train_step`` with R1 + path-length
regularisation: losses, the three gradient sets' effect = updated variables, pl_mean, w_avg).
A tiny ladder (32x8 image, 16/8 feature maps, z = 16) keeps the fixture < 1 MB; it exercises every
code path of the hot path (up/down convolutions, width-only stride, minibatch-std, style mixing,
dropout, truncation, double backward).  ASTER itself is absent from the reference, so the training
step runs with this repo's oracle ASTER stand-in behind the reference's own ``convert_inputs``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TBG_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)

torch.manual_seed(0)
torch.set_num_threads(4)

# ---- configure the reference's global cfg BEFORE its model modules are imported ----------------
from config import cfg  # noqa: E402

TINY = dict(char_height=8, char_width=4, max_char_number=8, embedding_out_dim=8, word_encoder_dense_dim=32,
            generator_resolutions=[(2, 8), (4, 16), (8, 32)], generator_feat_maps=[None, 16, 8],
            discrim_resolutions=[(8, 32), (8, 16), (4, 8), (4, 4)], discrim_feat_maps=[8, 16, 16, 16],
            z_dim=16, style_dim=16, n_mapping=2, batch_size_per_gpu=4)
for k, v in TINY.items():
    cfg[k] = v
cfg.image_width = cfg.char_width * cfg.max_char_number
cfg.generator_feat_maps[0] = int(cfg.word_encoder_dense_dim * cfg.max_char_number
                                 / (cfg.generator_resolutions[0][0] * cfg.generator_resolutions[0][1]))
cfg.batch_size = cfg.batch_size_per_gpu
assert cfg.cpu_only

from models.custom_stylegan2.discriminator import Discriminator  # noqa: E402
from models.custom_stylegan2.generator import Generator  # noqa: E402
from models.losses.gan_losses import discriminator_loss, generator_loss  # noqa: E402
from models.losses.ocr_losses import softmax_cross_entropy_loss  # noqa: E402
import training_step as ref_ts  # noqa: E402
from aster_ocr_utils.aster_inferer import AsterInferer as RefAster  # noqa: E402
from utils.utils import mask_text_box, string_to_aster_int_sequence, string_to_main_int_sequence  # noqa: E402

from oracle import aster as OA  # noqa: E402
from oracle import train_step as OT  # noqa: E402
from textboxgan_b200.config import Config  # noqa: E402


def plain(t):
    return t.detach().as_subclass(torch.Tensor).clone()


def g_params(G):
    """Reference Generator variables -> oracle flat names."""
    P = {}
    we = G.word_encoder
    P["word_encoder/w_embedding"] = plain(we.w_embedding)
    P["word_encoder/w0_embedding"] = plain(we.w0_embedding)
    P["word_encoder/fc/kernel"] = plain(we.fc.kernel)
    P["word_encoder/fc/bias"] = plain(we.fc.bias)
    le = G.latent_encoder
    for i, (d, b) in enumerate(zip(le.g_mapping.dense_layers, le.g_mapping.bias_act_layers)):
        P[f"latent_encoder/g_mapping/dense_{i}/w"] = plain(d.w)
        P[f"latent_encoder/g_mapping/bias_{i}/b"] = plain(b.b)
    P["latent_encoder/w_avg"] = plain(le.w_avg)
    syn = G.synthesis

    def modconv(prefix, m):
        P[prefix + "/w"] = plain(m.w)
        P[prefix + "/mod_dense/w"] = plain(m.mod_dense.w)
        P[prefix + "/mod_bias/b"] = plain(m.mod_bias.b)

    def torgb(prefix, t):
        modconv(prefix + "/conv", t.conv)
        P[prefix + "/bias/b"] = plain(t.apply_bias.b)

    res = cfg.generator_resolutions
    torgb(f"synthesis/{res[0][0]}x{res[0][1]}/ToRGB", syn.initial_torgb)
    for (h, w), blk, t in zip(res[1:], syn.synth_blocks, syn.torgbs):
        pb = f"synthesis/{h}x{w}/block"
        modconv(pb + "/conv_0", blk.conv_0)
        P[pb + "/noise_0/w"] = plain(blk.apply_noise_0.noise_strength)
        P[pb + "/bias_0/b"] = plain(blk.apply_bias_act_0.b)
        modconv(pb + "/conv_1", blk.conv_1)
        P[pb + "/noise_1/w"] = plain(blk.apply_noise_1.noise_strength)
        P[pb + "/bias_1/b"] = plain(blk.apply_bias_act_1.b)
        torgb(f"synthesis/{h}x{w}/ToRGB", t)
    return P


def d_params(D):
    P = {}
    res = cfg.discrim_resolutions
    r0 = res[0]
    P[f"{r0[0]}x{r0[1]}/FromRGB/conv/w"] = plain(D.initial_fromrgb.conv.w)
    P[f"{r0[0]}x{r0[1]}/FromRGB/bias/b"] = plain(D.initial_fromrgb.apply_bias_act.b)
    for (h, w), blk in zip(res[:-1], D.blocks):
        pb = f"{h}x{w}"
        P[pb + "/conv_0/w"] = plain(blk.conv_0.w)
        P[pb + "/bias_0/b"] = plain(blk.apply_bias_act_0.b)
        P[pb + "/conv_1/w"] = plain(blk.conv_1.w)
        P[pb + "/bias_1/b"] = plain(blk.apply_bias_act_1.b)
        P[pb + "/skip/w"] = plain(blk.conv_skip.w)
    rf = res[-1]
    pl = f"{rf[0]}x{rf[1]}/last"
    lb = D.last_block
    P[pl + "/conv_0/w"] = plain(lb.conv_0.w)
    P[pl + "/bias_0/b"] = plain(lb.apply_bias_act_0.b)
    P[pl + "/dense_1/w"] = plain(lb.dense_1.w)
    P[pl + "/bias_1/b"] = plain(lb.apply_bias_act_1.b)
    P["last_dense/w"] = plain(D.last_dense.w)
    P["last_bias/b"] = plain(D.last_bias.b)
    return P


def perturb(model):
    """Reference initialises biases / noise strengths / w_avg to zero; make them non-trivial."""
    g = torch.Generator().manual_seed(7)
    for v in model.weights:
        if v._tf_name in ("b", "bias", "w_avg") or (v._tf_name == "w" and v.dim() == 0):
            v.assign(torch.randn(v.shape, generator=g) * 0.1 if v.dim() > 0 else torch.tensor(0.2))


def queue_generator_draws(d, training, batch):
    q = []
    if training:
        q.append(("dropout", d["dropout_mask"]))
        q.append(("normal", d["z2"]))
        q.append(("uniform", torch.tensor(d["mix_coin"])))
        if d["mix_coin"] < 0.9:
            q.append(("uniform_int", torch.tensor(d["mix_cutoff"])))
    for n in d["noises"]:
        q.append(("normal", n[:batch]))
    return q


def main():
    B = cfg.batch_size_per_gpu
    ocfg = Config(char_height=cfg.char_height, char_width=cfg.char_width, max_char_number=cfg.max_char_number,
                  embedding_out_dim=cfg.embedding_out_dim, word_encoder_dense_dim=cfg.word_encoder_dense_dim,
                  generator_resolutions=list(cfg.generator_resolutions), generator_feat_maps=list(cfg.generator_feat_maps),
                  discrim_resolutions=list(cfg.discrim_resolutions), discrim_feat_maps=list(cfg.discrim_feat_maps),
                  z_dim=cfg.z_dim, style_dim=cfg.style_dim, n_mapping=cfg.n_mapping, batch_size_per_gpu=B)
    out = {}
    # ---- tokenisers through the reference's own functions ----
    words_txt = ["Hello", "w0rld!", "", "a-b'c.d", "ThisIsAVeryLongWord", "né", '"?,', "0"]
    out["tok_main"] = string_to_main_int_sequence(words_txt).astype(np.int32)
    out["tok_aster"] = string_to_aster_int_sequence(words_txt).astype(np.int32)
    out["tok_words"] = np.array(words_txt)

    # ---- build the reference models (model_loader.py:26-31,48-53 do the same dummy forward) ----
    G = Generator()
    G((tf.ones((1, cfg.max_char_number), dtype=tf.int32), tf.ones((1, cfg.z_dim))), batch_size=1)
    D = Discriminator()
    D(tf.ones((1, 3, cfg.char_height, cfg.image_width)))
    perturb(G)
    perturb(D)
    GP, DP = g_params(G), d_params(D)

    gen = torch.Generator().manual_seed(4444)
    real, words, labels = OT.synthetic_batch(ocfg, B, gen)
    draws = OT.make_draws(ocfg, B, gen, with_pl=True)
    draws["mix_coin"], draws["mix_cutoff"] = 0.3, 4          # exercise the mixing branch

    # ---- forward passes ----
    tf.DRAWS.queue = queue_generator_draws(draws, True, B)
    w_avg0 = plain(G.latent_encoder.w_avg)
    fake_train = G([tf.constant(words), tf.constant(draws["z"])], batch_size=B, training=True)
    assert not tf.DRAWS.queue
    out["w_avg_after_fwd"] = plain(G.latent_encoder.w_avg).numpy()
    G.latent_encoder.w_avg.assign(w_avg0)
    tf.DRAWS.queue = queue_generator_draws(draws, False, B)
    fake_eval, style_eval = G([tf.constant(words), tf.constant(draws["z"])], batch_size=B, ret_style=True,
                              truncation_psi=0.7, training=False)
    assert not tf.DRAWS.queue
    masked = mask_text_box(fake_train, tf.constant(words), cfg.char_width)
    scores_fake = D(masked)
    scores_real = D(tf.constant(real))
    out.update(fake_train=plain(fake_train).numpy(), fake_eval=plain(fake_eval).numpy(),
               style_eval=plain(style_eval).numpy(), masked=plain(masked).numpy(),
               scores_fake=plain(scores_fake).numpy(), scores_real=plain(scores_real).numpy(),
               g_loss=float(generator_loss(scores_fake)), d_loss=float(discriminator_loss(scores_fake, scores_real)))
    conv_in = RefAster.convert_inputs(masked, tf.constant(labels), blank_label=1)
    out["convert_inputs"] = plain(conv_in).numpy()
    AP = OA.init_aster_params()
    logits = OA.aster_inferer_call(plain(conv_in), AP, ocfg)
    out["ocr_sce"] = float(softmax_cross_entropy_loss(logits, tf.constant(labels)))

    # ---- full training step (R1 + PL), reference code, TF-Keras Adam semantics of the shim ----
    class StubAster:
        convert_inputs = staticmethod(RefAster.convert_inputs)

        def __call__(self, x):
            return OA.aster_inferer_call(x, AP, ocfg)

    upd = lambda o: dict(o, learning_rate=o["learning_rate"] * o["reg_interval"] / (o["reg_interval"] + 1),
                         beta1=o["beta1"] ** (o["reg_interval"] / (o["reg_interval"] + 1)),
                         beta2=o["beta2"] ** (o["reg_interval"] / (o["reg_interval"] + 1)))   # train.py:110-129
    go, do = upd(cfg.g_opt), upd(cfg.d_opt)
    mk = lambda o: tf.keras.optimizers.Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
    pl_mean = tf.Variable(0.0, name="pl_mean", trainable=False)
    ts = ref_ts.TrainingStep(G, D, StubAster(), mk(go), mk(go), mk(do), go["reg_interval"], do["reg_interval"], pl_mean)
    q = [("normal", draws["z"])] + queue_generator_draws(draws, True, B)
    pb = max(1, B // 2)
    q += [("normal", draws["pl_z"])] + [("normal", n) for n in draws["pl_noises"]] + [("normal", draws["pl_image_noise"])]
    tf.DRAWS.queue = q
    gen_l, disc_l, ocr_l = ts.dist_train_step(tf.constant(real), tf.constant(0.0), tf.constant(words),
                                              tf.constant(labels), True, True, 1e-4)
    assert not tf.DRAWS.queue, len(tf.DRAWS.queue)
    out["step_losses"] = np.array([float(v) for v in (*gen_l, *disc_l, ocr_l)], dtype=np.float64)
    out["pl_mean"] = float(pl_mean)
    GP2, DP2 = g_params(G), d_params(D)

    np.savez_compressed(
        os.path.join(HERE, "reference_tiny.npz"),
        cfg_json=np.array(repr(TINY)),
        real=real.numpy(), words=words.numpy(), labels=labels.numpy(),
        **{f"draw/{k}": (np.asarray(v) if not isinstance(v, list) else np.array([0])) for k, v in draws.items()
           if not isinstance(v, list)},
        **{f"draw/noises/{i}": n.numpy() for i, n in enumerate(draws["noises"])},
        **{f"draw/pl_noises/{i}": n.numpy() for i, n in enumerate(draws["pl_noises"])},
        **{f"G/{k}": v.numpy() for k, v in GP.items()}, **{f"D/{k}": v.numpy() for k, v in DP.items()},
        **{f"G2/{k}": v.numpy() for k, v in GP2.items()}, **{f"D2/{k}": v.numpy() for k, v in DP2.items()},
        **{f"out/{k}": np.asarray(v) for k, v in out.items()},
    )
    print("wrote", os.path.join(HERE, "reference_tiny.npz"))
    print("losses", out["step_losses"], "pl_mean", out["pl_mean"], "g_loss", out["g_loss"], "d_loss", out["d_loss"])


if __name__ == "__main__":
    main()
