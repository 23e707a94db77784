"""Host-side logic of the product against the oracle, with the tensor-core entry points replaced
by their CPU emulation (tests/emu.py): geometry algebra, weight re-layouts, autograd wiring of all
orders (R1 / path-length double backward), the three optimiser updates and the EMA."""
import copy

import pytest
import torch

from common import fractional_cfg, perturbed_params, rel_err, small_cfg
from emu import emulated_kernels
from oracle import aster as OA
from oracle import stylegan as OS
from oracle import train_step as OT
from textboxgan_b200.aster_inferer import AsterInferer
from textboxgan_b200.discriminator import Discriminator
from textboxgan_b200.generator import Generator
from textboxgan_b200.optimizers import Adam, update_optimizer_params
from textboxgan_b200.training_step import TrainingStep


def _build(cfg, GP, DP, with_ocr):
    G = Generator(cfg, device="cpu", seed=0)
    G.load_state_dict(GP)
    D = Discriminator(cfg, device="cpu", seed=0)
    D.load_state_dict(DP)
    aster = AsterInferer(cfg, device="cpu", synthetic_weights=True) if with_ocr else None
    g_opt = update_optimizer_params(cfg.g_opt)
    d_opt = update_optimizer_params(cfg.d_opt)
    mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
    pl_mean = torch.zeros(())
    ts = TrainingStep(G, D, aster, mk(g_opt), mk(g_opt), mk(d_opt), cfg.g_opt["reg_interval"],
                      cfg.d_opt["reg_interval"], pl_mean, cfg)
    return G, D, ts, pl_mean


@pytest.mark.parametrize("do_r1,do_pl,with_ocr,frac", [(False, False, True, False), (True, True, False, False),
                                                        (False, False, True, True), (True, True, True, False),
                                                        (False, True, True, False)])
def test_train_step_matches_oracle(do_r1, do_pl, with_ocr, frac):
    cfg = fractional_cfg(4) if frac else small_cfg(4)
    GP, DP, g = perturbed_params(cfg)
    real, words, labels = OT.synthetic_batch(cfg, 4, g)
    draws = OT.make_draws(cfg, 4, g, with_pl=do_pl)
    st = OT.StepState(copy.deepcopy(GP), copy.deepcopy(DP), OA.init_aster_params(), OT.make_adam(cfg.g_opt),
                      OT.make_adam(cfg.g_opt), OT.make_adam(cfg.d_opt), torch.zeros(()))
    ref_out, ref_grads, _ = OT.train_step(st, cfg, real, torch.zeros(()), words, labels, do_r1, do_pl, 1e-4, draws,
                                          fused=False, with_ocr=with_ocr, ret_grads=True)
    with emulated_kernels():
        G, D, ts, pl_mean = _build(cfg, GP, DP, with_ocr)
        d2 = dict(draws)
        d2["keep_grads"] = True
        out = ts.dist_train_step(real, torch.zeros(()), words, labels, do_r1, do_pl, 1e-4, draws=d2)
        # losses
        flat = lambda o: [float(v) for v in (*o[0], *o[1], o[2])]
        for a, b in zip(flat(out), flat(ref_out)):
            assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (flat(out), flat(ref_out))
        # gradients of the three groups (fp32 both sides; lrelu sign flips at |pre|~1e-7 make a
        # few elements differ, hence the 1e-2 bound relative to each tensor's max)
        g_grads, o_grads, d_grads = ts.last_grads
        for names, got, ref in ((ts._g_names, g_grads, ref_grads[0]), (ts._ocr_names, o_grads, ref_grads[1]),
                                (ts._d_names, d_grads, ref_grads[2])):
            if got is None:
                continue
            for n, a in zip(names, got):
                if n not in ref:
                    assert a is None or float(a.abs().max()) == 0.0
                    continue
                # scalar gradients (noise strengths) are fp32 sums over every pixel with heavy cancellation
                tol = 5e-2 if a.numel() == 1 else 1e-2
                assert rel_err(a, ref[n]) < tol, (n, rel_err(a, ref[n]))
        # updated weights (three Adam updates at pre-update gradients) and state
        # (OCR-pass gradients carry the 1e-4 loss weight: |g| approaches Adam's epsilon, where the update
        # g / (sqrt(v) + eps) is sensitive to fp32 rounding of g — hence 5e-3 of each tensor's maximum)
        for n, p in G.params.items():
            assert rel_err(p, st.G[n]) < 5e-3, n
        for n, p in D.params.items():
            assert rel_err(p, st.D[n]) < 2e-3, n
        if do_pl:
            assert abs(float(pl_mean) - float(st.pl_mean)) < 1e-4 * max(1.0, abs(float(st.pl_mean)))
        assert ts.g_optimizer.iterations.numpy() == 1 and ts.d_optimizer.iterations.numpy() == 1


def test_ema_matches_oracle():
    cfg = small_cfg(4)
    GP, DP, g = perturbed_params(cfg)
    GP2 = {k: v + 0.01 * torch.randn(v.shape, generator=g) for k, v in GP.items()}
    with emulated_kernels():
        G = Generator(cfg, device="cpu", seed=0)
        G.load_state_dict(GP2)
        Gc = Generator(cfg, device="cpu", seed=0)
        Gc.load_state_dict(GP)
        Gc.set_as_moving_average_of(G)
    clone = copy.deepcopy(GP)
    OT.set_as_moving_average_of(clone, GP2)
    for n, p in Gc.params.items():
        assert rel_err(p, clone[n]) < 1e-6, n


def test_adam_is_tf_keras_adam():
    """epsilon sits outside the bias-corrected sqrt (unlike torch.optim.Adam)."""
    with emulated_kernels():
        p = torch.tensor([1.0, -2.0, 3.0, 0.5, 1.5])
        g = torch.tensor([0.1, -0.2, 0.3, 1e-9, 0.0])
        opt = Adam(0.002, beta_1=0.0, beta_2=0.99, epsilon=1e-8)
        v = p.clone().requires_grad_(True)
        opt.apply_gradients([(g, v)])
        t = 1
        m = g
        vv = (1 - 0.99) * g * g
        lr_t = 0.002 * (1 - 0.99 ** t) ** 0.5 / (1 - 0.0 ** t)
        ref = p - lr_t * m / (vv.sqrt() + 1e-8)
        assert torch.allclose(v.detach(), ref, rtol=1e-6, atol=1e-9)
        assert opt.iterations.numpy() == 1


def test_four_consecutive_steps_follow_the_oracle_loss_curve():
    """State carried across iterations (Adam moments and step counters of all three optimisers, pl_mean, w_avg, EMA
    clone) — four steps with the lazy-regularisation schedule of train.py:182-183 (intervals 2 / 3 so that a PL step, an
    R1 step and a second PL step occur), every random draw injected on both sides.  The first step agrees to 5e-4; later
    steps to 2e-2: with beta1 = 0 the Adam update is lr * g / (|g| + eps), i.e. +-lr for every element however small
    its gradient, so fp32 rounding of near-zero gradients (and leaky-ReLU slope choices at |pre| ~ 1e-7) moves a few
    weights by 2*lr per step and the trajectories separate slowly — the reference's own runs differ from each other
    in the same way.  Weights after the fourth step stay within 4 steps * 2 * lr of the oracle's."""
    cfg = small_cfg(4)
    GP, DP, g = perturbed_params(cfg)
    st = OT.StepState(copy.deepcopy(GP), copy.deepcopy(DP), OA.init_aster_params(), OT.make_adam(cfg.g_opt),
                      OT.make_adam(cfg.g_opt), OT.make_adam(cfg.d_opt), torch.zeros(()), g_reg_interval=2,
                      d_reg_interval=3)
    with emulated_kernels():
        G = Generator(cfg, device="cpu", seed=0)
        G.load_state_dict(GP)
        clone = Generator(cfg, device="cpu", seed=0)
        clone.load_state_dict(GP)
        D = Discriminator(cfg, device="cpu", seed=0)
        D.load_state_dict(DP)
        aster = AsterInferer(cfg, device="cpu", synthetic_weights=True)
        g_opt, d_opt = update_optimizer_params(cfg.g_opt), update_optimizer_params(cfg.d_opt)
        mk = lambda o: Adam(o["learning_rate"], beta_1=o["beta1"], beta_2=o["beta2"], epsilon=o["epsilon"])
        pl_mean = torch.zeros(())
        ts = TrainingStep(G, D, aster, mk(g_opt), mk(g_opt), mk(d_opt), 2, 3, pl_mean, cfg)
        for step in range(4):
            do_r1, do_pl = (step + 1) % 3 == 0, (step + 1) % 2 == 0
            real, words, labels = OT.synthetic_batch(cfg, 4, g)
            draws = OT.make_draws(cfg, 4, g, with_pl=do_pl)
            ref = OT.train_step(st, cfg, real, torch.zeros(()), words, labels, do_r1, do_pl, 1e-4, draws, fused=False)
            out = ts.dist_train_step(real, torch.zeros(()), words, labels, do_r1, do_pl, 1e-4, draws=dict(draws))
            clone.set_as_moving_average_of(G)
            flat = lambda o: [float(v) for v in (*o[0], *o[1], o[2])]
            tol = 5e-4 if step == 0 else 2e-2
            for a, b in zip(flat(out), flat(ref)):
                assert abs(a - b) <= tol * max(1.0, abs(b)), (step, flat(out), flat(ref))
            assert (flat(out)[2] > 0) == do_pl and (flat(out)[5] > 0) == do_r1
        bound = 4 * 2 * 0.002 * 1.01        # steps * 2 * lr (synthesis weights get two updates per step: x2 below)
        for n, p in G.params.items():
            assert float((p.detach() - st.G[n]).abs().max()) <= 2 * bound, n
        for n, p in D.params.items():
            assert float((p.detach() - st.D[n]).abs().max()) <= bound, n
        assert abs(float(pl_mean) - float(st.pl_mean)) < 1e-4 * max(1.0, abs(float(st.pl_mean)))
        assert ts.g_optimizer.iterations.numpy() == 4 and ts.ocr_optimizer.iterations.numpy() == 4
        # steps 2-4 prepared their weights through the grouped plan recorded by step 1 (fused.StepWeights)
        assert ts._step_weights.plan is not None and len(ts._step_weights.keys) >= 10
