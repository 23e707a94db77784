"""The C-ABI library loads on a CPU-only box and exports every symbol include/tbg.h declares
(no compute calls without a GPU)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "tbg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tbg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    from textboxgan_b200 import build, lib

    path = build.build()
    assert path.exists()
    cdll = ctypes.CDLL(str(path))
    declared = _declared_symbols()
    assert "tbg_conv2d_igemm" in declared and "tbg_upfirdn2d" in declared
    for sym in declared:
        assert hasattr(cdll, sym), f"{sym} declared in include/tbg.h but not exported"
    # the Python binding covers exactly the declared symbols
    assert sorted(lib.exported_symbols()) == declared
    handle = lib.load()
    assert handle.tbg_version() >= 1
    assert handle.tbg_last_error() is not None


def test_invalid_arguments_return_status_not_exceptions():
    """Argument validation happens before any CUDA call, so it is checkable without a GPU
    (mirrors the OP_REQUIRES checks of upfirdn_2d.cu:241-256)."""
    from textboxgan_b200 import lib

    h = lib.load()
    st = h.tbg_conv2d_igemm(None, None)
    assert st == -1 and b"null" in h.tbg_last_error()
    a = lib.ConvArgs(x=1 << 20, w=1 << 20, out=1 << 20, B=1, H=4, W=4, Cin=48, Ho=4, Wo=4, n_total=64, cout=64,
                     taps_h=3, taps_w=3, pad_h=1, pad_w=1, stride_h=1, stride_w=1)
    st = h.tbg_conv2d_igemm(ctypes.byref(a), None)
    assert st == -1 and b"Cin" in h.tbg_last_error()
    st = h.tbg_upfirdn2d(1 << 20, 1 << 20, 1 << 20, 0, 1, 4, 4, 1, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, None)
    assert st == -1 and b"upx" in h.tbg_last_error()


def test_every_entry_point_validates_its_arguments_before_touching_the_gpu():
    """Each entry point called through the ctypes binding with its full argument list and invalid contents returns
    TBG_ERR_INVALID_ARG (-1) with a message — exercises the marshalling of every signature on a CPU-only box."""
    from textboxgan_b200 import lib

    h = lib.load()
    P = 1 << 20          # a non-null, 16-byte aligned fake pointer (never dereferenced: validation fails first)
    calls = {
        "tbg_conv2d_wgrad": lambda: h.tbg_conv2d_wgrad(None, None),
        "tbg_adam_step": lambda: h.tbg_adam_step(None, P, P, P, 8, 0.1, None, 0.0, 0.99, 1e-8, None),
        "tbg_ema_step": lambda: h.tbg_ema_step(None, P, 8, 0.99, None),
        "tbg_lstm_seq_fwd": lambda: h.tbg_lstm_seq_fwd(None, P, P, P, P, 2, 1, 4, 256, None),
        "tbg_lstm_seq_bwd": lambda: h.tbg_lstm_seq_bwd(None, P, P, P, P, 2, 1, 4, 256, None),
        "tbg_modulate": lambda: h.tbg_modulate(P, P, P, 1, 4, 12, None),
        "tbg_modulate_bwd": lambda: h.tbg_modulate_bwd(P, P, P, P, P, 1, 4, 12, None),
        "tbg_bias_act_bwd": lambda: h.tbg_bias_act_bwd(P, P, None, None, None, P, None, P, None, 1, 4, 64, 1, 1.0, 0, None),
        "tbg_torgb_fwd": lambda: h.tbg_torgb_fwd(P, P, None, P, 1, 4, 12, None),
        "tbg_torgb_bwd": lambda: h.tbg_torgb_bwd(P, P, P, P, P, 1, 4, 12, None),
        "tbg_fromrgb_fwd": lambda: h.tbg_fromrgb_fwd(P, P, P, P, 1, 4, 12, 1.0, 1.0, None),
        "tbg_fromrgb_bwd": lambda: h.tbg_fromrgb_bwd(P, P, P, P, P, P, None, 1, 4, 64, 1.0, 1.0, None),
        "tbg_set_tuning": lambda: h.tbg_set_tuning(b"no_such_key", 1),
        "tbg_bias_act_fwd": lambda: h.tbg_bias_act_fwd(None, None, None, None, P, 1, 4, 64, 1, 1.0, None),
        "tbg_rowdot": lambda: h.tbg_rowdot(P, P, None, 1, 4, 64, None),
        "tbg_batch_resize_normalize": lambda: h.tbg_batch_resize_normalize(P, P, P, None, P, P, 2, 16, 64, None),
        "tbg_dense_fwd": lambda: h.tbg_dense_fwd(None, P, None, P, 4, 8, 8, 1.0, 1.0, 0, 1.0, None),
        "tbg_dense_bwd": lambda: h.tbg_dense_bwd(P, None, P, P, P, P, P, P, 4, 8, 8, 1.0, 1.0, 0, 1.0, 0, None),
        "tbg_pixel_norm_fwd": lambda: h.tbg_pixel_norm_fwd(None, P, 4, 8, None),
        "tbg_pixel_norm_bwd": lambda: h.tbg_pixel_norm_bwd(None, P, P, 4, 8, None),
        "tbg_word_encoder_fwd": lambda: h.tbg_word_encoder_fwd(P, P, P, None, 0.7, P, P, P, P, P, 2, 8, 32, 256, 2, 8, 100, None),
        "tbg_word_encoder_bwd": lambda: h.tbg_word_encoder_bwd(P, None, 0.0, P, P, P, P, P, P, P, P, 2, 8, 32, 256, 2, 8, 128, None),
        "tbg_minibatch_std_fwd": lambda: h.tbg_minibatch_std_fwd(P, P, P, 6, 1, 4, 16, 512, 576, None),
        "tbg_minibatch_std_bwd": lambda: h.tbg_minibatch_std_bwd(P, P, P, 4, 1, 4, 16, 512, 512, None),
        "tbg_r1_sqnorm": lambda: h.tbg_r1_sqnorm(None, P, 4, 100, None),
        "tbg_r1_sqnorm_bwd": lambda: h.tbg_r1_sqnorm_bwd(P, None, P, 4, 100, None),
        "tbg_torgb_skip_fwd": lambda: h.tbg_torgb_skip_fwd(P, P, None, None, None, P, 1, 4, 4, 96, 0, 0, None),
        "tbg_image_grad_nhwc": lambda: h.tbg_image_grad_nhwc(None, None, P, 1, 4, 4, 0, None),
        "tbg_crop_resize_fwd": lambda: h.tbg_crop_resize_fwd(None, P, P, 1, 4, 4, 8, 8, 8, 1, 1, 1, None),
        "tbg_crop_resize_bwd": lambda: h.tbg_crop_resize_bwd(None, P, P, 1, 4, 4, 8, 8, 8, 1, 1, 1, None),
        "tbg_bias_act_rgb_bwd": lambda: h.tbg_bias_act_rgb_bwd(None, P, None, P, P, P, P, P, P, None, P, 1, 4, 12, 1, 1.0, None),
        "tbg_fir4": lambda: h.tbg_fir4(P, P, 1, 4, 4, 4, 4, 12, -1, -1, 1.0, None, None, None, None, 0, 1.0, None),
        "tbg_fir4_down": lambda: h.tbg_fir4_down(P, P, 1, 4, 4, 2, 2, 8, 3, -1, -1, 1.0, None),
        "tbg_fir4_down_adjoint": lambda: h.tbg_fir4_down_adjoint(P, None, P, 1, 4, 4, 2, 2, 12, 2, -1, -1, 1.0, None),
        "tbg_wprep": lambda: h.tbg_wprep(P, None, 1.0, 3, 3, 64, 64, 64, 64, P, P, None, None),
        "tbg_wprep_make_job": lambda: h.tbg_wprep_make_job(None, 0, P, None, 1.0, 3, 3, 64, 64, 64, 64, P, P, None),
        "tbg_wprep_group": lambda: h.tbg_wprep_group(None, 1, 1, None),
        "tbg_wfold": lambda: h.tbg_wfold(P, None, None, None, 1.0, 3, 3, 64, 64, 64, 64, P, None, None, 0, 0, None),
        "tbg_wfold_adj": lambda: h.tbg_wfold_adj(None, None, 1.0, 3, 3, 64, 64, 64, P, None, None, 0, 0, 0, None),
        "tbg_demod_coef": lambda: h.tbg_demod_coef(None, P, P, 1, 64, 64, 1e-8, None),
        "tbg_demod_bwd": lambda: h.tbg_demod_bwd(None, P, P, P, P, P, P, P, P, P, P, P, 1, 64, 64, None),
        "tbg_style_dense_fwd": lambda: h.tbg_style_dense_fwd(None, 0, P, 1, 3, 64, 1.0, None),
        "tbg_style_dense_bwd": lambda: h.tbg_style_dense_bwd(None, 0, P, P, 1, 3, 64, 1.0, None),
        "tbg_attn_decoder_fwd": lambda: h.tbg_attn_decoder_fwd(None, P, None, P, P, P, P, P, P, P, 1, 8, 4, None),
        "tbg_attn_decoder_bwd": lambda: h.tbg_attn_decoder_bwd(None, P, None, P, P, P, P, P, P, P, P, 1, 8, 4, None),
    }
    covered = set(calls) | {"tbg_conv2d_igemm", "tbg_upfirdn2d", "tbg_last_error", "tbg_version", "tbg_launch_count",
                            "tbg_reset_launch_count", "tbg_get_tuning", "tbg_crc32c", "tbg_wprep_job_bytes"}
    assert covered == set(lib.exported_symbols()), set(lib.exported_symbols()) ^ covered
    for name, call in calls.items():
        st = call()
        assert st == -1, (name, st)
        assert len(h.tbg_last_error()) > 0, name


def test_tuning_switches_round_trip_and_no_environment_reads():
    """tbg_set_tuning / tbg_get_tuning are the only way to steer kernel selection: the library sources read no
    environment variables."""
    from textboxgan_b200 import lib

    assert lib.get_tuning("no_such_key") == -1
    for key in ("conv_halo", "wgrad_halo", "halo_staged", "wgrad_staged", "lstm_cluster", "halo_cta2"):
        old = lib.get_tuning(key)
        assert old in (0, 1)
        lib.set_tuning(key, 1 - old)
        assert lib.get_tuning(key) == 1 - old
        lib.set_tuning(key, old)
    lib.set_tuning("halo_b_stages", 6)
    assert lib.get_tuning("halo_b_stages") == 6
    lib.set_tuning("halo_b_stages", 4)
    for f in (ROOT / "textboxgan_b200" / "csrc").glob("*.cu*"):
        assert "getenv" not in f.read_text(), f


def test_product_does_not_import_the_oracle():
    for f in (ROOT / "textboxgan_b200").glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
