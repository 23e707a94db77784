"""Configuration surface of the training step — field names follow the reference's global
``cfg`` EasyDict (config/config.py:12-149) so that ``TrainingStep`` / ``ModelLoader`` read the
same attributes.  Unlike the reference, building a config has no side effects (the reference
constructs a ``MirroredStrategy`` at import, config/config.py:140); the strategy object is
attached explicitly with :func:`Config.attach_strategy`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction
from typing import Any, List, Optional, Tuple, Union


@dataclass
class Config:
    # Text boxes specs (config/config.py:40-43)
    char_height: int = 64
    char_width: Union[int, Fraction] = 32   # Fraction: BASELINE shapes whose width is not a multiple of max_char_number
    max_char_number: int = 8
    # Model (config/config.py:45-78)
    embedding_out_dim: int = 32
    word_encoder_dense_dim: int = 256
    generator_resolutions: List[Tuple[int, int]] = field(default_factory=list)
    generator_feat_maps: List[int] = field(default_factory=list)
    discrim_resolutions: List[Tuple[int, int]] = field(default_factory=list)
    discrim_feat_maps: List[int] = field(default_factory=list)
    z_dim: int = 512
    style_dim: int = 512
    n_mapping: int = 5
    # Optimizers (config/config.py:81-94)
    g_opt: dict = field(default_factory=lambda: {
        "learning_rate": 0.002, "beta1": 0.0, "beta2": 0.99, "epsilon": 1e-08, "reg_interval": 8})
    d_opt: dict = field(default_factory=lambda: {
        "learning_rate": 0.002, "beta1": 0.0, "beta2": 0.99, "epsilon": 1e-08, "reg_interval": 16})
    batch_size_per_gpu: int = 4
    # OCR (config/config.py:107-111)
    ocr_loss_weight: float = 0.0001
    ocr_loss_type: str = "softmax_crossentropy"
    aster_image_dims: Tuple[int, int] = (64, 256)
    aster_weights: Optional[str] = None
    # explicit opt-in to seeded synthetic recogniser weights when aster_weights is None (tests, benchmarks); the
    # product entry points refuse to train against a random recogniser otherwise
    aster_synthetic_weights: bool = False
    # Summaries / checkpoints (config/config.py:97-103)
    summary_steps_frequency: dict = field(default_factory=lambda: {"print_steps": [50, 500], "log_losses": [False, True]})
    image_summary_step_frequency: int = 500
    validation_step_frequency: int = 10000
    save_step_frequency: int = 10000
    num_ckpts_to_keep: int = 5
    # Others
    shuffle_seed: int = 4444
    max_steps: int = 130000
    # derived / runtime
    image_width: int = 0
    batch_size: int = 0
    num_replicas: int = 1
    strategy: Any = None
    cpu_only: bool = False
    # B200 build: compute dtype of activations on the device path
    compute_dtype: str = "bf16"

    def __post_init__(self) -> None:
        assert self.ocr_loss_type in ["softmax_crossentropy", "mse"]  # config/config.py:111
        iw = self.char_width * self.max_char_number                 # config/config.py:122
        assert iw == int(iw), "char_width * max_char_number must be an integer image width"
        self.image_width = int(iw)
        if isinstance(self.char_width, Fraction) and self.char_width.denominator == 1:
            self.char_width = int(self.char_width)
        if not self.generator_resolutions:
            g_res, g_fm, d_res, d_fm = derive_ladders(self.char_height, self.image_width)
            self.generator_resolutions, self.generator_feat_maps = g_res, g_fm
            self.discrim_resolutions, self.discrim_feat_maps = d_res, d_fm
        # config/config.py:130-136 — feature maps of the word-encoder output
        r0 = self.generator_resolutions[0]
        self.generator_feat_maps = list(self.generator_feat_maps)
        self.generator_feat_maps[0] = int(self.word_encoder_dense_dim * self.max_char_number / (r0[0] * r0[1]))
        self.batch_size = self.batch_size_per_gpu * self.num_replicas  # config/config.py:141
        # config/config.py:145-149
        assert (tuple(self.generator_resolutions[-1]) == tuple(self.discrim_resolutions[0])
                == (self.char_height, self.image_width)), "ladders must end/start at (char_height, image_width)"

    def attach_strategy(self, strategy: Any) -> "Config":
        self.strategy = strategy
        self.num_replicas = int(strategy.num_replicas_in_sync)
        self.batch_size = self.batch_size_per_gpu * self.num_replicas
        return self

    @property
    def n_style(self) -> int:  # generator.py:16
        return 3 * (len(self.generator_resolutions) - 1)


# The reference hard-codes one ladder (config/config.py:48-74, 256x64).  Other BASELINE.json
# shapes use the head of the generator list / tail of the discriminator list.
_REF_G_FM = [None, 512, 256, 256, 128, 128, 64, 32]
_REF_D_FM_TAIL = [512, 512, 256, 256, 128, 128, 64, 32, 16]  # read right-to-left from (4,4)


def derive_ladders(char_height: int, image_width: int):
    """Resolution / feature-map ladders for a (char_height x image_width) text box with W = 4H."""
    assert image_width == 4 * char_height, "ladders are defined for 4:1 text boxes (base grid 2x8)"
    g_res = [(2, 8)]
    while g_res[-1][0] < char_height:
        g_res.append((g_res[-1][0] * 2, g_res[-1][1] * 2))
    assert g_res[-1] == (char_height, image_width)
    g_fm = _REF_G_FM[: len(g_res)]
    # discriminator: halve both until (8,32), then (8,16),(4,8),(4,4)  (config/config.py:65-73)
    d_res = [(char_height, image_width)]
    while d_res[-1][0] > 8:
        d_res.append((d_res[-1][0] // 2, d_res[-1][1] // 2))
    d_res += [(8, 16), (4, 8), (4, 4)]
    d_fm = list(reversed(_REF_D_FM_TAIL[: len(d_res)]))
    return g_res, g_fm, d_res, d_fm


def reference_default() -> Config:
    """The reference's shipped configuration (config/config.py:40-78)."""
    return Config()


def baseline_config(index: int, n_gpus: int = 1) -> Config:
    """BASELINE.json ``configs[index]`` mapped onto reference fields."""
    table = {
        0: dict(batch=4, mcn=8, z=128, h=16, w=64),
        1: dict(batch=32, mcn=8, z=512, h=32, w=128),
        2: dict(batch=64, mcn=12, z=512, h=64, w=256),
        3: dict(batch=256, mcn=12, z=512, h=64, w=256),
        4: dict(batch=512, mcn=16, z=512, h=128, w=512),
    }[index]
    # mcn=12 @ 256 (configs 2, 3): 256/12 is not an integer.  The reference derives image_width from an integer
    # char_width (config/config.py:122) and tf.repeat needs integer repeats (utils/utils.py:30-36), so that
    # shape is outside its domain; here char_width becomes the exact fraction W/mcn and column x belongs to
    # character floor(x*mcn/W) (identical to x // char_width whenever char_width is an integer).
    per_gpu = table["batch"] // n_gpus
    return Config(char_height=table["h"], char_width=Fraction(table["w"], table["mcn"]), max_char_number=table["mcn"],
                  z_dim=table["z"], style_dim=table["z"], batch_size_per_gpu=per_gpu, num_replicas=n_gpus)


cfg = reference_default()
