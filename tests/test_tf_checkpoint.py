"""Row f4: TensorFlow tensor-bundle / object-graph checkpoints read and written without TensorFlow.  Known answers of the
published formats (CRC-32C check value, LevelDB table magic, varints) + round trips through this module's own writer —
no TensorFlow-written file exists offline (PARITY NOTE in textboxgan_b200/tf_checkpoint.py)."""
import os
import struct

import numpy as np
import pytest
import torch

from common import small_cfg
from textboxgan_b200 import tf_checkpoint as T


def test_crc32c_and_varint_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                     # CRC-32C (Castagnoli) check value
    assert T.crc32c(b"") == 0 and T.crc32c(b"a") == 0xC1D04330
    assert T.crc32c(b"56789", T.crc32c(b"1234")) == 0xE3069283      # incremental == one shot
    blob = os.urandom(100003)                                       # native (libtbg.so) and pure-Python routines agree
    assert T.crc32c(blob) == T.crc32c(blob, pure_python=True) and T.crc32c(b"123456789", pure_python=True) == 0xE3069283
    assert T.mask_crc(0) == 0xA282EAD8                              # LevelDB mask: rotate right 15, add kMaskDelta
    for v in (0, 1, 127, 128, 300, 2 ** 32 + 5):
        assert T._read_varint(T._varint(v), 0) == (v, len(T._varint(v)))
    assert T._varint(300) == b"\xac\x02"


def test_table_round_trip_across_many_blocks(tmp_path):
    items = [(f"key/{i:05d}/with/a/long/shared/prefix".encode(), os.urandom(1 + i % 97)) for i in range(1500)]
    path = str(tmp_path / "t.index")
    T.write_table(path, items, block_size=512)
    data = open(path, "rb").read()
    assert struct.unpack_from("<Q", data, len(data) - 8)[0] == 0xDB4775248B80FB57 and len(data[-48:]) == 48
    assert T.read_table(path) == sorted(items)
    # a flipped byte inside a block is caught by the block checksum
    bad = bytearray(data)
    bad[100] ^= 0xFF
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        T.read_table(path)


def test_tensor_bundle_round_trip_and_checksums(tmp_path):
    g = np.random.default_rng(0)
    tensors = {"a/w" + T.VAR_SUFFIX: g.standard_normal((3, 3, 8, 4)).astype(np.float32),
               "scalar": np.float32(2.5).reshape(()), "ids": np.arange(12, dtype=np.int32).reshape(3, 4),
               "step": np.int64(7).reshape(()), "blob": b"\x00\x01serialized-proto\xff"}
    prefix = str(tmp_path / "ckpt-3")
    T.save_tensor_bundle(prefix, tensors)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    back = T.load_tensor_bundle(prefix + ".index")
    assert set(back) == set(tensors) and back["blob"] == tensors["blob"]
    for k, v in tensors.items():
        if not isinstance(v, bytes):
            assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape and np.array_equal(back[k], v)
    listed = {k: (s, d) for k, s, d in T.list_variables(str(tmp_path))}      # directory -> newest checkpoint
    assert listed["ids"] == ([3, 4], 3) and listed["a/w" + T.VAR_SUFFIX] == ([3, 3, 8, 4], 1)
    # corrupt one tensor byte: the per-tensor crc32c catches it
    shard = prefix + ".data-00000-of-00001"
    raw = bytearray(open(shard, "rb").read())
    raw[-5] ^= 0x10
    open(shard, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        T.load_tensor_bundle(prefix)
    assert T.load_tensor_bundle(prefix, verify_crc=False)["ids"].shape == (3, 4)


def test_object_graph_resolution():
    paths = {("generator", "synthesis", "synth_blocks", "0", "conv_0", "w"): "k0",
             ("generator", "synthesis", "synth_blocks", "1", "conv_0", "w"): "k1",
             ("generator", "latent_encoder", "w_avg"): "k2"}
    g = T.ObjectGraph(T.ObjectGraph.build(paths))
    for p, k in paths.items():
        assert g.resolve(p) == k
    with pytest.raises(KeyError, match="no child 'synth_blocks_x'"):
        g.resolve(("generator", "synthesis", "synth_blocks_x"))
    with pytest.raises(KeyError, match="holds no VARIABLE_VALUE"):
        g.resolve(("generator", "synthesis"))


def test_models_round_trip_through_a_reference_layout_checkpoint(tmp_path):
    """export_tf_checkpoint writes the reference's object layout (generator / g_clone / discriminator roots, attribute
    paths of the reference classes); loading it back restores every variable bit for bit, and the key names are the ones
    tf.train.Checkpoint derives from those attribute paths."""
    from textboxgan_b200.discriminator import Discriminator
    from textboxgan_b200.generator import Generator

    cfg = small_cfg(2)
    G, C, D = Generator(cfg, device="cpu", seed=1), Generator(cfg, device="cpu", seed=2), Discriminator(cfg, device="cpu", seed=3)
    prefix = str(tmp_path / "ckpt-10")
    T.export_tf_checkpoint(prefix, generator=G, g_clone=C, discriminator=D)
    keys = {k for k, _, _ in T.list_variables(prefix)}
    assert "generator/synthesis/synth_blocks/0/conv_0/mod_dense/w/.ATTRIBUTES/VARIABLE_VALUE" in keys
    assert "g_clone/latent_encoder/g_mapping/dense_layers/4/w/.ATTRIBUTES/VARIABLE_VALUE" in keys
    assert "discriminator/last_block/dense_1/w/.ATTRIBUTES/VARIABLE_VALUE" in keys and T.OBJECT_GRAPH_KEY in keys
    assert len(keys) == 1 + 2 * len(G.params) + len(D.params)
    G2, C2, D2 = Generator(cfg, device="cpu", seed=7), Generator(cfg, device="cpu", seed=8), Discriminator(cfg, device="cpu", seed=9)
    T.load_generator_from_tf_checkpoint(G2, str(tmp_path))
    T.load_generator_from_tf_checkpoint(C2, prefix, is_g_clone=True)
    T.load_discriminator_from_tf_checkpoint(D2, prefix + ".index")
    for a, b in ((G, G2), (C, C2), (D, D2)):
        for n in a.params:
            assert torch.equal(a.params[n].detach(), b.params[n].detach()), n
    assert not torch.equal(G2.params["word_encoder/w_embedding"], C2.params["word_encoder/w_embedding"])


def test_model_loader_restores_reference_checkpoints(tmp_path):
    """ModelLoader.load_generator(ckpt_dir=...) accepts a directory holding a TensorFlow checkpoint of the reference
    (models/model_loader.py:22-44: restores {"g_clone": generator} or {"generator": generator})."""
    from textboxgan_b200.generator import Generator
    from textboxgan_b200.model_loader import ModelLoader

    cfg = small_cfg(2)
    G, C = Generator(cfg, device="cpu", seed=1), Generator(cfg, device="cpu", seed=2)
    T.export_tf_checkpoint(str(tmp_path / "ckpt-225000"), generator=G, g_clone=C)
    got = ModelLoader(cfg, device="cpu").load_generator(is_g_clone=True, ckpt_dir=str(tmp_path))
    for n in C.params:
        assert torch.equal(C.params[n].detach(), got.params[n].detach()), n
