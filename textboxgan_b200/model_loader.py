"""ModelLoader — builds the discriminator, the generator and its EMA clone, and restores / saves
checkpoints (mirror of models/model_loader.py:10-81).

``load_checkpoint`` returns a manager with ``.save(checkpoint_number=)`` and ``.latest_checkpoint``
like ``tf.train.CheckpointManager``; checkpoints are ``torch.save`` files ``ckpt-<n>.pt`` holding the
flat variable buffers, the non-trainable state, the three Adam states and ``pl_mean`` (the objects
of train.py:94-108).  A missing checkpoint is a silent fresh start, as in the reference.
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional, Tuple

import torch

from .config import Config, cfg as default_cfg
from .discriminator import Discriminator
from .generator import Generator
from .model_base import Model
from .optimizers import Adam


class CheckpointManager:
    def __init__(self, objects: Dict[str, object], directory: str, max_to_keep: Optional[int] = None):
        self.objects = objects
        self.directory = directory
        self.max_to_keep = max_to_keep

    def _all(self):
        if not os.path.isdir(self.directory):
            return []
        found = []
        for f in os.listdir(self.directory):
            m = re.fullmatch(r"ckpt-(\d+)\.pt", f)
            if m:
                found.append((int(m.group(1)), os.path.join(self.directory, f)))
        return sorted(found)

    @property
    def latest_checkpoint(self) -> Optional[str]:
        allc = self._all()
        return allc[-1][1] if allc else None

    def save(self, checkpoint_number: int) -> str:
        os.makedirs(self.directory, exist_ok=True)
        state = {}
        for name, obj in self.objects.items():
            if isinstance(obj, Model):
                state[name] = obj.state_dict()
            elif isinstance(obj, Adam):
                state[name] = obj.state_dict()
            elif torch.is_tensor(obj):
                state[name] = obj.detach().cpu().clone()
            else:
                raise TypeError(f"cannot checkpoint {name}: {type(obj)}")
        path = os.path.join(self.directory, f"ckpt-{int(checkpoint_number)}.pt")
        torch.save(state, path)
        if self.max_to_keep:
            for _, old in self._all()[: -self.max_to_keep]:
                os.remove(old)
        return path

    def restore(self, path: Optional[str], expect_partial: bool) -> bool:
        if not path or not os.path.exists(path):
            return False
        state = torch.load(path, map_location="cpu")
        models = {n: o for n, o in self.objects.items() if isinstance(o, Model)}
        for name, obj in self.objects.items():
            if name not in state:
                if expect_partial:
                    continue
                raise KeyError(f"checkpoint {path} has no entry '{name}'")
            if isinstance(obj, Model):
                obj.load_state_dict(state[name])
            elif isinstance(obj, Adam):
                owner = models.get({"g_optimizer": "generator", "ocr_optimizer": "generator",
                                    "d_optimizer": "discriminator"}.get(name, ""), None)
                obj.load_state_dict(state[name], owner)
            elif torch.is_tensor(obj):
                obj.copy_(state[name].to(obj.device))
        return True


class ModelLoader:
    """Loads the different sub models."""

    def __init__(self, cfg: Optional[Config] = None, device="cuda"):
        self.cfg = cfg if cfg is not None else default_cfg
        self.device = device

    def initiate_models(self) -> Tuple[Discriminator, Generator, Generator]:
        """models/model_loader.py:13-20"""
        discriminator = self._load_discriminator()
        generator = self.load_generator(is_g_clone=False, ckpt_dir=None)
        g_clone = self.load_generator(is_g_clone=True, ckpt_dir=None)
        # one process per GPU: every replica starts from rank 0's initialisation (MirroredStrategy mirrors the
        # variables, config/config.py:140); gradients are SUM-all-reduced afterwards, so the replicas stay identical
        discriminator.broadcast_from(0)
        generator.broadcast_from(0)
        g_clone.set_weights(generator.get_weights())          # set initial g_clone weights same as generator
        return discriminator, generator, g_clone

    def load_generator(self, is_g_clone: bool = False, ckpt_dir: Optional[str] = None) -> Generator:
        """models/model_loader.py:22-44 (the dummy batch-1 forward that builds the Keras variables
        is unnecessary here: variables are created in the constructor)."""
        generator = Generator(self.cfg, device=self.device)
        if ckpt_dir is not None:
            ckpt_kwargs = {"g_clone": generator} if is_g_clone else {"generator": generator}
            self.load_checkpoint(ckpt_kwargs=ckpt_kwargs, model_description="Generator", expect_partial=True,
                                 ckpt_dir=ckpt_dir)
        return generator

    def _load_discriminator(self) -> Discriminator:
        return Discriminator(self.cfg, device=self.device)

    def load_checkpoint(self, ckpt_kwargs: dict, model_description: str, expect_partial: bool, ckpt_dir: str,
                        max_to_keep: Optional[int] = None, resume_step: int = -1) -> CheckpointManager:
        """models/model_loader.py:57-81"""
        manager = CheckpointManager(ckpt_kwargs, ckpt_dir, max_to_keep=max_to_keep)
        resume_checkpoint = os.path.join(ckpt_dir, f"ckpt-{resume_step}.pt") if resume_step != -1 \
            else manager.latest_checkpoint
        if manager.restore(resume_checkpoint, expect_partial):
            print("{} restored from {}".format(model_description, resume_checkpoint))
        elif self._restore_tf_checkpoint(ckpt_kwargs, ckpt_dir, resume_step):
            print("{} restored from the TensorFlow checkpoint in {}".format(model_description, ckpt_dir))
        return manager

    @staticmethod
    def _restore_tf_checkpoint(ckpt_kwargs: dict, ckpt_dir: str, resume_step: int) -> bool:
        """A directory written by the reference itself (tf.train.Checkpoint: ``ckpt-N.index`` + data shards, e.g. the
        authors' "trained model", README.md:60-66): models are restored through the TensorBundle reader of
        :mod:`textboxgan_b200.tf_checkpoint`; optimiser slots of a TensorFlow checkpoint are not imported."""
        if not os.path.isdir(ckpt_dir):
            return False
        if resume_step != -1:
            path = os.path.join(ckpt_dir, f"ckpt-{resume_step}")
            if not os.path.exists(path + ".index"):
                return False
        elif any(f.endswith(".index") for f in os.listdir(ckpt_dir)):
            path = ckpt_dir
        else:
            return False
        from . import tf_checkpoint as T

        done = False
        for name, obj in ckpt_kwargs.items():
            if isinstance(obj, Generator) and name in ("generator", "g_clone"):
                T.load_generator_from_tf_checkpoint(obj, path, is_g_clone=(name == "g_clone"))
                done = True
            elif isinstance(obj, Discriminator) and name == "discriminator":
                T.load_discriminator_from_tf_checkpoint(obj, path)
                done = True
        return done
