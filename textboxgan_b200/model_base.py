"""Minimal Keras-like variable container (names, trainable flag, get/set_weights) used by the
Generator / Discriminator mirrors.  The reference relies on ``tf.keras.Model`` for exactly these
services (``trainable_variables`` training_step.py:233, ``weights`` generator.py:52,
``get_weights/set_weights`` models/model_loader.py:19)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import torch


def fresh_seed() -> int:
    """Initialisation seed for a model built without an explicit one: drawn from torch's global CPU generator, so a
    user's ``torch.manual_seed`` governs it and is not discarded (``torch.seed()`` would reseed the global RNG)."""
    return int(torch.randint(0, 2 ** 31 - 1, (), dtype=torch.int64))


class Model:
    ALIGN = 64   # floats; start alignment of every variable inside the flat buffer

    def __init__(self, name: str = ""):
        self.name = name
        self.params: Dict[str, torch.Tensor] = {}
        self._non_trainable: set = set()

    # -- construction ------------------------------------------------------------------------
    def add_weight(self, name: str, value: torch.Tensor, trainable: bool = True) -> torch.Tensor:
        t = value.detach().clone().float()
        t.requires_grad_(trainable)
        self.params[name] = t
        if not trainable:
            self._non_trainable.add(name)
        return t

    # -- Keras surface -----------------------------------------------------------------------
    @property
    def weights(self) -> List[torch.Tensor]:
        return list(self.params.values())

    @property
    def weight_names(self) -> List[str]:
        return list(self.params.keys())

    def trainable_names(self, scopes: Sequence[str] = ("",)) -> List[str]:
        return [n for n in self.params if n not in self._non_trainable and any(n.startswith(s) for s in scopes)]

    @property
    def trainable_variables(self) -> List[torch.Tensor]:
        return [self.params[n] for n in self.trainable_names()]

    def get_weights(self) -> List[torch.Tensor]:
        return [p.detach().clone() for p in self.params.values()]

    def set_weights(self, values: Iterable[torch.Tensor]) -> None:
        values = list(values)
        assert len(values) == len(self.params)
        with torch.no_grad():
            for p, v in zip(self.params.values(), values):
                assert p.shape == v.shape
                p.copy_(v)

    # -- flat-dict interchange (tests / checkpoints) --------------------------------------------
    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach().clone() for k, v in self.params.items()}

    def load_state_dict(self, state: Dict[str, torch.Tensor]) -> None:
        missing = [k for k in self.params if k not in state]
        if missing:
            raise KeyError(f"missing weights: {missing[:5]}")
        with torch.no_grad():
            for k, p in self.params.items():
                p.copy_(state[k].to(p.device, p.dtype).reshape(p.shape))

    def to(self, device) -> "Model":
        """Move to ``device`` and re-home every variable as a view of ONE flat fp32 buffer
        (``self.flat``), in insertion order.  Optimiser groups that are contiguous name ranges
        (see :meth:`flat_range`) then update with a single kernel and all-reduce with a single
        collective — the B200 replacement for per-variable ResourceApplyAdam + per-variable
        NCCL all-reduce behind ``apply_gradients`` (training_step.py:233-235)."""
        # Non-trainable state (w_avg is assigned during the forward pass, latent_encoder.py:39-45)
        # lives outside the flat buffer: views share their base's autograd version counter, so an
        # in-place state update would invalidate every saved parameter of the step.
        # every variable starts on a 256-byte boundary (ALIGN floats): vectorised kernel paths and the library
        # GEMMs' aligned kernels apply to each view; the padding stays zero forever (zero gradient => Adam no-op)
        A = self.ALIGN
        total = sum((p.numel() + A - 1) // A * A for k, p in self.params.items() if k not in self._non_trainable)
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        self.segments = {}
        for k, p in list(self.params.items()):
            if k in self._non_trainable:
                self.params[k] = p.detach().to(device).clone()
                continue
            n = p.numel()
            flat[off: off + n].copy_(p.detach().reshape(-1))
            q = flat[off: off + n].view(p.shape)
            q.requires_grad_(True)
            self.params[k] = q
            self.segments[k] = (off, n)
            off += (n + A - 1) // A * A
        self.flat = flat
        return self

    def broadcast_from(self, src: int = 0) -> None:
        """Make every replica start from rank ``src``'s variables (trainable flat buffer + non-trainable state).  The
        reference's MirroredStrategy mirrors identical variables on all replicas (config/config.py:140); with one
        process per GPU each rank would otherwise initialise from its own RNG."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        with torch.no_grad():
            dist.broadcast(self.flat, src)
            for k in sorted(self._non_trainable):
                dist.broadcast(self.params[k], src)

    def flat_range(self, names: Sequence[str]):
        """(start, end) of the flat buffer covered by ``names``; they must be contiguous."""
        A = self.ALIGN
        segs = sorted(self.segments[n] for n in names)
        for (o0, n0), (o1, _) in zip(segs[:-1], segs[1:]):
            if (o0 + n0 + A - 1) // A * A != o1:
                raise ValueError("variables are not contiguous in the flat buffer")
        return segs[0][0], (segs[-1][0] + segs[-1][1] + A - 1) // A * A


class Submodel:
    """A named view (prefix scope) over a parent's variables — what ``generator.synthesis`` /
    ``generator.word_encoder`` / ``generator.latent_encoder`` are to the training step
    (training_step.py:196,203)."""

    def __init__(self, parent: Model, scope: str):
        self.parent = parent
        self.scope = scope

    @property
    def trainable_variables(self) -> List[torch.Tensor]:
        return [self.parent.params[n] for n in self.parent.trainable_names((self.scope,))]

    @property
    def trainable_names(self) -> List[str]:
        return self.parent.trainable_names((self.scope,))
