"""``tf.keras.optimizers.Adam`` (optimizer_v2, non-amsgrad) as used by train.py:58-75, over the
flat parameter buffers of :class:`textboxgan_b200.model_base.Model`.

``apply_gradients`` mirrors what the reference gets implicitly in replica context
(training_step.py:233-235): cross-replica SUM of the gradients (one all-reduce over the flat
gradient buffer instead of one per variable), then the Adam update (one ``tbg_adam_step`` launch
instead of one ResourceApplyAdam per variable).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import kernels as K


class _Iterations:
    """``optimizer.iterations`` with the ``.numpy()`` accessor train.py:179,211 uses."""

    def __init__(self):
        self.value = 0

    def numpy(self) -> int:
        return self.value

    def __int__(self) -> int:
        return self.value


class Adam:
    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-7):
        self.learning_rate = float(learning_rate)
        self.beta_1 = float(beta_1)
        self.beta_2 = float(beta_2)
        self.epsilon = float(epsilon)
        self.iterations = _Iterations()
        self._lr_dev: Optional[torch.Tensor] = None   # device copy of lr_t (CUDA-graph replay)
        self._pending = None                           # (p, g, m, v, work) between begin_apply and finish_apply
        self.defer_iteration = False                   # True while a captured graph owns the update
        self._slots: Dict[Tuple[int, int, int], Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}

    # -- helpers -------------------------------------------------------------------------------
    @staticmethod
    def _aligned_like(n: int, phase: int, device) -> torch.Tensor:
        """fp32 buffer of ``n`` elements whose address has the same 16-byte phase as the
        parameter range it shadows, so that the vectorised kernel path applies to all four."""
        buf = torch.zeros(n + 4, dtype=torch.float32, device=device)
        return buf[phase: phase + n]

    def _lr_t(self) -> float:
        t = self.iterations.value + 1
        return self.learning_rate * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)

    def use_device_lr(self, device) -> None:
        """Route lr_t through a device scalar; call :meth:`refresh_device_lr` before each replay."""
        if self._lr_dev is None:
            self._lr_dev = torch.zeros(1, dtype=torch.float32, device=device)
        self.refresh_device_lr()

    def refresh_device_lr(self) -> None:
        if self._lr_dev is not None:
            self._lr_dev.fill_(self._lr_t())

    def _lr_arg(self):
        return self._lr_dev if self._lr_dev is not None else self._lr_t()

    # -- Keras surface -------------------------------------------------------------------------
    def apply_gradients(self, grads_and_vars: Iterable[Tuple[torch.Tensor, torch.Tensor]], model=None,
                        names: Optional[Sequence[str]] = None) -> None:
        """Flat-buffer fast path when ``model``/``names`` identify a contiguous variable range
        (the three groups of training_step.py:194-213 all do); per-variable path otherwise."""
        grads_and_vars = [(g, v) for g, v in grads_and_vars]
        if model is not None and names is not None:
            self.begin_apply(model, list(names), [g for g, _ in grads_and_vars])
            self.finish_apply()
            return
        else:
            for g, v in grads_and_vars:
                if g is None:
                    continue
                self._apply_flat_range(v.detach().reshape(-1), g.detach().reshape(-1).float().contiguous(), id(v))
        if not self.defer_iteration:
            self.iterations.value += 1

    # -- split form of apply_gradients: gather + cross-replica SUM first, Adam later ---------------------------------
    def begin_apply(self, model, names: List[str], grads: List[Optional[torch.Tensor]]) -> None:
        """Gather the group's gradients into its flat buffer and START the cross-replica SUM (asynchronous all-reduce on
        the communication stream).  The caller keeps computing — the training step issues the discriminator group's
        reduce before the two generator backward passes, so the 62 MB exchange overlaps them — and calls
        :meth:`finish_apply` where the reference calls ``apply_gradients`` (training_step.py:233-235)."""
        assert getattr(self, "_pending", None) is None, "begin_apply called twice without finish_apply"
        start, end = model.flat_range(names)
        p = model.flat[start:end]
        key = (id(model), start, end)
        if key not in self._slots:
            phase = start % 4
            self._slots[key] = tuple(self._aligned_like(end - start, phase, p.device) for _ in range(3))
        g, m, v = self._slots[key]
        dsts, srcs = [], []
        for n, gr in zip(names, grads):
            o, cnt = model.segments[n]
            view = g[o - start: o - start + cnt]
            if gr is None:
                # Keras apply_gradients skips variables without a gradient (their m / v slots stay untouched); a zero
                # gradient pushed through Adam would decay v instead.  Every variable of the three groups of
                # training_step.py:194-213 is used on every step, so a missing gradient is a wiring error.
                raise RuntimeError(f"Adam.apply_gradients: variable '{n}' of the flat group has no gradient")
            dsts.append(view)
            srcs.append(gr.detach().reshape(-1))
        if dsts:
            torch._foreach_copy_(dsts, srcs)
        work = None
        if dist.is_initialized() and dist.get_world_size() > 1:
            # SUM: every loss already carries 1/global_batch
            work = dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True)
        self._pending = (p, g, m, v, work)

    def finish_apply(self) -> None:
        p, g, m, v, work = self._pending
        self._pending = None
        if work is not None:
            work.wait()                      # the compute stream waits for the reduce; the host does not block
        K.adam_step(p, g, m, v, self._lr_arg(), self.beta_1, self.beta_2, self.epsilon)
        if not self.defer_iteration:
            self.iterations.value += 1

    def _apply_flat_range(self, p: torch.Tensor, g: torch.Tensor, key_id: int) -> None:
        key = (key_id, 0, p.numel())
        if key not in self._slots:
            self._slots[key] = (None, torch.zeros_like(p), torch.zeros_like(p))
        _, m, v = self._slots[key]
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
        K.adam_step(p, g, m, v, self._lr_arg(), self.beta_1, self.beta_2, self.epsilon)

    # -- checkpointing ---------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return {
            "iterations": self.iterations.value,
            "hyper": (self.learning_rate, self.beta_1, self.beta_2, self.epsilon),
            "slots": {f"{k[1]}:{k[2]}": (m.detach().cpu().clone(), v.detach().cpu().clone())
                      for k, (_, m, v) in self._slots.items()},
        }

    def load_state_dict(self, state: dict, model=None) -> None:
        self.iterations.value = int(state["iterations"])
        self._pending_slots = state.get("slots", {})
        if model is not None:
            for rng, (m, v) in self._pending_slots.items():
                start, end = (int(s) for s in rng.split(":"))
                key = (id(model), start, end)
                p = model.flat[start:end]
                g, mm, vv = (self._aligned_like(end - start, start % 4, p.device) for _ in range(3))
                mm.copy_(m)
                vv.copy_(v)
                self._slots[key] = (g, mm, vv)


def update_optimizer_params(params: dict) -> dict:
    """train.py:110-129 (lazy-regularisation correction of lr and betas)."""
    p = dict(params)
    mb_ratio = p["reg_interval"] / (p["reg_interval"] + 1)
    p["learning_rate"] = p["learning_rate"] * mb_ratio
    p["beta1"] = p["beta1"] ** mb_ratio
    p["beta2"] = p["beta2"] ** mb_ratio
    return p
