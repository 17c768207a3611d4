"""The learning-rate schedule of the reference's training step, as a closed form of the iteration number.

Every `Trainer.global_step` ends with `self.scheduler.step(...)` (src/defaults/trainer.py:143) on a `MixedLRScheduler`
(src/utils/_utils.py:369-413); the APLA configs fill it with `["LinearWarmup", "CosineAnnealingLR"]`
(params/finetune/dinov2/NABirds/vit_b/apla.yml:13-17; `DefaultWrapper.init_scheduler`, src/defaults/wrappers.py:224-308).
The step engine takes its rate as a number (`FineTuneEngine.lr`, a device-side scalar the captured graph reads), so the
schedule is restated here as `lr_at(k)` = the rate the reference's optimiser holds when iteration k (0-based) calls
`optimizer.step()`.  The restatement keeps what the composition of the two schedulers actually does, including:

  * `LinearWarmup` starts from `eta_min = 1e-8` and takes its first increment inside its constructor
    (`_LRScheduler.__init__` steps once), so iteration 0 already runs at `eta_min + (max_lr - eta_min) / W` -- 7e-8 for the
    NABirds config -- and the ramp ends one increment ABOVE `max_lr` (`last_epoch <= warmup_iters` is inclusive,
    _utils.py:152-158);
  * `warmup_iters = 0` is replaced by 1 (_utils.py:131-133);
  * the cosine part only starts stepping after the warm-up (`self.iter > self.warmup_iters`, _utils.py:409-411), anneals
    from wherever the ramp ended down to its own `eta_min` (1e-6 in the yml) over `T_max = total - W` iterations
    (wrappers.py:284-289), by torch's chainable formula, whose product telescopes to the half-cosine used below.

Pinned against sequences recorded from the unmodified classes: tests/golden/make_golden_lr.py, tests/test_schedule.py.
"""
from __future__ import annotations

import math
from typing import Iterator


class WarmupCosineSchedule:
    def __init__(self, max_lr: float, warmup_iters: int, total_iters: int, cosine_eta_min: float = 1e-6,
                 warmup_eta_min: float = 1e-8, cosine: bool = True):
        self.max_lr = float(max_lr)
        self.warmup_iters = int(warmup_iters) if warmup_iters else 1
        self.total_iters = int(total_iters)
        self.cosine_eta_min = float(cosine_eta_min)
        self.warmup_eta_min = float(warmup_eta_min)
        self.cosine = bool(cosine)
        if self.cosine and self.total_iters <= self.warmup_iters:
            raise ValueError("total_iters must exceed warmup_iters (CosineAnnealingLR needs T_max > 0)")

    @classmethod
    def from_reference_config(cls, optimization_params: dict, steps_per_epoch: int, epochs: int) -> "WarmupCosineSchedule":
        """From the `optimization_params.default` mapping of a reference yml (optimizer.params.lr, scheduler.type,
        scheduler.params.{LinearWarmup, CosineAnnealingLR})."""
        sch = optimization_params["scheduler"]
        types = sch["type"] if isinstance(sch["type"], (list, tuple)) else [sch["type"]]
        unknown = [t for t in types if t not in ("LinearWarmup", "CosineAnnealingLR", None)]
        if unknown or "LinearWarmup" not in types:
            raise NotImplementedError(f"only the APLA configs' LinearWarmup (+ CosineAnnealingLR) pair is restated, got {types}")
        wp = sch["params"].get("LinearWarmup", {})
        warmup = steps_per_epoch * wp["warmup_epochs"] if wp.get("warmup_epochs") else wp.get("warmup_iters", 0)
        return cls(optimization_params["optimizer"]["params"]["lr"], warmup, steps_per_epoch * epochs,
                   cosine_eta_min=sch["params"].get("CosineAnnealingLR", {}).get("eta_min", 0.0),
                   cosine="CosineAnnealingLR" in types)

    def lr_at(self, iteration: int) -> float:
        """Rate used by `optimizer.step()` of training iteration `iteration` (0-based)."""
        W, lo = self.warmup_iters, self.warmup_eta_min
        inc = (self.max_lr - lo) / W
        if iteration <= W:
            return lo + (iteration + 1) * inc
        top = lo + (W + 1) * inc
        if not self.cosine:
            return top
        t, T = iteration - W, self.total_iters - W
        return self.cosine_eta_min + (top - self.cosine_eta_min) * (1.0 + math.cos(math.pi * t / T)) / 2.0

    def __iter__(self) -> Iterator[float]:
        return (self.lr_at(k) for k in range(self.total_iters))

    def apply(self, engine, iteration: int) -> float:
        """Set `engine.lr` for the step about to run (the captured graph reads it from its device-side scalar)."""
        engine.lr = self.lr_at(iteration)
        return engine.lr


class CosineSchedule:
    """The per-iteration schedule of the self-supervised (DINOv2) step: `freeze_iters` zeros, a linear ramp of
    `warmup_iters` values from `start_warmup_value` to `base_value` INCLUSIVE of both ends (numpy linspace), then a
    half cosine from `base_value` to `final_value` over the remaining iterations; reads past the end return
    `final_value`.  Restates `CosineScheduler` (src/self_supervised/dinov2/dinov2_utils.py:143-166) without numpy arrays."""

    def __init__(self, base_value: float, final_value: float, total_iters: int, warmup_iters: int = 0,
                 start_warmup_value: float = 0.0, freeze_iters: int = 0):
        self.base_value, self.final_value = float(base_value), float(final_value)
        self.total_iters, self.warmup_iters, self.freeze_iters = int(total_iters), int(warmup_iters), int(freeze_iters)
        self.start_warmup_value = float(start_warmup_value)
        self.zero_until = 0                     # iterations forced to 0 (the frozen last layer, trainer.py:45-47)
        if self.total_iters < self.warmup_iters + self.freeze_iters:
            raise ValueError("total_iters is shorter than freeze + warm-up")

    def __getitem__(self, it: int) -> float:
        if it >= self.total_iters:
            return self.final_value
        if it < self.zero_until or it < self.freeze_iters:
            return 0.0
        k = it - self.freeze_iters
        if k < self.warmup_iters:
            if self.warmup_iters == 1:
                return self.start_warmup_value
            return self.start_warmup_value + (self.base_value - self.start_warmup_value) * k / (self.warmup_iters - 1)
        n = self.total_iters - self.warmup_iters - self.freeze_iters
        j = k - self.warmup_iters
        return self.final_value + 0.5 * (self.base_value - self.final_value) * (1.0 + math.cos(math.pi * j / n))


def build_ssl_schedules(*, lr: float, lr_eta_min: float, lr_warmup_epochs: int, weight_decay: float, momentum_teacher: float,
                        final_momentum_teacher: float, warmup_teacher_temp: float, teacher_temp: float,
                        warmup_teacher_temp_epochs: int, freeze_last_layer_epochs: int, iters_per_epoch: int, epochs: int):
    """-> dict(lr, wd, momentum, teacher_temp, last_layer_lr) of `CosineSchedule`s, composed as `build_schedulers` does
    (src/self_supervised/dinov2/trainer.py:7-56): lr warm-up from 0 then cosine to the yml's CosineAnnealingLR.eta_min;
    weight decay cosine from its base value to the hard-coded 1e-4; teacher momentum cosine to its final value; teacher
    temperature a linear ramp that then stays; the last layer's rate = lr with its first `freeze_last_layer_epochs`
    epochs at 0.  Per step (trainer.py:106-140): `lr[k]`, `wd[k]` go to the optimiser, `teacher_temp[k]` into the
    forward (`SSLMetaArch.forward(..., teacher_temp=)`), `momentum[k]` into `update_teacher`."""
    total = iters_per_epoch * epochs
    lr_kw = dict(base_value=lr, final_value=lr_eta_min, total_iters=total, warmup_iters=lr_warmup_epochs * iters_per_epoch,
                 start_warmup_value=0.0)
    last = CosineSchedule(**lr_kw)
    last.zero_until = freeze_last_layer_epochs * iters_per_epoch
    tt_iters = warmup_teacher_temp_epochs * iters_per_epoch
    return dict(
        lr=CosineSchedule(**lr_kw),
        wd=CosineSchedule(base_value=weight_decay, final_value=1e-4, total_iters=total),
        momentum=CosineSchedule(base_value=momentum_teacher, final_value=final_momentum_teacher, total_iters=total),
        teacher_temp=CosineSchedule(base_value=teacher_temp, final_value=teacher_temp, total_iters=tt_iters,
                                    warmup_iters=tt_iters, start_warmup_value=warmup_teacher_temp),
        last_layer_lr=last)
