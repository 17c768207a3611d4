"""Golden learning-rate sequences from the UNMODIFIED reference schedulers (run in the build container only).

    python tests/golden/make_golden_lr.py

The reference ends every training iteration with `self.scheduler.step(...)` (src/defaults/trainer.py:143) on a
`MixedLRScheduler` (src/utils/_utils.py:369-413) that `DefaultWrapper.init_scheduler` (src/defaults/wrappers.py:224-308)
fills from the yml: for the APLA configs `["LinearWarmup", "CosineAnnealingLR"]` (params/finetune/dinov2/NABirds/vit_b/apla.yml:13-17
over __common__.yml:157-161).  This script builds exactly that pair around a torch AdamW, records the learning rate the
optimiser holds at every iteration, and stores the sequences in lr_schedule.json; tests/test_schedule.py pins
apla_b200/schedule.py against them on any machine.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def _accept_verbose():
    """torch >= 2.7 removed the `verbose` argument the reference still passes positionally
    (`super().__init__(optimizer, last_epoch, verbose)`, _utils.py:145): accept and drop it.  Nothing else is touched."""
    base = torch.optim.lr_scheduler.LRScheduler
    if getattr(base, "_apla_verbose_shim", False):
        return
    orig = base.__init__

    def init(self, optimizer, last_epoch=-1, *ignored_verbose):
        orig(self, optimizer, last_epoch)
    base.__init__ = init
    base._apla_verbose_shim = True


def sequence(max_lr, warmup_iters, steps_per_epoch, epochs, cosine_eta_min, types):
    mg.import_reference()
    _accept_verbose()
    from utils._utils import LinearWarmup, MixedLRScheduler
    p = torch.nn.Parameter(torch.zeros(2, 2))
    opt = torch.optim.AdamW([p], lr=max_lr, weight_decay=1e-5)
    scheds, stypes, wi = [None], [None], 0
    for t in types:                                              # same order and arguments as init_scheduler
        if t == "LinearWarmup":
            s = LinearWarmup(opt, max_lr=max_lr, warmup_iters=warmup_iters, warmup_epochs=0, steps_per_epoch=steps_per_epoch)
            wi = s.warmup_iters
        elif t == "CosineAnnealingLR":
            t_max = steps_per_epoch * epochs - (wi if "LinearWarmup" in types else 0)
            s = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=t_max, eta_min=cosine_eta_min)
        scheds.append(s); stypes.append(t)
    mixed = MixedLRScheduler(schedulers=scheds, scheduler_types=stypes, steps_per_epoch=steps_per_epoch)
    lrs = []
    for _ in range(steps_per_epoch * epochs):
        lrs.append(opt.param_groups[0]["lr"])                    # the rate this iteration's optimizer.step() uses
        p.grad = torch.ones_like(p)
        opt.step()
        mixed.step(None, None)
    return lrs


def main():
    cases = {}
    for name, kw in dict(
        nabirds_apla=dict(max_lr=3e-5, warmup_iters=500, steps_per_epoch=200, epochs=6, cosine_eta_min=1e-6,
                          types=["LinearWarmup", "CosineAnnealingLR"]),
        short=dict(max_lr=5e-4, warmup_iters=5, steps_per_epoch=8, epochs=4, cosine_eta_min=1e-6,
                   types=["LinearWarmup", "CosineAnnealingLR"]),
        warmup_only=dict(max_lr=5e-4, warmup_iters=10, steps_per_epoch=8, epochs=3, cosine_eta_min=1e-6,
                         types=["LinearWarmup"]),
        no_warmup_given=dict(max_lr=1e-3, warmup_iters=0, steps_per_epoch=6, epochs=3, cosine_eta_min=1e-6,
                             types=["LinearWarmup", "CosineAnnealingLR"]),
    ).items():
        cases[name] = dict(config=kw, lr=sequence(**kw))
    with open(os.path.join(HERE, "lr_schedule.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cases=cases), f)
    for k, v in cases.items():
        print(k, len(v["lr"]), v["lr"][:3], v["lr"][-2:])


if __name__ == "__main__":
    main()
