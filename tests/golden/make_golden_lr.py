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


# ---- the self-supervised (DINOv2) step: CosineScheduler (src/self_supervised/dinov2/dinov2_utils.py:143-166) --------------
def ssl_sequences():
    """The five schedules `build_schedulers` creates (src/self_supervised/dinov2/trainer.py:7-56) with the ISIC2019 values
    (params/pretrain/dinov2/ISIC2019/vit_b/__common__.yml:135-140,169-195 + apla.yml:12), at a reduced iteration count."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_dinov2_utils", os.path.join(mg.REF, "self_supervised/dinov2/dinov2_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    CS = mod.CosineScheduler
    ipe, epochs = 7, 12                       # iterations per epoch, epochs (the yml has 300 epochs)
    total = ipe * epochs
    cfg = dict(lr=1e-3, eta_min=1e-6, lr_warmup_epochs=2, weight_decay=1e-5, momentum_teacher=0.994, final_momentum_teacher=1.0,
               warmup_teacher_temp=0.04, teacher_temp=0.07, warmup_teacher_temp_epochs=3, freeze_last_layer_epochs=1,
               iters_per_epoch=ipe, epochs=epochs)
    lr = CS(base_value=cfg["lr"], final_value=cfg["eta_min"], total_iters=total, warmup_iters=cfg["lr_warmup_epochs"] * ipe,
            start_warmup_value=0)
    wd = CS(base_value=cfg["weight_decay"], final_value=1e-4, total_iters=total, warmup_iters=0)
    mom = CS(base_value=cfg["momentum_teacher"], final_value=cfg["final_momentum_teacher"], total_iters=total, warmup_iters=0)
    tt = CS(base_value=cfg["teacher_temp"], final_value=cfg["teacher_temp"], total_iters=cfg["warmup_teacher_temp_epochs"] * ipe,
            warmup_iters=cfg["warmup_teacher_temp_epochs"] * ipe, start_warmup_value=cfg["warmup_teacher_temp"])
    last = CS(base_value=cfg["lr"], final_value=cfg["eta_min"], total_iters=total, warmup_iters=cfg["lr_warmup_epochs"] * ipe,
              start_warmup_value=0)
    last.schedule[: cfg["freeze_last_layer_epochs"] * ipe] = 0
    n = total + 3                                # a few reads past the end (-> final_value)
    return dict(config=cfg, lr=[float(lr[i]) for i in range(n)], wd=[float(wd[i]) for i in range(n)],
                momentum=[float(mom[i]) for i in range(n)], teacher_temp=[float(tt[i]) for i in range(n)],
                last_layer_lr=[float(last[i]) for i in range(n)])


if __name__ == "__main__":
    with open(os.path.join(HERE, "lr_schedule.json")) as f:
        d = json.load(f)
    d["ssl"] = ssl_sequences()
    with open(os.path.join(HERE, "lr_schedule.json"), "w") as f:
        json.dump(d, f)
    print("ssl", {k: (len(v), v[:2], v[-1]) for k, v in d["ssl"].items() if k != "config"})
