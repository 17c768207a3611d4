"""apla_b200/schedule.py against learning-rate sequences recorded from the reference's own scheduler classes
(LinearWarmup + CosineAnnealingLR inside MixedLRScheduler; tests/golden/make_golden_lr.py)."""
import json
import os

import pytest

from apla_b200.schedule import WarmupCosineSchedule

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "lr_schedule.json")) as f:
    GOLD = json.load(f)["cases"]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_lr_sequence_matches_reference(name):
    cfg, want = GOLD[name]["config"], GOLD[name]["lr"]
    s = WarmupCosineSchedule(cfg["max_lr"], cfg["warmup_iters"], cfg["steps_per_epoch"] * cfg["epochs"],
                             cosine_eta_min=cfg["cosine_eta_min"], cosine="CosineAnnealingLR" in cfg["types"])
    got = list(s)
    assert len(got) == len(want)
    for k, (a, b) in enumerate(zip(got, want)):
        assert abs(a - b) <= 1e-9 * abs(b), (name, k, a, b)


def test_known_quirks_of_the_reference_schedule():
    s = WarmupCosineSchedule(3e-5, 500, 1200)
    assert abs(s.lr_at(0) - 6.998e-8) < 1e-12                 # first step far below lr
    assert s.lr_at(500) > 3e-5                                # the ramp overshoots max_lr by one increment
    assert s.lr_at(501) < s.lr_at(500) and s.lr_at(1199) < 1.01e-6
    assert WarmupCosineSchedule(1e-3, 0, 18).warmup_iters == 1


def test_from_reference_yaml_mapping_and_apply():
    opt = dict(optimizer=dict(params=dict(lr=3e-5, weight_decay=1e-5)),
               scheduler=dict(type=["LinearWarmup", "CosineAnnealingLR"],
                              params=dict(LinearWarmup=dict(warmup_epochs=0, warmup_iters=500),
                                          CosineAnnealingLR=dict(eta_min=1e-6))))
    s = WarmupCosineSchedule.from_reference_config(opt, steps_per_epoch=200, epochs=6)
    want = GOLD["nabirds_apla"]["lr"]
    assert abs(s.lr_at(700) - want[700]) <= 1e-9 * want[700]

    class Eng:
        lr = 0.0
    e = Eng()
    assert s.apply(e, 3) == e.lr == s.lr_at(3)
    with pytest.raises(NotImplementedError):
        WarmupCosineSchedule.from_reference_config(dict(optimizer=opt["optimizer"], scheduler=dict(type="OneCycleLR", params={})), 10, 1)


def test_ssl_schedules_match_reference_cosine_scheduler():
    """The five schedules of the DINOv2 step against `CosineScheduler` / `build_schedulers` of the reference."""
    from apla_b200.schedule import build_ssl_schedules
    with open(os.path.join(HERE, "golden", "lr_schedule.json")) as f:
        ssl = json.load(f)["ssl"]
    c = ssl["config"]
    s = build_ssl_schedules(lr=c["lr"], lr_eta_min=c["eta_min"], lr_warmup_epochs=c["lr_warmup_epochs"],
                            weight_decay=c["weight_decay"], momentum_teacher=c["momentum_teacher"],
                            final_momentum_teacher=c["final_momentum_teacher"], warmup_teacher_temp=c["warmup_teacher_temp"],
                            teacher_temp=c["teacher_temp"], warmup_teacher_temp_epochs=c["warmup_teacher_temp_epochs"],
                            freeze_last_layer_epochs=c["freeze_last_layer_epochs"], iters_per_epoch=c["iters_per_epoch"],
                            epochs=c["epochs"])
    for name in ("lr", "wd", "momentum", "teacher_temp", "last_layer_lr"):
        want = ssl[name]
        for k, b in enumerate(want):
            a = s[name][k]
            assert abs(a - b) <= 1e-12 + 1e-9 * abs(b), (name, k, a, b)
    assert s["last_layer_lr"][0] == 0.0 and s["last_layer_lr"][c["iters_per_epoch"]] == s["lr"][c["iters_per_epoch"]]
    assert s["teacher_temp"][10 ** 6] == c["teacher_temp"]
