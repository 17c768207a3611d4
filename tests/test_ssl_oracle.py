"""Pins oracle/ssl_oracle.py (DINOv2 heads and losses) against vectors recorded from the unmodified
reference classes (tests/golden/make_golden_ssl.py).  CPU only; both sides are fp32 CPU torch: bar 1e-5 relative."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ssl_oracle as S


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "ssl_small.json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(golden_dir, "ssl_small.npz"))
    heads = {who: {k: torch.as_tensor(arr[f"{who}/{k}"]) for k in meta["head_keys"]}
             for who in ("student_head", "teacher_head")}
    return meta, arr, heads


def test_mask_bookkeeping(gold):
    meta, arr, _ = gold
    masks = torch.as_tensor(arr["masks"])
    assert torch.equal(S.mask_indices_of(masks), torch.as_tensor(arr["mask_indices_list"]))
    assert rel(S.masks_weight_of(masks), arr["masks_weight"]) < 1e-7


def test_components_match_reference(gold):
    meta, arr, heads = gold
    B, K, ng, nl = meta["B"], meta["K"], meta["n_global"], meta["n_local"]
    masks = torch.as_tensor(arr["masks"])
    idx = S.mask_indices_of(masks)
    mw = S.masks_weight_of(masks)
    n_masked = idx.shape[0]
    dino_c, ibot_c = torch.zeros(1, K), torch.zeros(1, 1, K)
    for step, temp in enumerate(meta["teacher_temps"]):
        tag = f"s{step}/"
        g = lambda n: torch.as_tensor(arr[tag + n])                                   # noqa: E731
        s_local = g("in/s_local").requires_grad_(True)
        s_global = g("in/s_global").requires_grad_(True)
        s_patch = g("in/s_patch").requires_grad_(True)
        t_cls, t_patch = g("in/t_cls"), g("in/t_patch")
        # teacher: head, targets, centres (the centre used at step k is the one updated with step k-1's batch)
        a, b = t_cls.chunk(2)
        t_out = S.dino_head_forward(heads["teacher_head"], torch.cat((torch.cat((b, a)), t_patch.flatten(0, 1)[idx])))
        assert rel(t_out, g("teacher_head_out")) < 1e-5
        assert rel(dino_c, g("dino_center_used")) < 1e-5 or float(dino_c.abs().sum()) == 0
        assert rel(ibot_c, g("ibot_center_used")) < 1e-5 or float(ibot_c.abs().sum()) == 0
        t_cls_out, t_patch_out = t_out[:ng * B], t_out[ng * B:]
        t_dino = S.softmax_center_teacher(t_cls_out, dino_c, temp)
        t_ibot = S.softmax_center_teacher(t_patch_out.unsqueeze(0), ibot_c, temp).squeeze(0)
        assert rel(t_dino, g("t_dino")) < 1e-5 and rel(t_ibot, g("t_ibot")) < 1e-5
        dino_c = S.dino_center_update(dino_c, t_cls_out, meta["center_momentum"])
        ibot_c = S.ibot_center_update(ibot_c, t_patch_out.unsqueeze(0), meta["center_momentum"])
        # student: head and the four loss terms
        student = {k: v.clone().requires_grad_(True) for k, v in heads["student_head"].items()}
        s_out = S.dino_head_forward(student, torch.cat((s_local, s_global, s_patch.flatten(0, 1)[idx])))
        assert rel(s_out, g("student_head_out")) < 1e-5
        s_l, s_g, s_p = s_out[:nl * B], s_out[nl * B:(nl + ng) * B], s_out[(nl + ng) * B:]
        t_list = t_dino.view(ng, -1, K)
        l_local = S.dino_loss(s_l.chunk(nl), list(t_list), meta["student_temp"])
        l_global = S.dino_loss([s_g], [t_list.flatten(0, 1)], meta["student_temp"])
        l_koleo = sum(S.koleo_loss(p) for p in s_global.chunk(2))
        l_ibot = S.ibot_loss_masked(s_p, t_ibot, masks, n_masked_patches=n_masked, masks_weight=mw)
        for name, v in dict(dino_local=l_local, dino_global=l_global, koleo=l_koleo, ibot=l_ibot,
                            ibot_default_weight=S.ibot_loss_masked(s_p, t_ibot, masks)).items():
            ref = float(arr[tag + "loss/" + name])
            assert abs(float(v) - ref) <= 1e-5 * abs(ref), (name, float(v), ref)
        (0.3 * l_local + 0.7 * l_global + 0.1 * l_koleo + 0.5 * l_ibot).backward()
        assert rel(s_local.grad, g("grad/s_local")) < 1e-5
        assert rel(s_global.grad, g("grad/s_global")) < 1e-5
        assert rel(s_patch.grad, g("grad/s_patch")) < 1e-5
        # the reference parametrises the last layer by (weight_g, weight_v): same keys, same gradients
        for k, p in student.items():
            assert rel(p.grad, g("grad/head/" + k)) < 2e-5, k
    assert rel(dino_c, arr["final/dino_center"]) < 1e-5
    assert rel(ibot_c, arr["final/ibot_center"]) < 1e-5


def test_objective_assembly_is_consistent_with_its_pinned_parts(gold):
    """`ssl_objective` (models.py:212-433, assembly unpinned) must equal the documented weighting of the pinned parts:
    dino terms / (2 + 16), global term and iBOT term x 2 (loss_scales), iBOT x 1/2, KoLeo x its weight."""
    meta, arr, heads = gold
    B, K, ng, nl = meta["B"], meta["K"], meta["n_global"], meta["n_local"]
    masks = torch.as_tensor(arr["masks"])
    g = lambda n: torch.as_tensor(arr["s0/" + n])                                      # noqa: E731
    total, parts, (c_d, c_i) = S.ssl_objective(
        heads["student_head"], heads["teacher_head"], student_local_cls=g("in/s_local"), student_global_cls=g("in/s_global"),
        student_global_patch=g("in/s_patch"), teacher_global_cls=g("in/t_cls"), teacher_global_patch=g("in/t_patch"),
        masks=masks, dino_center=torch.zeros(1, K), ibot_center=torch.zeros(1, 1, K), teacher_temp=meta["teacher_temps"][0],
        n_local_crops=nl, n_global_crops=ng, dino_loss_weight=1.0, koleo_loss_weight=0.1, ibot_loss_weight=1.0)
    terms = ng * (ng - 1) + nl * ng
    want = (float(arr["s0/loss/dino_local"]) / terms + float(arr["s0/loss/dino_global"]) * 2 / terms
            + 0.1 * float(arr["s0/loss/koleo"]) + float(arr["s0/loss/ibot"]) * 2 * 0.5)
    assert abs(float(total) - want) <= 1e-5 * abs(want)
    assert abs(float(parts["ibot_loss"]) - float(arr["s0/loss/ibot"]) / 2) <= 1e-5 * abs(float(arr["s0/loss/ibot"]))
    assert rel(c_d, arr["s1/dino_center_used"]) < 1e-5 and rel(c_i, arr["s1/ibot_center_used"]) < 1e-5


def test_ema_update():
    t = {"w": torch.ones(3), "inds": torch.arange(3)}
    s = {"w": torch.full((3,), 3.0), "inds": torch.arange(3) + 1}
    S.ema_update(t, s, 0.75)
    assert torch.allclose(t["w"], torch.full((3,), 1.5)) and torch.equal(t["inds"], torch.arange(3))


def test_weight_norm_matches_torch():
    lin = torch.nn.utils.weight_norm(torch.nn.Linear(7, 5, bias=False))
    with torch.no_grad():
        lin.weight_g.mul_(torch.rand(5, 1) + 0.5)
    x = torch.randn(4, 7)
    assert torch.allclose(torch.nn.functional.linear(x, S.weight_norm_weight(lin.weight_g, lin.weight_v)), lin(x), atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# the whole self-supervised step, pinned against the UNMODIFIED reference (tests/golden/make_golden_ssl_step.py runs
# DINOv2.forward / update_teacher on the CPU through tests/golden/xformers_shim.py: "pinned modulo the xformers shim")
# ---------------------------------------------------------------------------------------------------------------------
def _seeded_fill(shapes, int_arrays, seed):
    """The generator's `seeded_fill`: floating tensors in sorted key order from one torch.Generator."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        if k in int_arrays:
            sd[k] = torch.as_tensor(int_arrays[k])
            continue
        n = torch.randn(shapes[k], generator=g)
        gain = k.endswith(("norm.weight", "norm1.weight", "norm2.weight", ".gamma", "weight_g"))
        sd[k] = 1.0 + 0.1 * n if gain else 0.05 * n
    return sd


def test_ssl_step_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "ssl_step_tiny.json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(golden_dir, "ssl_step_tiny.npz"))
    cfg = meta["cfg"]
    ints = lambda who: {k[len(who) + 5:]: arr[k] for k in arr.files if k.startswith(who + "_int/")}    # noqa: E731
    student = _seeded_fill(meta["student_keys"], ints("student"), 11)
    teacher = _seeded_fill(meta["teacher_keys"], ints("teacher"), 12)
    # the reference builds the teacher as its own model and then loads the student's state dict into it, buffers
    # included (models.py:138): same column indices on both sides, bit for bit
    for k, v in ints("student").items():
        assert np.array_equal(v, ints("teacher")[k])
    assert meta["teacher_trainable"] == []
    trainable = meta["trainable"]
    assert all(("proj_weight1" in n or "proj_bias1" in n or n.startswith("dino_head.")) for n in trainable)
    for n in trainable:
        student[n].requires_grad_(True)
    glob, loc = torch.as_tensor(arr["in/global"]), torch.as_tensor(arr["in/local"])
    masks = torch.as_tensor(arr["in/masks"])
    dino_c, ibot_c = torch.zeros(1, cfg["K"]), torch.zeros(1, 1, cfg["K"])
    for step in range(2):
        tag = f"s{step}/"
        for n in trainable:
            student[n].grad = None
        loss, parts, (dino_c, ibot_c) = S.ssl_step(
            student, teacher, glob, loc, masks, dino_c, ibot_c, teacher_temp=cfg["teacher_temp"], patch=cfg["patch"],
            depth=cfg["depth"], num_heads=cfg["num_heads"], n_local_crops=cfg["n_local"],
            n_global_crops=cfg["n_global"], dino_loss_weight=cfg["dino_w"], koleo_loss_weight=cfg["koleo_w"],
            ibot_loss_weight=cfg["ibot_w"])
        loss.backward()
        assert rel(loss.detach(), arr[tag + "loss"]) < 1e-5, (step, float(loss), float(arr[tag + "loss"]))
        for k in ("dino_local_crops_loss", "dino_global_crops_loss", "koleo_loss", "ibot_loss"):
            assert rel(parts[k].detach(), arr[tag + "loss/" + k]) < 1e-5, (step, k)
        for n in trainable:
            assert rel(student[n].grad, arr[tag + "grad/" + n]) < 2e-5, (step, n, rel(student[n].grad, arr[tag + "grad/" + n]))
        S.ema_update(teacher, student, cfg["momentum"])
        with torch.no_grad():
            for n in trainable:
                student[n].add_(student[n].grad, alpha=-0.05)
        for k in ("backbone.blocks.1.attn.proj_weight1", "dino_head.mlp.0.weight"):
            assert rel(teacher[k], arr[tag + "teacher_after/" + k]) < 1e-6, (step, k)
    assert rel(dino_c, arr["final/dino_center"]) < 1e-5
    assert rel(ibot_c, arr["final/ibot_center"]) < 1e-5


def test_sinkhorn_knopp_matches_reference(golden_dir):
    arr = np.load(os.path.join(golden_dir, "ssl_sk_small.npz"))
    t_cls, t_patch = torch.as_tensor(arr["t_cls"]), torch.as_tensor(arr["t_patch"])
    assert rel(S.sinkhorn_knopp(t_cls, 0.04), arr["dino_sk_t0.04_it3"]) < 1e-5
    assert rel(S.sinkhorn_knopp(t_cls, 0.07, n_iterations=1), arr["dino_sk_t0.07_it1"]) < 1e-5
    assert rel(S.sinkhorn_knopp(t_patch, 0.04, n_samples_world=23), arr["ibot_sk_t0.04_it3"]) < 1e-5
    assert rel(S.sinkhorn_knopp(t_cls, 0.04), arr["dino_sk_t0.04_it3_in_group"]) < 1e-5
    assert float((S.sinkhorn_knopp(t_cls, 0.04).sum(-1) - 1).abs().max()) < 1e-5        # an assignment per sample
