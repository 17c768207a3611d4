"""Checkpoint / index interchange with the reference (apla_b200/checkpoint.py).  CPU only.

The optimiser side is held against torch.optim.AdamW itself, built over the reference's two parameter groups
(src/defaults/wrappers.py:205-221): what we write must load into it unchanged and continue to the same parameters,
and what it writes must come back out of `split_optimizer_state` tensor for tensor."""
import json
import os

import numpy as np
import pytest
import torch

from apla_b200 import checkpoint as C
from apla_b200.config import AplaConfig
from apla_b200.hostvit import build_classifier
from helpers import GOLDEN, TINY


def _tiny(r=16, **kw):
    return build_classifier(TINY, img_size=56, patch_size=14, n_classes=10, apla_config=AplaConfig(r, **kw), seed=0,
                            is_multi_gpu=bool(kw))


def _ref_groups(model):
    """DefaultWrapper.get_params_groups, wrappers.py:205-221."""
    reg, noreg = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (noreg if (name.endswith(".bias") or len(p.shape) == 1) else reg).append(p)
    return [{"params": reg}, {"params": noreg, "weight_decay": 0.0}]


def _trainable(model):
    return [(n, tuple(p.shape)) for n, p in model.named_parameters() if p.requires_grad]


def _fake_steps(model, opt, n, seed=0):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n):
        for _, p in model.named_parameters():
            if p.requires_grad:
                p.grad = torch.randn(p.shape, generator=g) * 1e-2
        opt.step()


def test_param_group_order_is_the_reference_order():
    model = _tiny()
    named = _trainable(model)
    reg, noreg = C.optimizer_param_order(named)
    assert reg == [f"backbone.blocks.{i}.attn.proj_weight1" for i in range(2)] + ["fc.weight"]
    assert noreg == [f"backbone.blocks.{i}.attn.proj_bias1" for i in range(2)] + ["fc.bias"]
    groups = _ref_groups(model)
    by_id = {id(p): n for n, p in model.named_parameters()}
    assert [by_id[id(p)] for p in groups[0]["params"]] == reg
    assert [by_id[id(p)] for p in groups[1]["params"]] == noreg


def test_optimizer_state_round_trip_against_torch_adamw():
    model = _tiny()
    named = _trainable(model)
    opt = torch.optim.AdamW(_ref_groups(model), lr=3e-5, weight_decay=1e-5)
    _fake_steps(model, opt, 2)
    ref_sd = opt.state_dict()

    m, v, step, hyper = C.split_optimizer_state(ref_sd, named)
    assert step == 2 and hyper == dict(lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    ours = C.optimizer_state_dict(named, m, v, step, **hyper)
    assert ours["param_groups"] == ref_sd["param_groups"]
    assert sorted(ours["state"]) == sorted(ref_sd["state"])
    for pid, st in ref_sd["state"].items():
        assert float(ours["state"][pid]["step"]) == float(st["step"])
        assert torch.equal(ours["state"][pid]["exp_avg"], st["exp_avg"])
        assert torch.equal(ours["state"][pid]["exp_avg_sq"], st["exp_avg_sq"])

    # a second, fresh reference optimiser accepts our dict and continues to bit-identical parameters
    model2 = _tiny()
    model2.load_state_dict(model.state_dict())
    opt2 = torch.optim.AdamW(_ref_groups(model2), lr=1.0, weight_decay=0.5)     # overwritten by the loaded groups
    opt2.load_state_dict(ours)
    _fake_steps(model, opt, 1, seed=5)
    _fake_steps(model2, opt2, 1, seed=5)
    for (n, p), (_, q) in zip(model.named_parameters(), model2.named_parameters()):
        assert torch.equal(p, q), n


def test_optimizer_state_rejects_foreign_dicts():
    model = _tiny()
    named = _trainable(model)
    opt = torch.optim.AdamW(_ref_groups(model), lr=3e-5)
    _fake_steps(model, opt, 1)
    sd = opt.state_dict()
    with pytest.raises(ValueError):
        C.split_optimizer_state(sd, named[:-1])
    other = _tiny(r=32)
    with pytest.raises(ValueError):
        C.split_optimizer_state(sd, _trainable(other))
    # untouched optimiser: empty state -> zero moments, step 0
    fresh = torch.optim.AdamW(_ref_groups(model), lr=3e-5).state_dict()
    m, v, step, _ = C.split_optimizer_state(fresh, named)
    assert step == 0 and all(float(t.abs().sum()) == 0 for t in m.values())
    assert C.optimizer_state_dict(named, m, v, 0, lr=3e-5)["state"] == {}


def test_session_file_format_and_loaders(tmp_path):
    model = _tiny()
    named = _trainable(model)
    opt = torch.optim.AdamW(_ref_groups(model), lr=3e-5, weight_decay=1e-5)
    _fake_steps(model, opt, 3)
    m, v, step, hyper = C.split_optimizer_state(opt.state_dict(), named)
    path = C.session_path(save_dir=str(tmp_path / "ck"), model_name="apla_vit_tiny")
    assert path.endswith("ck/apla_vit_tiny.pth")
    assert C.session_path(model_path=str(tmp_path / "x")) == str(tmp_path / "x") + ".pth"
    with pytest.raises(AttributeError):
        C.session_path()
    C.save_session(path, state_dict=C.model_to_cpu_state(model), optimizer=C.optimizer_state_dict(named, m, v, step, **hyper),
                   iters=3, epoch=1, parameters={"partial_size": 16}, best_val_target=0.5)
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    assert tuple(ckpt) == C.SESSION_KEYS                          # bases.py:455-463, same keys in the same order
    assert ckpt["iters"] == 3 and ckpt["epoch"] == 1
    # I6: the state dict of an APLA classifier
    for k in ("inds", "qkv.weight", "qkv.bias", "proj_weight1", "proj_weight2", "proj_bias1", "proj_bias2"):
        assert f"backbone.blocks.0.attn.{k}" in ckpt["state_dict"]

    # the reference's load_session: plain load_state_dict into a same-config model
    fresh = _tiny()
    fresh.load_state_dict(C.load_session_file(path)["state_dict"])
    for (n, p), (_, q) in zip(model.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(p, q), n

    # load_from_pretrained: 'apla' in the path -> non-strict branch with its assertions
    other = _tiny()
    missing, unexpected = C.load_from_pretrained(other, path)
    assert missing == [] and unexpected == []
    assert torch.equal(other.fc.weight, model.fc.weight)
    # ... a model with more tensors than the file has -> missing keys -> AssertionError (pretrained_loader.py:29)
    bigger = build_classifier(TINY, img_size=56, patch_size=14, n_classes=10, apla_config=AplaConfig(16), seed=0)
    bigger.extra = torch.nn.Linear(2, 2)
    with pytest.raises(AssertionError):
        C.load_from_pretrained(bigger, path)
    # strict branch (no 'apla' / 'fastadapt' in the path)
    plain = str(tmp_path / "ck" / "vit_tiny.pth")
    os.rename(path, plain)
    assert "apla" not in plain
    C.load_from_pretrained(_tiny(), plain)
    with pytest.raises(RuntimeError):
        C.load_from_pretrained(bigger, plain)
    with pytest.raises(FileNotFoundError, match="is not present in"):
        C.load_from_pretrained(_tiny(), str(tmp_path / "nope.pth"))


def test_inds_json_round_trip(tmp_path):
    model = _tiny()
    p = C.save_inds_json(model, str(tmp_path / "inds-tiny-rand_16.json"))
    with open(p) as f:
        table = json.load(f)
    assert sorted(table) == ["block_0", "block_1"] and all(len(v) == 16 for v in table.values())
    # a multi-GPU partial build from that file (apla_vit.py:77, 20-24) selects the same rows in the same order,
    # frozen rows = ascending complement
    again = _tiny(inds_path=p)
    for a, b in zip(model.backbone.blocks, again.backbone.blocks):
        assert torch.equal(torch.as_tensor(a.attn.trainable_inds), torch.as_tensor(b.attn.trainable_inds))
        rest = torch.as_tensor(b.attn.freezed_inds)
        assert torch.equal(rest, rest.sort().values)
    assert C.inds_table(again) == table
    full = C.load_inds_json(p, 128)
    assert all(sorted(v.tolist()) == list(range(128)) for v in full.values())
    with open(tmp_path / "bad.json", "w") as f:
        json.dump({"block_0": [1, 1, 2]}, f)
    with pytest.raises(ValueError):
        C.load_inds_json(str(tmp_path / "bad.json"), 128)


def test_reference_inds_fixture_matches_the_reference_run():
    """The reference's own index file, read our way, gives the permutation the reference registered as `inds`
    (recorded by tests/golden/make_golden.py from the unmodified reference, case c2_vitb14_inds128)."""
    full = C.load_inds_json(os.path.join(GOLDEN, "inds-vit_b-rand_128.json"), 768)
    arr = np.load(os.path.join(GOLDEN, "c2_vitb14_inds128.npz"))
    for i in range(12):
        assert np.array_equal(full[f"block_{i}"].numpy().astype(np.int16), arr[f"inds/backbone.blocks.{i}.attn.inds"])
