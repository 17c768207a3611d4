"""Checkpoint and index interchange with the reference.

What the reference writes and reads, and what this module therefore produces and accepts:

  * session file `<save_dir>/<model_name>.pth` (src/defaults/bases.py:448-468): a `torch.save`d dict with keys
    `iters, state_dict, original_state, optimizer, epoch, parameters, best_val_target` (+ `scaler` under fp16 AMP,
    which the bf16 path has no use for).  `state_dict` is the Classifier's CPU state dict -- for an APLA model it
    carries `backbone.blocks.{i}.attn.{inds, qkv.*, proj_weight1, proj_weight2, proj_bias1, proj_bias2}`;
    `optimizer` is `torch.optim.AdamW.state_dict()` moved to the CPU (src/utils/_utils.py:58-75) over the two
    parameter groups of `DefaultWrapper.get_params_groups` (src/defaults/wrappers.py:205-221): group 0 = trainable
    tensors with more than one dimension whose name does not end in `.bias` (weight-decayed), group 1 = the rest
    (`weight_decay` 0), parameter ids counted in `named_parameters()` order inside each group, group 0 first.
  * `load_session` (bases.py:405-433): `model.load_state_dict(checkpoint['state_dict'])`, then iters / epoch /
    optimizer unless `restore_only_model`.
  * `load_from_pretrained` (src/utils/pretrained_loader.py:23-39): for paths containing 'apla' / 'fastadapt' a
    non-strict load that must have no missing keys and only `partial_size` unexpected ones; strict otherwise.
  * index files `inds-*.json` (src/apla/apla_vit.py:20-24; params/**/inds-vit_b-rand_128.json): `{"block_i": [r ints]}`,
    the trainable rows of every block in the order they sit in `proj_weight1`.

The functions here are host-side format code (CPU torch, exercised by the CPU tests against torch.optim.AdamW itself);
`FineTuneEngine.{state_dict, optimizer_state_dict, load_state_dict, load_optimizer_state_dict, save_session,
load_session}` in engine.py move the tensors between these dicts and the device arenas.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

SESSION_KEYS = ("iters", "state_dict", "original_state", "optimizer", "epoch", "parameters", "best_val_target")


def is_regularized(name: str, shape: Sequence[int]) -> bool:
    """wrappers.py:216-219: biases and 1-D tensors are not weight-decayed."""
    return not (name.endswith(".bias") or len(shape) == 1)


def optimizer_param_order(named_shapes: Sequence[Tuple[str, Sequence[int]]]) -> Tuple[List[str], List[str]]:
    """Trainable names in `named_parameters()` order -> (group 0 names, group 1 names): torch numbers the parameters
    of an optimizer state dict consecutively through the groups, so the id of a tensor is its position in
    `group0 + group1`."""
    reg = [n for n, s in named_shapes if is_regularized(n, s)]
    noreg = [n for n, s in named_shapes if not is_regularized(n, s)]
    return reg, noreg


def optimizer_state_dict(named_shapes: Sequence[Tuple[str, Sequence[int]]], exp_avg: Dict[str, torch.Tensor],
                         exp_avg_sq: Dict[str, torch.Tensor], step: int, *, lr: float, betas=(0.9, 0.999),
                         eps: float = 1e-8, weight_decay: float = 1e-5) -> dict:
    """The `torch.optim.AdamW.state_dict()` the reference would hold after `step` updates with these moments.

    The `param_groups` entries are taken from a real AdamW built over empty tensors of the same shapes, so the set of
    hyper-parameter keys is whatever the installed torch writes (it changed between releases) and
    `torch.optim.AdamW.load_state_dict` accepts the result unchanged."""
    reg, noreg = optimizer_param_order(named_shapes)
    shapes = dict((n, tuple(s)) for n, s in named_shapes)
    dummy = {n: torch.nn.Parameter(torch.empty(shapes[n])) for n in reg + noreg}
    groups = [{"params": [dummy[n] for n in reg]}, {"params": [dummy[n] for n in noreg], "weight_decay": 0.0}]
    opt = torch.optim.AdamW(groups, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
    sd = opt.state_dict()
    state = {}
    if step > 0:
        for i, n in enumerate(reg + noreg):
            state[i] = {"step": torch.tensor(float(step)),
                        "exp_avg": exp_avg[n].detach().to("cpu", torch.float32).reshape(shapes[n]).clone(),
                        "exp_avg_sq": exp_avg_sq[n].detach().to("cpu", torch.float32).reshape(shapes[n]).clone()}
    return {"state": state, "param_groups": sd["param_groups"]}


def split_optimizer_state(opt_sd: dict, named_shapes: Sequence[Tuple[str, Sequence[int]]]):
    """Inverse of `optimizer_state_dict`: -> (exp_avg by name, exp_avg_sq by name, step, hyper-parameters of group 0).

    Raises ValueError when the dict does not describe AdamW over exactly these tensors (count, grouping or shapes)."""
    reg, noreg = optimizer_param_order(named_shapes)
    shapes = dict((n, tuple(s)) for n, s in named_shapes)
    groups = opt_sd["param_groups"]
    if len(groups) != 2 or len(groups[0]["params"]) != len(reg) or len(groups[1]["params"]) != len(noreg):
        raise ValueError(f"optimizer state has groups of {[len(g['params']) for g in groups]} tensors, "
                         f"this model trains {len(reg)} + {len(noreg)}")
    ids = list(groups[0]["params"]) + list(groups[1]["params"])
    m, v, steps = {}, {}, set()
    for pid, n in zip(ids, reg + noreg):
        st = opt_sd["state"].get(pid)
        if not st:                      # a parameter the optimiser has not touched yet
            m[n] = torch.zeros(shapes[n])
            v[n] = torch.zeros(shapes[n])
            steps.add(0)
            continue
        if tuple(st["exp_avg"].shape) != shapes[n]:
            raise ValueError(f"optimizer state {pid} has shape {tuple(st['exp_avg'].shape)}, {n} is {shapes[n]}")
        m[n] = st["exp_avg"].detach().to("cpu", torch.float32)
        v[n] = st["exp_avg_sq"].detach().to("cpu", torch.float32)
        steps.add(int(float(st["step"])))
    if len(steps) != 1:
        raise ValueError(f"optimizer state holds different step counts {sorted(steps)}: not one AdamW over all tensors")
    g0 = groups[0]
    hyper = dict(lr=float(g0["lr"]), betas=tuple(g0["betas"]), eps=float(g0["eps"]),
                 weight_decay=float(g0["weight_decay"]))
    return m, v, steps.pop(), hyper


def model_to_cpu_state(model: torch.nn.Module) -> "OrderedDict[str, torch.Tensor]":
    """src/utils/_utils.py:49-55."""
    return OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())


def session_path(save_dir: Optional[str] = None, model_name: Optional[str] = None, model_path: Optional[str] = None) -> str:
    """bases.py:435-446: `<save_dir>/<model_name>.pth`, or `abspath(model_path) + '.pth'`."""
    if model_path is not None:
        return os.path.abspath(model_path) + ".pth"
    if save_dir is None or model_name is None:
        raise AttributeError("save_dir not found. Please specify the saving directory")
    os.makedirs(save_dir, exist_ok=True)
    return os.path.join(save_dir, model_name) + ".pth"


def save_session(path: str, *, state_dict, optimizer: dict, iters: int, epoch: int, original_state=None,
                 parameters=None, best_val_target=None) -> str:
    """Write the reference's session file (bases.py:455-466).  `path` is the full file name."""
    state = {"iters": int(iters), "state_dict": state_dict,
             "original_state": original_state if original_state is not None else state_dict,
             "optimizer": optimizer, "epoch": int(epoch), "parameters": parameters,
             "best_val_target": best_val_target}
    torch.save(state, path)
    return path


def load_session_file(path: str) -> dict:
    """`torch.load` of a session file; FileNotFoundError with the reference's wording when it is absent
    (pretrained_loader.py:12-20).  The reference pickles an EasyDict under `parameters`, hence weights_only=False --
    only load files you wrote or trust."""
    path = os.path.abspath(path)
    if not os.path.isfile(path):
        raise FileNotFoundError('Model "{}" is not present in "{}"'.format(os.path.basename(path), os.path.dirname(path)))
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if "state_dict" not in ckpt:
        raise KeyError(f"{path} is not a session file: no 'state_dict' (keys: {sorted(ckpt)})")
    return ckpt


def load_from_pretrained(model: torch.nn.Module, path: str, strict: bool = False):
    """pretrained_loader.py:23-39 with its two branches and its assertions."""
    sd = load_session_file(path)["state_dict"]
    if "fastadapt" in path or "apla" in path:
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert missing == [], f"There are unexpected keys!: \n{unexpected}\n"
        assert all("partial_size" in k for k in unexpected)
        return missing, unexpected
    res = model.load_state_dict(sd, strict=True)
    return res.missing_keys, res.unexpected_keys


# ---------------------------------------------------------------------------------------------------------------------
# index files
# ---------------------------------------------------------------------------------------------------------------------
def inds_table(model: torch.nn.Module) -> Dict[str, List[int]]:
    """`{"block_i": trainable rows}` of an APLA model (backbone or Classifier), in `proj_weight1` row order."""
    bb = getattr(model, "backbone", model)
    out = {}
    for i, blk in enumerate(bb.blocks):
        at = blk.attn
        if not hasattr(at, "trainable_inds"):
            raise ValueError(f"block {i} carries no APLA attention (multi-GPU 'full' mode trains every row)")
        out[f"block_{i}"] = [int(j) for j in torch.as_tensor(at.trainable_inds).tolist()]
    return out


def save_inds_json(model: torch.nn.Module, path: str) -> str:
    """Write the index file a multi-GPU partial run needs (apla_vit.py:77): every rank then reads the same rows."""
    with open(path, "w") as f:
        json.dump(inds_table(model), f)
    return path


def load_inds_json(path: str, dim: int) -> Dict[str, torch.Tensor]:
    """`{"block_i": full index permutation}`: the file's trainable rows followed by the ascending complement
    (apla_vit.py:20-24)."""
    with open(path) as f:
        table = json.load(f)
    out = {}
    for k, chosen in table.items():
        taken = set(chosen)
        if len(taken) != len(chosen) or (chosen and not (0 <= min(chosen) and max(chosen) < dim)):
            raise ValueError(f"{path}: {k} is not a set of distinct rows in [0, {dim})")
        out[k] = torch.tensor(list(chosen) + [j for j in range(dim) if j not in taken])
    return out
