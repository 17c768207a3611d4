"""Shared test helpers: build the product-side model the way the golden fixtures were built on the reference."""
import json
import os

import numpy as np
import torch

from apla_b200.config import AplaConfig
from apla_b200.hostvit import VitArch, build_classifier

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY = VitArch(128, 2, 2)

CASES = {
    # name: (arch, table img size, patch)
    "tiny_r16": (TINY, 56, 14),
    "tiny_interp_r128": (TINY, 56, 14),
    "tiny_full_multigpu": (TINY, 56, 14),
    "c1_vits16_r32": ("vit_small", 224, 16),
    "c1_vits16_r32_pert": ("vit_small", 224, 16),
    "c2_vitb14_r8": ("vit_base", 518, 14),
    "c3_vitb14_r768": ("vit_base", 518, 14),
    "c2_vitb14_inds128": ("vit_base", 518, 14),
    "c5_vitb14_518_r768": ("vit_base", 518, 14),      # 1370 tokens: the streaming long-sequence attention kernels
    "vitl14_r128": ("vit_large", 518, 14),            # C4's backbone shape (D 1024, 16 heads, 24 blocks)
}


def perturb_module(model, seed=7, scale=0.05):
    """Same perturbation as oracle.perturb_state / make_golden.perturb_module (sorted keys, one generator)."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    with torch.no_grad():
        for k in sorted(sd.keys()):
            t = sd[k]
            if not t.is_floating_point():
                continue
            t.add_(torch.randn(t.shape, generator=g) * scale * (0.2 if t.dim() > 1 else 1.0))


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(GOLDEN, name + ".npz"))
    return meta, arr


def build_case(name):
    """-> (model on CPU, meta, arrays) constructed exactly like tests/golden/make_golden.py did on the reference."""
    meta, arr = load_golden(name)
    m = meta["meta"]
    arch, table_img, patch = CASES[name]
    kw = {}
    if name == "c2_vitb14_inds128":
        kw["inds_path"] = os.path.join(GOLDEN, "inds-vit_b-rand_128.json")
    cfg = AplaConfig(m["apla_cfg"]["partial_size"], **kw)
    model = build_classifier(arch, img_size=table_img, patch_size=patch, n_classes=m["n_classes"], apla_config=cfg,
                             is_multi_gpu=m["is_multi_gpu"], seed=0)
    if m["perturb"]:
        perturb_module(model)
    return model, meta, arr


def synthetic_batch(batch, img, n_classes, seed=1234, rank=0):
    g = torch.Generator().manual_seed(seed + rank)
    images = torch.randn(batch, 3, img, img, generator=g)
    labels = torch.randint(0, n_classes, (batch,), generator=g)
    return images, labels


def rel(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))
