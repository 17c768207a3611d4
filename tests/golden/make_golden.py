"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src with three stub modules for absent, unused imports (easydict,
matplotlib, timm), builds the reference model exactly as
``Classifier.__init__`` does (src/defaults/models.py:39-65), runs ``Trainer.global_step``'s
arithmetic in fp32 on CPU (src/defaults/trainer.py:106-138 without AMP) and stores small
fixtures next to this file.  Nothing here is used at run time by the product; the fixtures are
what ``tests/test_oracle_golden.py`` pins ``oracle/apla_oracle.py`` against, on any machine.
"""
import hashlib
import json
import os
import sys
import types
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"


class _AttrDict(dict):
    """Stand-in for easydict.EasyDict: answers hasattr() and `in` (apla_vit.py:12,77)."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def import_reference():
    ed = types.ModuleType("easydict"); ed.EasyDict = _AttrDict
    sys.modules.setdefault("easydict", ed)
    mpl = types.ModuleType("matplotlib"); mpl.pylab = types.ModuleType("matplotlib.pylab")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pylab", mpl.pylab)
    sys.modules.setdefault("timm", types.ModuleType("timm"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import builtins
    from utils.transformers import vit
    from apla import apla_vit
    return vit, apla_vit


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


class RefClassifier(nn.Module):
    """Same construction order as src/defaults/models.py:39-65 without the params plumbing."""
    def __init__(self, vit, apla_vit, factory_kw, apla_cfg, n_classes, is_multi_gpu, ctor=None):
        super().__init__()
        model = ctor(vit) if ctor else None
        self.backbone = apla_vit.build_apla(config=apla_cfg, model=model, attn_class="apla_attn",
                                            is_multi_gpu=is_multi_gpu)
        self.backbone.fc = nn.Identity()
        self.fc = nn.Linear(self.backbone.num_features, n_classes)

    def forward(self, x):
        return self.fc(self.backbone(x))


def perturb_module(model: nn.Module, seed=7, scale=0.05):
    """Mirror of oracle.perturb_state: sorted state_dict keys, floating tensors only."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    with torch.no_grad():
        for k in sorted(sd.keys()):
            t = sd[k]
            if not t.is_floating_point():
                continue
            t.add_(torch.randn(t.shape, generator=g) * scale * (0.2 if t.dim() > 1 else 1.0))
    # APLA modules keep CPU copies of parameters in .data, state_dict tensors alias them -> in place is enough


def param_groups(model):
    """DefaultWrapper.get_params_groups  src/defaults/wrappers.py:205-221."""
    reg, noreg = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (noreg if (name.endswith(".bias") or len(p.shape) == 1) else reg).append(p)
    return [{"params": reg}, {"params": noreg, "weight_decay": 0.0}]


ONLY = set(sys.argv[1:])        # optional: names of the cases to (re)generate; default all


def run_case(name, vit, apla_vit, *, ctor, apla_cfg, n_classes, batch, img, is_multi_gpu=False,
             perturb=True, sub=1, steps=1, full_grads=False):
    import contextlib, io
    if ONLY and name not in ONLY:
        return
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = RefClassifier(vit, apla_vit, None, _AttrDict(apla_cfg), n_classes, is_multi_gpu, ctor=ctor)
    if perturb:
        perturb_module(model)
    g = torch.Generator().manual_seed(1234)
    images = torch.randn(batch, 3, img, img, generator=g)
    labels = torch.randint(0, n_classes, (batch,), generator=g)

    out = {"meta": {"name": name, "n_classes": n_classes, "batch": batch, "img": img,
                    "apla_cfg": {k: v for k, v in apla_cfg.items() if k != "inds_path"},
                    "is_multi_gpu": is_multi_gpu, "perturb": perturb, "sub": sub}}
    arrays = {}
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    out["weights_digest"] = {k: digest(v) for k, v in sd0.items()}
    out["state_keys"] = list(sd0.keys())
    out["trainable"] = [n for n, p in model.named_parameters() if p.requires_grad]
    for k, v in sd0.items():
        if k.endswith(".inds"):
            arrays["inds/" + k] = v.numpy().astype(np.int16)

    opt = torch.optim.AdamW(param_groups(model), lr=3e-5, weight_decay=1e-5)
    for s in range(steps):
        opt.zero_grad()
        logits = model(images)
        loss = F.cross_entropy(logits, labels)
        loss.backward()
        raw = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        tag = f"s{s}/"
        arrays[tag + "logits"] = logits.detach().numpy()
        arrays[tag + "loss"] = loss.detach().numpy()
        arrays[tag + "grad_norm"] = gn.detach().numpy()
        for n, gr in raw.items():
            flat = gr.flatten()
            arrays[tag + "gnorm/" + n] = torch.linalg.vector_norm(flat).numpy()
            arrays[tag + "grad/" + n] = (flat if full_grads else flat[::sub]).numpy()
        for n, p in model.named_parameters():
            if p.requires_grad:
                flat = p.detach().flatten()
                arrays[tag + "param/" + n] = (flat if full_grads else flat[::sub]).clone().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"{name}: loss {float(loss):.6f} grad_norm {float(gn):.6f} "
          f"{os.path.getsize(os.path.join(HERE, name + '.npz')) / 1e3:.0f} kB")


def main():
    vit, apla_vit = import_reference()
    ln = partial(nn.LayerNorm, eps=1e-6)
    bc = _AttrDict(has_layerscale=True, layerscale_init_values=1.0)

    # the reference's only data fixture: params/**/inds-vit_b-rand_128.json
    src = "/root/reference/params/finetune/dinov2/ImageNet/vit_b/inds-vit_b-rand_128.json"
    dst = os.path.join(HERE, "inds-vit_b-rand_128.json")
    with open(src) as f:
        inds128 = json.load(f)
    with open(dst, "w") as f:
        json.dump(inds128, f)

    # tiny: direct VisionTransformer ctor, everything stored in full, two optimiser steps
    def tiny(v):
        return v.VisionTransformer(img_size=[56], patch_size=14, embed_dim=128, depth=2, num_heads=2,
                                   mlp_ratio=4, qkv_bias=True, norm_layer=ln, block_conf=bc)
    run_case("tiny_r16", vit, apla_vit, ctor=tiny, apla_cfg={"partial_size": 16}, n_classes=10,
             batch=4, img=56, full_grads=True, steps=2)
    # tiny with pos-embed interpolation (table built for 56px = 16 patches, input 28px = 4 patches)
    run_case("tiny_interp_r128", vit, apla_vit, ctor=tiny, apla_cfg={"partial_size": 128}, n_classes=10,
             batch=3, img=28, full_grads=True)
    # multi-GPU 'full': stock attention kept, attn.proj trainable (apla_vit.py:65-75)
    run_case("tiny_full_multigpu", vit, apla_vit, ctor=tiny, apla_cfg={"partial_size": "full"}, n_classes=10,
             batch=4, img=56, is_multi_gpu=True, full_grads=True)

    # C1: ViT-S/16 r=32 B=8 224px (BASELINE.json configs[0]) -- unperturbed AND perturbed
    def c1(v):
        return v.vit_small(pretrained=False, img_size=[224], patch_size=16, pretrained_type="dinov2",
                           is_memory_efficient=True, block_conf=bc)
    run_case("c1_vits16_r32", vit, apla_vit, ctor=c1, apla_cfg={"partial_size": 32}, n_classes=555,
             batch=8, img=224, perturb=False, sub=29)
    run_case("c1_vits16_r32_pert", vit, apla_vit, ctor=c1, apla_cfg={"partial_size": 32}, n_classes=555,
             batch=8, img=224, perturb=True, sub=29)

    # C2 shape: ViT-B/14, 518-px table interpolated to 224 px, r=8, batch cut to 2 for CPU
    def c2(v):
        return v.vit_base(pretrained=False, img_size=[518], patch_size=14, pretrained_type="dinov2",
                          is_memory_efficient=True, block_conf=bc)
    run_case("c2_vitb14_r8", vit, apla_vit, ctor=c2, apla_cfg={"partial_size": 8}, n_classes=555,
             batch=2, img=224, sub=31)
    # C3 shape single-process: partial_size == dim
    run_case("c3_vitb14_r768", vit, apla_vit, ctor=c2, apla_cfg={"partial_size": 768}, n_classes=555,
             batch=2, img=224, sub=257)
    # inds_path fixture, multi-GPU partial mode (apla_vit.py:77, 20-24)
    run_case("c2_vitb14_inds128", vit, apla_vit, ctor=c2,
             apla_cfg={"partial_size": 128, "inds_path": dst}, n_classes=555,
             batch=2, img=224, is_multi_gpu=True, sub=127)

    # C5 backbone shape: ViT-B/14 at 518 px = 1370 tokens, pos table used un-interpolated (vit.py:424-425), r = 768,
    # 2 images per GPU (mmseg config); the segmentation decoder is a third-party framework, so the step ends in the
    # same classifier head as C2 -- what is pinned here is the long-sequence block path
    run_case("c5_vitb14_518_r768", vit, apla_vit, ctor=c2, apla_cfg={"partial_size": 768}, n_classes=555,
             batch=2, img=518, sub=1543)

    # C4 backbone shape: ViT-L/14 (D 1024, 16 heads, 24 blocks), r = 128, supervised head, batch 2
    def vl(v):
        return v.vit_large(pretrained=False, img_size=[518], patch_size=14, pretrained_type="dinov2",
                           is_memory_efficient=True, block_conf=bc)
    run_case("vitl14_r128", vit, apla_vit, ctor=vl, apla_cfg={"partial_size": 128}, n_classes=555,
             batch=2, img=224, sub=509)


if __name__ == "__main__":
    main()
