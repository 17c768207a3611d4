"""Golden vectors for ONE self-supervised step of the reference's DINOv2 meta-architecture with APLA (BASELINE config C4,
scaled down), from the UNMODIFIED reference code run on the CPU.

    python tests/golden/make_golden_ssl_step.py        # build container only

`self_supervised/dinov2/models.py` refuses to import without xformers; `xformers_shim.py` (next to this file) supplies
`memory_efficient_attention`, `unbind` and `fmha.BlockDiagonalMask` in plain torch, so the fixture pins every line of the
reference on this path -- DINOv2.forward (models.py:207-433), DinoVisionTransformer.forward_features_list
(dinov2_vits.py:208-267), NestedTensorBlock.forward_nested (layers/block.py:244-288), APLA_MemEffAttention.forward
(apla/appla_attn_mem_eff.py:27-67), DINOHead and the three losses -- EXCEPT the xformers kernels (pinned modulo the shim).
Other accommodations, none touching reference source: `Tensor.cuda` is made a no-op (models.py:216-224 moves the batch
to the GPU), and a 2-block, 64-wide factory `vit_tiny_test` is registered in `dinov2_vits.__dict__` (build_model looks
factories up there, models.py:50) so that the fixture stays small.

Weights are NOT stored: after construction every floating tensor of student and teacher is overwritten by a seeded
fill (`seeded_fill`, sorted key order), which tests/test_ssl_oracle.py repeats; indices, inputs and all outputs are stored.
Output: tests/golden/ssl_step_tiny.npz / .json."""
import contextlib
import io
import json
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import xformers_shim  # noqa: E402

CFG = dict(embed_dim=64, depth=2, num_heads=1, patch=14, global_px=56, local_px=28, B=2, n_global=2, n_local=8,
           partial_size=16, K=256, head_hidden=128, head_bottleneck=64, teacher_temp=0.05, koleo_w=0.1, dino_w=1.0,
           ibot_w=1.0, momentum=0.994)


def seeded_fill(sd, seed):
    """Overwrite every floating tensor of a state dict in sorted key order: gains ~ 1 + 0.1 n, everything else 0.05 n."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k in sorted(sd.keys()):
            t = sd[k]
            if not t.is_floating_point():
                continue
            n = torch.randn(t.shape, generator=g)
            gain = k.endswith(("norm.weight", "norm1.weight", "norm2.weight", ".gamma", "weight_g"))
            t.copy_(1.0 + 0.1 * n if gain else 0.05 * n)


def make_batch(cfg, seed=1234):
    g = torch.Generator().manual_seed(seed)
    B, P = cfg["B"], (cfg["global_px"] // cfg["patch"]) ** 2
    glob = torch.randn(cfg["n_global"] * B, 3, cfg["global_px"], cfg["global_px"], generator=g)
    loc = torch.randn(cfg["n_local"] * B, 3, cfg["local_px"], cfg["local_px"], generator=g)
    masks = torch.zeros(cfg["n_global"] * B, P, dtype=torch.bool)
    for i in range(0, cfg["n_global"] * B, 2):              # mask_sample_probability 0.5: every other crop is masked
        n = int(torch.randint(2, P // 2 + 1, (1,), generator=g))
        masks[i, torch.randperm(P, generator=g)[:n]] = True
    return glob, loc, masks


def main():
    xformers_shim.install()
    from make_golden import _AttrDict, import_reference
    import_reference()
    from self_supervised.dinov2 import dinov2_vits as vits
    from self_supervised.dinov2 import models as M
    from self_supervised.dinov2.layers import MemEffAttention, NestedTensorBlock

    cfg = CFG

    def vit_tiny_test(patch_size=16, num_register_tokens=0, **kw):
        return vits.DinoVisionTransformer(patch_size=patch_size, embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                                          num_heads=cfg["num_heads"], mlp_ratio=4,
                                          block_fn=partial(NestedTensorBlock, attn_class=MemEffAttention),
                                          num_register_tokens=num_register_tokens, **kw)
    vits.__dict__["vit_tiny_test"] = vit_tiny_test
    torch.Tensor.cuda = lambda self, *a, **k: self          # the batch stays on the CPU

    A = _AttrDict
    head = dict(head_n_prototypes=cfg["K"], head_bottleneck_dim=cfg["head_bottleneck"], head_nlayers=3,
                head_hidden_dim=cfg["head_hidden"])
    params = A(
        system_params=A(which_GPUs="0"),
        crops_params=A(n_global_crops=cfg["n_global"], n_local_crops=cfg["n_local"]),
        model_params=A(
            backbone_type="vit_tiny_test", pretrained=False,
            transformers_params=A(student=A(pretrained_type="LVD142M-SSL", pre_img_size=cfg["global_px"],
                                            patch_size=cfg["patch"], drop_path_rate=0, drop_path_uniform=False,
                                            layerscale=1.0, ffn_layer="mlp", block_chunks=0, num_register_tokens=0,
                                            interpolate_antialias=False, interpolate_offset=0.1)),
            adaptation=A(mode="apla", params=A(partial_size=cfg["partial_size"])),
            dinov2=A(dino=A(loss_weight=cfg["dino_w"], koleo_loss_weight=cfg["koleo_w"], **head),
                     ibot=A(loss_weight=cfg["ibot_w"], mask_sample_probability=0.5, mask_ratio_min_max=[0.1, 0.5],
                            separate_head=False, **head),
                     centering="centering")))
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = M.DINOv2(params)
    model.train()
    student_sd, teacher_sd = model.student.state_dict(), model.teacher.state_dict()
    seeded_fill(student_sd, seed=11)
    seeded_fill(teacher_sd, seed=12)

    arrays = {}
    meta = dict(cfg=cfg, student_keys={k: list(v.shape) for k, v in student_sd.items()},
                teacher_keys={k: list(v.shape) for k, v in teacher_sd.items()},
                trainable=[n for n, p in model.student.named_parameters() if p.requires_grad],
                teacher_trainable=[n for n, p in model.teacher.named_parameters() if p.requires_grad])
    for k, v in student_sd.items():
        if not v.is_floating_point():
            arrays["student_int/" + k] = v.numpy()
    for k, v in teacher_sd.items():
        if not v.is_floating_point():
            arrays["teacher_int/" + k] = v.numpy()

    glob, loc, masks = make_batch(cfg)
    idx = masks.flatten().nonzero().flatten()
    mw = (1 / masks.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(masks)[masks]
    upperbound = int(idx.shape[0]) + 3                      # collate's bound is >= the number of masked patches
    batch = {"collated_global_crops": glob, "collated_local_crops": loc, "collated_masks": masks,
             "mask_indices_list": idx, "masks_weight": mw, "upperbound": upperbound,
             "n_masked_patches": torch.full((1,), idx.shape[0], dtype=torch.long)}
    arrays["in/global"], arrays["in/local"], arrays["in/masks"] = glob.numpy(), loc.numpy(), masks.numpy()

    for step in range(2):
        tag = f"s{step}/"
        for p in model.parameters():
            p.grad = None
        loss, loss_dict = model(batch, teacher_temp=cfg["teacher_temp"])
        loss.backward()
        arrays[tag + "loss"] = loss.detach().numpy()
        for k, v in loss_dict.items():
            arrays[tag + "loss/" + k] = v.detach().numpy()
        for n, p in model.student.named_parameters():
            if p.requires_grad:
                arrays[tag + "grad/" + n] = p.grad.numpy().copy()
        # the trainer's EMA of the teacher (models.py:437-447), then a plain SGD nudge of the student so that the second
        # step sees a different student, teacher and centres
        model.update_teacher(cfg["momentum"])
        with torch.no_grad():
            for n, p in model.student.named_parameters():
                if p.requires_grad:
                    p.add_(p.grad, alpha=-0.05)
        arrays[tag + "teacher_after/backbone.blocks.1.attn.proj_weight1"] = \
            model.teacher["backbone"].blocks[1].attn.proj_weight1.detach().numpy().copy()
        arrays[tag + "teacher_after/dino_head.mlp.0.weight"] = \
            model.teacher["dino_head"].mlp[0].weight.detach().numpy().copy()
    model.dino_loss.apply_center_update()
    model.ibot_patch_loss.apply_center_update()
    arrays["final/dino_center"] = model.dino_loss.center.numpy()
    arrays["final/ibot_center"] = model.ibot_patch_loss.center.numpy()

    np.savez_compressed(os.path.join(HERE, "ssl_step_tiny.npz"), **arrays)
    with open(os.path.join(HERE, "ssl_step_tiny.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(f"ssl_step_tiny: loss {float(loss):.6f}; {len(arrays)} arrays, "
          f"{os.path.getsize(os.path.join(HERE, 'ssl_step_tiny.npz')) / 1e3:.0f} kB; trainable {len(meta['trainable'])}")
    print({k: float(v) for k, v in loss_dict.items()})


if __name__ == "__main__":
    main()
