"""Golden vectors for the DINOv2 heads and losses, from the UNMODIFIED reference.

    python tests/golden/make_golden_ssl.py          # build container only (/root/reference must exist)

The four reference files are loaded BY PATH (importlib) because ``self_supervised/dinov2/__init__.py`` pulls in
``models.py``, which raises without xformers; the files themselves need only torch:
  layers/dino_head.py, loss/dino_clstoken_loss.py, loss/ibot_patch_loss.py (its own no-xformers ``lossfunc``),
  loss/koleo_loss.py.
Shapes are a scaled-down C4 step: 4 images -> 8 global-crop and 32 local-crop CLS tokens, 16 patches per global crop
of which some are masked, embed 64, head 64 -> 96 -> 96 -> 32 -> 256 prototypes.  Everything the oracle
(oracle/ssl_oracle.py) restates is recorded: head outputs, teacher targets, losses, gradients w.r.t. the student inputs
and the head parameters, centre updates over two consecutive steps.  Output: tests/golden/ssl_small.npz (+ .json)."""
import importlib.util
import json
import os
import random

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/self_supervised/dinov2"


def load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    head_m = load("layers/dino_head.py", "ref_dino_head")
    dino_m = load("loss/dino_clstoken_loss.py", "ref_dino_loss")
    ibot_m = load("loss/ibot_patch_loss.py", "ref_ibot_loss")
    koleo_m = load("loss/koleo_loss.py", "ref_koleo_loss")

    torch.manual_seed(0)
    random.seed(0)
    B, D, P, K = 4, 64, 16, 256
    n_global, n_local = 2, 8
    head = head_m.DINOHead(in_dim=D, out_dim=K, hidden_dim=96, bottleneck_dim=32, nlayers=3)
    with torch.no_grad():                       # make weight_g non-trivial (the reference trains it away from 1)
        head.last_layer.weight_g.mul_(1.0 + 0.1 * torch.randn_like(head.last_layer.weight_g))
        for p in head.mlp.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    teacher = head_m.DINOHead(in_dim=D, out_dim=K, hidden_dim=96, bottleneck_dim=32, nlayers=3)
    teacher.load_state_dict(head.state_dict())
    with torch.no_grad():
        for p in teacher.parameters():
            p.add_(0.01 * torch.randn_like(p))
    arrays, meta = {}, dict(B=B, D=D, P=P, K=K, n_global=n_global, n_local=n_local, student_temp=0.1,
                            center_momentum=0.9, teacher_temps=[0.04, 0.05])
    for k, v in head.state_dict().items():
        arrays["student_head/" + k] = v.detach().numpy()
    for k, v in teacher.state_dict().items():
        arrays["teacher_head/" + k] = v.detach().numpy()
    meta["head_keys"] = list(head.state_dict().keys())

    dino = dino_m.DINOLoss(out_dim=K)
    ibot = ibot_m.iBOTPatchLoss(patch_out_dim=K)
    koleo = koleo_m.KoLeoLoss()

    # masks the way collate_data_and_cast builds its outputs (dinov2_utils.py:43-48), random boolean grids here
    masks = torch.zeros(n_global * B, P, dtype=torch.bool)
    for i in range(0, n_global * B, 2):         # half of the crops carry a mask (mask_sample_probability 0.5)
        n = random.randint(2, 9)
        masks[i, torch.randperm(P)[:n]] = True
    idx = masks.flatten().nonzero().flatten()
    mw = (1 / masks.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(masks)[masks]
    arrays["masks"] = masks.numpy()
    arrays["mask_indices_list"] = idx.numpy()
    arrays["masks_weight"] = mw.numpy()
    n_masked = int(idx.shape[0])

    for step, temp in enumerate(meta["teacher_temps"]):
        tag = f"s{step}/"
        s_local = torch.randn(n_local * B, D, requires_grad=True)
        s_global = torch.randn(n_global * B, D, requires_grad=True)
        s_patch = torch.randn(n_global * B, P, D, requires_grad=True)
        t_cls = torch.randn(n_global * B, D)
        t_patch = torch.randn(n_global * B, P, D)
        for n_, t in dict(s_local=s_local, s_global=s_global, s_patch=s_patch, t_cls=t_cls, t_patch=t_patch).items():
            arrays[tag + "in/" + n_] = t.detach().numpy()

        # ---- teacher side (models.py:237-296, the shared-head branch) with the reference's own classes
        with torch.no_grad():
            a, b = t_cls.chunk(2)
            t_in = torch.cat((torch.cat((b, a)), t_patch.flatten(0, 1).index_select(0, idx)))
            t_out = teacher(t_in)
            t_cls_out, t_patch_out = t_out[:n_global * B], t_out[n_global * B:]
            arrays[tag + "teacher_head_out"] = t_out.numpy()
            t_dino = dino.softmax_center_teacher(t_cls_out, teacher_temp=temp)
            arrays[tag + "dino_center_used"] = dino.center.numpy().copy()
            dino.update_center(t_cls_out)
            t_ibot = ibot.softmax_center_teacher(t_patch_out.unsqueeze(0)[:, :n_masked], teacher_temp=temp).squeeze(0)
            arrays[tag + "ibot_center_used"] = ibot.center.numpy().copy()
            ibot.update_center(t_patch_out.unsqueeze(0)[:n_masked])
            arrays[tag + "t_dino"] = t_dino.numpy()
            arrays[tag + "t_ibot"] = t_ibot.numpy()

        # ---- student side: one head pass over the concatenation (models.py:335-371)
        s_in = torch.cat((s_local, s_global, s_patch.flatten(0, 1).index_select(0, idx)))
        s_out = head(s_in)
        arrays[tag + "student_head_out"] = s_out.detach().numpy()
        s_l, s_g, s_p = s_out[:n_local * B], s_out[n_local * B:(n_local + n_global) * B], s_out[(n_local + n_global) * B:]
        t_list = t_dino.view(n_global, -1, K)
        l_local = dino(student_output_list=s_l.chunk(n_local), teacher_out_softmaxed_centered_list=t_list)
        l_global = dino(student_output_list=[s_g], teacher_out_softmaxed_centered_list=[t_list.flatten(0, 1)])
        l_koleo = sum(koleo(p) for p in s_global.chunk(2))
        l_ibot = ibot.forward_masked(s_p, t_ibot, student_masks_flat=masks, n_masked_patches=n_masked, masks_weight=mw)
        l_ibot_default_w = ibot.forward_masked(s_p, t_ibot, student_masks_flat=masks)
        for n_, v in dict(dino_local=l_local, dino_global=l_global, koleo=l_koleo, ibot=l_ibot,
                          ibot_default_weight=l_ibot_default_w).items():
            arrays[tag + "loss/" + n_] = v.detach().numpy()
        # gradients of a fixed weighting of the four terms (the weights are not the model's: every term must show up)
        total = 0.3 * l_local + 0.7 * l_global + 0.1 * l_koleo + 0.5 * l_ibot
        head.zero_grad()
        total.backward()
        arrays[tag + "grad/s_local"] = s_local.grad.numpy()
        arrays[tag + "grad/s_global"] = s_global.grad.numpy()
        arrays[tag + "grad/s_patch"] = s_patch.grad.numpy()
        for k, p in head.named_parameters():
            arrays[tag + "grad/head/" + k] = p.grad.numpy().copy()
    # centres after the last pending update is applied
    dino.apply_center_update()
    ibot.apply_center_update()
    arrays["final/dino_center"] = dino.center.numpy()
    arrays["final/ibot_center"] = ibot.center.numpy()

    np.savez_compressed(os.path.join(HERE, "ssl_small.npz"), **arrays)
    with open(os.path.join(HERE, "ssl_small.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(f"ssl_small: {len(arrays)} arrays, {os.path.getsize(os.path.join(HERE, 'ssl_small.npz')) / 1e3:.0f} kB, "
          f"n_masked {n_masked}")


if __name__ == "__main__":
    main()
