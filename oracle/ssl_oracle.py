"""CPU oracle for the DINOv2 self-supervised heads and losses  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

SURVEY.md 8(f) row f2: what sits on top of the block path in BASELINE config C4 (ViT-L/14 student / teacher over
multi-crop batches).  No CUDA path exists for these yet; this file is the first step of that row -- the restatement the
kernels of a later round will be held to -- and, like ``oracle/apla_oracle.py``, may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs.

Plain fp32 PyTorch-on-CPU restatement (functional, weights in flat dicts keyed like the reference state dict) of

* ``DINOHead.forward``                    src/self_supervised/dinov2/layers/dino_head.py:12-58
* ``DINOLoss`` (softmax-center teacher, cross-entropy over crop pairs, centre EMA)
                                          src/self_supervised/dinov2/loss/dino_clstoken_loss.py:12-98
* ``iBOTPatchLoss.forward_masked`` + centre EMA
                                          src/self_supervised/dinov2/loss/ibot_patch_loss.py:22-145
* ``KoLeoLoss``                           src/self_supervised/dinov2/loss/koleo_loss.py:17-45
* the loss assembly of ``DINOv2.forward`` and ``update_teacher``
                                          src/self_supervised/dinov2/models.py:212-447
* the mask bookkeeping of ``collate_data_and_cast``
                                          src/self_supervised/dinov2/dinov2_utils.py:21-62

Pinning: ``tests/golden/make_golden_ssl.py`` loads the four reference files above BY PATH (their package ``__init__``
imports xformers, which this image lacks) and records head outputs, teacher targets, the three losses, their
gradients and the centre updates; ``tests/test_ssl_oracle.py`` replays them against this file (<= 1e-5 relative).
Those components are therefore *pinned*.  ``ssl_objective`` (the assembly in ``models.py``, which cannot be imported
without xformers) is restated from the source with every scale cited -- **assembly parity unpinned**; so is the
xformers ``cross_entropy`` the reference prefers for the iBOT term when xformers is present (the fallback it defines
itself, ibot_patch_loss.py:26-27, is what is pinned; the two are the same function of their inputs).
Sinkhorn-Knopp centring (dino_clstoken_loss.py:33-60) is not restated: every shipped config uses "centering".
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---------------------------------------------------------------------------------------------------------------------
# DINO head
# ---------------------------------------------------------------------------------------------------------------------
def weight_norm_weight(g: Tensor, v: Tensor) -> Tensor:
    """torch.nn.utils.weight_norm(dim=0) of a Linear: W[o, :] = g[o] * v[o, :] / ||v[o, :]||  (dino_head.py:27)."""
    return v * (g / v.norm(dim=1, keepdim=True))


def dino_head_forward(sd: Dict[str, Tensor], x: Tensor, prefix: str = "") -> Tensor:
    """dino_head.py:36-41 with the MLP of ``_build_mlp`` (:44-58; Linear / GELU pairs, ``use_bn`` False in every shipped
    config): x [..., in_dim] -> [..., out_dim] prototype scores.  ``sd`` holds ``mlp.{0,2,...}.weight/bias`` (or
    ``mlp.weight/bias`` for nlayers == 1) and ``last_layer.weight_g`` [out,1] / ``last_layer.weight_v`` [out, bottleneck]."""
    if prefix + "mlp.weight" in sd:                                  # nlayers == 1 (:45-46)
        x = F.linear(x, sd[prefix + "mlp.weight"], sd.get(prefix + "mlp.bias"))
    else:
        idx = sorted(int(k[len(prefix) + 4:].split(".")[0]) for k in sd
                     if k.startswith(prefix + "mlp.") and k.endswith(".weight"))
        for j, i in enumerate(idx):
            x = F.linear(x, sd[f"{prefix}mlp.{i}.weight"], sd.get(f"{prefix}mlp.{i}.bias"))
            if j + 1 < len(idx):
                x = F.gelu(x)                                        # exact GELU (:51,56)
    eps = 1e-6 if x.dtype == torch.float16 else 1e-12               # :38
    x = F.normalize(x, dim=-1, p=2, eps=eps)
    w = weight_norm_weight(sd[prefix + "last_layer.weight_g"], sd[prefix + "last_layer.weight_v"])
    return F.linear(x, w)


# ---------------------------------------------------------------------------------------------------------------------
# teacher targets and centres
# ---------------------------------------------------------------------------------------------------------------------
def softmax_center_teacher(teacher_out: Tensor, center: Tensor, teacher_temp: float) -> Tensor:
    """dino_clstoken_loss.py:28-31 / ibot_patch_loss.py:39-51 (after the pending centre update has been applied)."""
    return F.softmax((teacher_out - center) / teacher_temp, dim=-1)


def dino_center_update(center: Tensor, teacher_out: Tensor, momentum: float = 0.9, world_size: int = 1,
                       summed_over_ranks: Optional[Tensor] = None) -> Tensor:
    """dino_clstoken_loss.py:76-98: centre <- centre * m + mean_over_all_ranks(teacher_out) * (1 - m).
    teacher_out [n, K]; ``summed_over_ranks`` stands for the all-reduced row sum when world_size > 1."""
    s = teacher_out.sum(dim=0, keepdim=True) if summed_over_ranks is None else summed_over_ranks
    return center * momentum + s / (len(teacher_out) * world_size) * (1 - momentum)


def ibot_center_update(center: Tensor, teacher_patch_out: Tensor, momentum: float = 0.9, world_size: int = 1) -> Tensor:
    """ibot_patch_loss.py:123-145: the batch statistic is sum over dim 0 of the mean over dim 1, divided by
    len(tensor) * world.  teacher_patch_out [b, n, K] (the reference passes [1, n_masked, K], models.py:296)."""
    s = teacher_patch_out.mean(1).sum(dim=0, keepdim=True)
    return center * momentum + s / (len(teacher_patch_out) * world_size) * (1 - momentum)


# ---------------------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------------------
def dino_loss(student_output_list: Sequence[Tensor], teacher_softmaxed_centered_list: Sequence[Tensor],
              student_temp: float = 0.1) -> Tensor:
    """dino_clstoken_loss.py:62-74: - sum over (student crop, teacher crop) pairs of mean_b sum_k t * log_softmax(s / T)."""
    total = torch.zeros((), dtype=torch.float32)
    for s in student_output_list:
        lsm = F.log_softmax(s / student_temp, dim=-1)
        for t in teacher_softmaxed_centered_list:
            total = total - torch.sum(t * lsm, dim=-1).mean()
    return total


def ibot_loss_masked(student_masked: Tensor, teacher_masked: Tensor, student_masks_flat: Tensor,
                     n_masked_patches: Optional[int] = None, masks_weight: Optional[Tensor] = None,
                     student_temp: float = 0.1) -> Tensor:
    """ibot_patch_loss.py:102-121 (``lossfunc`` = the file's own fallback :26-27)."""
    loss = torch.sum(teacher_masked * F.log_softmax(student_masked / student_temp, dim=-1), dim=-1)
    if masks_weight is None:
        masks_weight = masks_weight_of(student_masks_flat)
    if n_masked_patches is not None:
        loss = loss[:n_masked_patches]
    loss = loss * masks_weight
    return -loss.sum() / student_masks_flat.shape[0]


def koleo_loss(student_output: Tensor, eps: float = 1e-8) -> Tensor:
    """koleo_loss.py:23-45: -mean log distance to the nearest neighbour among the L2-normalised rows
    (nn.PairwiseDistance(2, eps=1e-8): ||a - b + eps||_2)."""
    x = F.normalize(student_output.float(), eps=eps, p=2, dim=-1)
    dots = x @ x.t()
    n = x.shape[0]
    dots = dots.masked_fill(torch.eye(n, dtype=torch.bool), -1.0)    # :30 fills the diagonal with -1
    nn_idx = dots.argmax(dim=1)
    dist = torch.linalg.vector_norm(x - x[nn_idx] + 1e-8, ord=2, dim=-1)
    return -torch.log(dist + eps).mean()


# ---------------------------------------------------------------------------------------------------------------------
# mask bookkeeping (collate_data_and_cast)
# ---------------------------------------------------------------------------------------------------------------------
def mask_indices_of(collated_masks: Tensor) -> Tensor:
    """dinov2_utils.py:46: positions of the masked patches in the flattened [2B * P] global-crop patch grid."""
    return collated_masks.flatten().nonzero().flatten()


def masks_weight_of(collated_masks: Tensor) -> Tensor:
    """dinov2_utils.py:48: every masked patch weighs 1 / (masked patches of its crop)."""
    return (1 / collated_masks.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(collated_masks)[collated_masks]


# ---------------------------------------------------------------------------------------------------------------------
# the objective of one step (models.py:212-433), shared dino / ibot head, "centering"
# ---------------------------------------------------------------------------------------------------------------------
def ssl_objective(student_head: Dict[str, Tensor], teacher_head: Dict[str, Tensor], *, student_local_cls: Tensor,
                  student_global_cls: Tensor, student_global_patch: Tensor, teacher_global_cls: Tensor,
                  teacher_global_patch: Tensor, masks: Tensor, dino_center: Tensor, ibot_center: Tensor,
                  teacher_temp: float, n_local_crops: int = 8, n_global_crops: int = 2, dino_loss_weight: float = 1.0,
                  koleo_loss_weight: float = 0.1, ibot_loss_weight: float = 1.0, student_temp: float = 0.1,
                  center_momentum: float = 0.9, world_size: int = 1):
    """ASSEMBLY PARITY UNPINNED (see the module docstring).  Inputs are backbone outputs after the final norm:
    student_local_cls [n_local*B, D], student_global_cls [2B, D], student_global_patch [2B, P, D], the teacher's
    [2B, D] / [2B, P, D] on the same global crops, masks bool [2B, P].  -> (loss, loss_dict, new centres)."""
    assert n_global_crops == 2                                                            # :214
    idx = mask_indices_of(masks)
    n_masked = idx.shape[0]
    mw = masks_weight_of(masks)
    n_local_terms = max(n_local_crops * n_global_crops, 1)                                # :227
    n_global_terms = (n_global_crops - 1) * n_global_crops                                # :228
    ibot_loss_scale = 1.0 / n_global_crops                                                # :234

    # ---- teacher (:237-318): CLS tokens of the two global crops swapped so that crop A is matched to crop B
    with torch.no_grad():
        a, b = teacher_global_cls.chunk(n_global_crops)
        t_cls = torch.cat((b, a))
        n_cls = t_cls.shape[0]
        t_in = torch.cat((t_cls, teacher_global_patch.flatten(0, 1).index_select(0, idx)))   # :249-257 minus the padding
        t_out = dino_head_forward(teacher_head, t_in)
        t_cls_out, t_patch_out = t_out[:n_cls], t_out[n_cls:n_cls + n_masked]
        # softmax_center_teacher first applies the update left pending by the PREVIOUS step (:30,41); callers pass the
        # centres in that applied state and get back the ones to use at the next step
        t_dino = softmax_center_teacher(t_cls_out, dino_center, teacher_temp).view(n_global_crops, -1, t_cls_out.shape[-1])
        new_dino_center = dino_center_update(dino_center, t_cls_out, center_momentum, world_size)           # :288
        t_ibot = softmax_center_teacher(t_patch_out.unsqueeze(0), ibot_center, teacher_temp).squeeze(0)    # :290-295
        new_ibot_center = ibot_center_update(ibot_center, t_patch_out.unsqueeze(0), center_momentum, world_size)  # :296

    # ---- student head over [local cls | global cls | masked global patches] in one pass (:335-371)
    s_patch_in = student_global_patch.flatten(0, 1).index_select(0, idx)
    s_out = dino_head_forward(student_head, torch.cat((student_local_cls, student_global_cls, s_patch_in)))
    n_l, n_g = student_local_cls.shape[0], student_global_cls.shape[0]
    s_local, s_global, s_patch = s_out[:n_l], s_out[n_l:n_l + n_g], s_out[n_l + n_g:]

    losses = {}
    total = torch.zeros((), dtype=torch.float32)
    if n_local_crops > 0:                                                                 # :374-386
        l = dino_loss(s_local.chunk(n_local_crops), list(t_dino), student_temp) / (n_global_terms + n_local_terms)
        losses["dino_local_crops_loss"] = l
        total = total + dino_loss_weight * l
    loss_scales = 2                                                                       # :389
    g = dino_loss([s_global], [t_dino.flatten(0, 1)], student_temp) * loss_scales / (n_global_terms + n_local_terms)
    losses["dino_global_crops_loss"] = g                                                  # :392-404
    total = total + dino_loss_weight * g
    if koleo_loss_weight > 0:                                                             # :412-420
        k = koleo_loss_weight * sum(koleo_loss(p) for p in student_global_cls.chunk(2))
        losses["koleo_loss"] = k / loss_scales
        total = total + k
    i = ibot_loss_masked(s_patch, t_ibot, masks, n_masked_patches=n_masked, masks_weight=mw,
                         student_temp=student_temp) * loss_scales * ibot_loss_scale      # :423-433
    losses["ibot_loss"] = i / 2
    total = total + ibot_loss_weight * i
    return total, losses, (new_dino_center, new_ibot_center)


def ema_update(teacher: Dict[str, Tensor], student: Dict[str, Tensor], m: float) -> None:
    """models.py:437-447 ``update_teacher``: teacher <- m * teacher + (1 - m) * student, parameter by parameter, in place."""
    with torch.no_grad():
        for k, t in teacher.items():
            if t.is_floating_point() and k in student:
                t.mul_(m).add_(student[k].detach(), alpha=1 - m)
