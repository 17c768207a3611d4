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
Those components are therefore *pinned*.  The whole step -- ``ssl_step``: teacher backbone on the unmasked global crops,
student backbone over the packed [masked global | local] crop list, ``ssl_objective``, ``ema_update`` -- is pinned against
``DINOv2.forward`` / ``update_teacher`` of the UNMODIFIED reference run on the CPU for two steps
(``tests/golden/make_golden_ssl_step.py`` -> ``ssl_step_tiny``; loss, the four loss terms, every trainable gradient, the
teacher after the EMA and both centres, <= 2e-5).  The reference's ``models.py`` cannot be imported without xformers, so
that generator supplies ``tests/golden/xformers_shim.py`` (plain-torch ``memory_efficient_attention``, ``unbind``,
``BlockDiagonalMask``): **pinned modulo the xformers shim** -- the kernels of xformers 0.0.18 themselves and its
``cross_entropy`` (the reference's own fallback, ibot_patch_loss.py:26-27, is what ran) are the unpinned residue.
Sinkhorn-Knopp centring (``sinkhorn_knopp_teacher``, dino_clstoken_loss.py:33-60, ibot_patch_loss.py:53-83; no shipped config
selects it) is restated as ``sinkhorn_knopp`` and pinned by ``tests/golden/make_golden_ssl_sk.py`` -> ``ssl_sk_small``.
"""
from __future__ import annotations

import math
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


def sinkhorn_knopp(teacher_out: Tensor, teacher_temp: float, n_iterations: int = 3, n_samples_world: Optional[int] = None,
                   all_reduce=None) -> Tensor:
    """dino_clstoken_loss.py:33-60 / ibot_patch_loss.py:53-83, in the [samples, K] layout (the reference works on the
    transpose): Q = exp(t / temp) normalised to total mass 1, then n_iterations of {each prototype's mass -> 1/K, each
    sample's mass -> 1/B}, finally x B so that every sample's row sums to 1.  ``n_samples_world`` = B (samples over all
    ranks; iBOT passes the all-reduced number of masked patches), ``all_reduce`` = in-place sum over ranks (None = 1 rank)."""
    red = all_reduce if all_reduce is not None else (lambda x: x)
    Q = torch.exp(teacher_out.float() / teacher_temp)
    B = Q.shape[0] if n_samples_world is None else n_samples_world
    K = Q.shape[1]
    Q = Q / red(Q.sum())
    for _ in range(n_iterations):
        Q = Q / red(Q.sum(dim=0, keepdim=True))
        Q = Q / K
        Q = Q / Q.sum(dim=1, keepdim=True)
        Q = Q / B
    return Q * B


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
    """Pinned through ``ssl_step`` (see the module docstring).  Inputs are backbone outputs after the final norm:
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


# ---------------------------------------------------------------------------------------------------------------------
# the backbone as the SSL step drives it (DinoVisionTransformer, dinov2_vits.py) and the whole step
# ---------------------------------------------------------------------------------------------------------------------
def dinov2_pos_embed(pos_embed: Tensor, w: int, h: int, patch: int, interpolate_offset: float = 0.1,
                     antialias: bool = False) -> Tensor:
    """dinov2_vits.py:176-208: bicubic resize of the patch grid of the table to (w // patch, h // patch); with the
    shipped ``interpolate_offset`` 0.1 the resize is driven by scale factors (w0 + 0.1) / M, not by the output size."""
    N = pos_embed.shape[1] - 1
    w0, h0 = w // patch, h // patch
    if w0 * h0 == N and w == h:
        return pos_embed
    M = int(math.sqrt(N))
    assert N == M * M
    dim = pos_embed.shape[-1]
    grid = pos_embed[:, 1:].float().reshape(1, M, M, dim).permute(0, 3, 1, 2)
    if interpolate_offset:
        kw = dict(scale_factor=(float(w0 + interpolate_offset) / M, float(h0 + interpolate_offset) / M))
    else:
        kw = dict(size=(w0, h0))
    grid = F.interpolate(grid, mode="bicubic", antialias=antialias, **kw)
    assert (w0, h0) == tuple(grid.shape[-2:])
    return torch.cat((pos_embed[:, :1].float(), grid.permute(0, 2, 3, 1).reshape(1, -1, dim)), dim=1)


def dinov2_prepare_tokens(sd: Dict[str, Tensor], images: Tensor, masks: Optional[Tensor], patch: int,
                          pre: str = "backbone.") -> Tensor:
    """dinov2_vits.py:210-231 (no register tokens in any shipped config): patch embedding, masked patches replaced by
    ``mask_token`` BEFORE the position table is added, CLS token prepended."""
    _, _, w, h = images.shape
    x = F.conv2d(images, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)
    if masks is not None:
        x = torch.where(masks.unsqueeze(-1), sd[pre + "mask_token"].to(x.dtype).unsqueeze(0), x)
    x = torch.cat((sd[pre + "cls_token"].expand(x.shape[0], -1, -1), x), dim=1)
    return x + dinov2_pos_embed(sd[pre + "pos_embed"], w, h, patch)


def dinov2_backbone(sd: Dict[str, Tensor], crops, masks, *, patch: int, depth: int, num_heads: int,
                    eps: float = 1e-6, pre: str = "backbone."):
    """``DinoVisionTransformer.forward_features`` (dinov2_vits.py:269-289) for one tensor of crops, or
    ``forward_features_list`` (:249-267) for a list: every block then runs ONCE over the crops of all resolutions packed
    into one [1, sum b_i n_i, D] sequence under a block-diagonal mask (layers/block.py:244-288 -- attention per original
    sequence, everything else token-wise), and the final LayerNorm is applied per resolution.
    -> dict(cls [b, D], patch [b, P, D]) or a list of them."""
    from oracle.apla_oracle import block_forward
    is_list = isinstance(crops, (list, tuple))
    xs = [dinov2_prepare_tokens(sd, c, m, patch, pre) for c, m in zip(crops, masks)] if is_list \
        else [dinov2_prepare_tokens(sd, crops, masks, patch, pre)]
    D = xs[0].shape[-1]
    if is_list:
        seqlens = [x.shape[1] for x in xs for _ in range(x.shape[0])]
        packed = torch.cat([x.reshape(1, -1, D) for x in xs], dim=1)
        for i in range(depth):
            packed = block_forward(sd, f"{pre}blocks.{i}.", packed, num_heads, eps, seqlens=seqlens)
        outs, o = [], 0
        for x in xs:
            n = x.shape[0] * x.shape[1]
            outs.append(packed[:, o:o + n].reshape(x.shape))
            o += n
    else:
        x = xs[0]
        for i in range(depth):
            x = block_forward(sd, f"{pre}blocks.{i}.", x, num_heads, eps)
        outs = [x]
    res = []
    for x in outs:
        xn = F.layer_norm(x, (D,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], eps)
        res.append(dict(cls=xn[:, 0], patch=xn[:, 1:]))
    return res if is_list else res[0]


def ssl_step(student: Dict[str, Tensor], teacher: Dict[str, Tensor], global_crops: Tensor, local_crops: Tensor,
             masks: Tensor, dino_center: Tensor, ibot_center: Tensor, *, teacher_temp: float, patch: int, depth: int,
             num_heads: int, n_local_crops: int = 8, n_global_crops: int = 2, dino_loss_weight: float = 1.0,
             koleo_loss_weight: float = 0.1, ibot_loss_weight: float = 1.0, center_momentum: float = 0.9,
             world_size: int = 1):
    """``DINOv2.forward`` end to end (models.py:207-433): the teacher backbone sees the global crops UNMASKED as one
    plain batch (:240), the student backbone sees [global crops with masks, local crops] as a packed list (:322-324);
    then ``ssl_objective``.  ``student`` / ``teacher`` are flat state dicts (``backbone.*``, ``dino_head.*``).
    PINNED against the unmodified reference modulo the xformers shim (tests/golden/make_golden_ssl_step.py).
    -> (loss, loss_dict, (dino_center, ibot_center) for the next step)."""
    kw = dict(patch=patch, depth=depth, num_heads=num_heads)
    with torch.no_grad():
        t = dinov2_backbone(teacher, global_crops, None, **kw)
    s_glob, s_loc = dinov2_backbone(student, [global_crops, local_crops], [masks, None], **kw)
    head = lambda sd: {k[len("dino_head."):]: v for k, v in sd.items() if k.startswith("dino_head.")}  # noqa: E731
    return ssl_objective(head(student), head(teacher), student_local_cls=s_loc["cls"], student_global_cls=s_glob["cls"],
                         student_global_patch=s_glob["patch"], teacher_global_cls=t["cls"],
                         teacher_global_patch=t["patch"], masks=masks, dino_center=dino_center, ibot_center=ibot_center,
                         teacher_temp=teacher_temp, n_local_crops=n_local_crops, n_global_crops=n_global_crops,
                         dino_loss_weight=dino_loss_weight, koleo_loss_weight=koleo_loss_weight,
                         ibot_loss_weight=ibot_loss_weight, center_momentum=center_momentum, world_size=world_size)
