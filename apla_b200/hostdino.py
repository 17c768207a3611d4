"""Host side of one DINOv2 self-supervised step (BASELINE config C4) for tests, smoke and bench: stand-ins for the
reference's `DinoVisionTransformer` (src/self_supervised/dinov2/dinov2_vits.py:60-330) and for the `DINOv2`
meta-architecture (src/self_supervised/dinov2/models.py:29-453), which do not exist on the GPU box.  Like
`apla_b200/hostvit.py` this is NOT accelerated code: it is the model the accelerated pieces are dropped into --
`apla_b200.apla` (APLA attention / `fuse_apla_blocks`: the packed multi-crop block path), `apla_b200.dinov2.DINOHead`
and the three loss classes.  All step arithmetic happens in those.

Contract kept with the reference so that its state dicts and the golden vectors recorded from it interchange:
  * backbone keys `patch_embed.proj`, `cls_token`, `pos_embed`, `mask_token`, `blocks.N.*`, `norm` (block_chunks = 0,
    no register tokens -- every shipped config); meta-architecture keys `student.backbone.*`, `student.dino_head.*`,
    `teacher.*` (shared DINO / iBOT head, `ibot.separate_head = False`, "centering" -- every shipped config);
  * `forward_features(x, masks)` / `forward_features_list` return the reference's dict keys;
  * position table resized with the 0.1-offset scale factors of dinov2_vits.py:176-208, masked patches replaced by
    `mask_token` before the table is added (:210-231);
  * `SSLMetaArch.forward(batch, teacher_temp)` takes the dict `collate_data_and_cast` produces (dinov2_utils.py:21-62)
    and returns `(loss, loss_dict)` with the reference's four keys and scales (models.py:374-433).
NOT claimed: the random-number order of the reference's constructor (tests fill the weights explicitly)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .hostvit import ARCHS, HostViT, VitArch


class HostDinoViT(HostViT):
    def __init__(self, arch: VitArch, img_size: int = 518, patch_size: int = 14, layerscale: Optional[float] = 1.0,
                 eps: float = 1e-6, interpolate_offset: float = 0.1, interpolate_antialias: bool = False):
        super().__init__(arch, img_size=img_size, patch_size=patch_size, layerscale=layerscale, eps=eps)
        self.patch_size = patch_size
        self.num_register_tokens = 0
        self.interpolate_offset = interpolate_offset
        self.interpolate_antialias = interpolate_antialias
        self.mask_token = nn.Parameter(torch.zeros(1, arch.embed_dim))
        self.head = nn.Identity()

    def interpolate_pos_encoding(self, x, w, h):
        npatch, N = x.shape[1] - 1, self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        # a frozen table resized to the same grid gives the same result every forward (the reference recomputes the
        # bicubic resize each time, dinov2_vits.py:176-208: 1.7 ms per crop size at ViT-L): cached per (table, grid, dtype)
        if not self.pos_embed.requires_grad:
            key = (self.pos_embed.data_ptr(), self.pos_embed._version, str(self.pos_embed.device), w, h, x.dtype)
            hit = getattr(self, "_pos_cache", {}).get(key)
            if hit is None:
                hit = self._interpolate_pos_encoding(x, w, h).detach()
                cache = getattr(self, "_pos_cache", {})
                if len(cache) > 8:
                    cache.clear()
                cache[key] = hit
                self._pos_cache = cache
            return hit
        return self._interpolate_pos_encoding(x, w, h)

    def _interpolate_pos_encoding(self, x, w, h):
        npatch, N = x.shape[1] - 1, self.pos_embed.shape[1] - 1
        table = self.pos_embed.float()
        D = x.shape[-1]
        w0, h0 = w // self.patch_size, h // self.patch_size
        M = int(math.sqrt(N))
        if self.interpolate_offset:
            kw = dict(scale_factor=(float(w0 + self.interpolate_offset) / M, float(h0 + self.interpolate_offset) / M))
        else:
            kw = dict(size=(w0, h0))
        grid = F.interpolate(table[:, 1:].reshape(1, M, M, D).permute(0, 3, 1, 2), mode="bicubic",
                             antialias=self.interpolate_antialias, **kw)
        if (w0, h0) != tuple(grid.shape[-2:]):
            raise AssertionError("position grid does not match the patch grid")
        return torch.cat((table[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, D)), dim=1).to(x.dtype)

    def prepare_tokens_with_masks(self, x, masks=None):
        _, _, w, h = x.shape
        x = self.patch_embed(x)
        if masks is not None:
            x = torch.where(masks.unsqueeze(-1), self.mask_token.to(x.dtype).unsqueeze(0), x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1)
        return x + self.interpolate_pos_encoding(x, w, h)

    @staticmethod
    def _out(x_norm, x, masks):
        return {"x_norm_clstoken": x_norm[:, 0], "x_norm_regtokens": x_norm[:, 1:1], "x_norm_patchtokens": x_norm[:, 1:],
                "x_prenorm": x, "masks": masks}

    def forward_features_list(self, x_list, masks_list):
        xs = [self.prepare_tokens_with_masks(x, m) for x, m in zip(x_list, masks_list)]
        for blk in self.blocks:
            if getattr(blk, "accepts_crop_lists", False) or type(blk).__name__ == "FusedAplaBlock":
                xs = blk(xs)                      # one packed, block-diagonal pass (dinov2/layers/block.py:244-288)
            else:
                xs = [blk(x) for x in xs]         # same numbers: every op but attention is token-wise
        return [self._out(self.norm(x), x, m) for x, m in zip(xs, masks_list)]

    def forward_features(self, x, masks=None):
        if isinstance(x, (list, tuple)):
            return self.forward_features_list(x, masks)
        x = self.prepare_tokens_with_masks(x, masks)
        for blk in self.blocks:
            x = blk(x)
        return self._out(self.norm(x), x, masks)

    def forward(self, *args, is_training=False, **kwargs):
        ret = self.forward_features(*args, **kwargs)
        return ret if is_training else self.head(ret["x_norm_clstoken"])


def build_dino_backbone(arch, *, img_size: int, patch_size: int, apla_config, attn_class: str = "apla_attn_mem_eff",
                        layerscale: Optional[float] = 1.0, fuse: bool = True,
                        indices: Optional[Sequence[torch.Tensor]] = None) -> HostDinoViT:
    """backbone -> build_apla (freeze policy: proj_weight1 / proj_bias1 trainable) -> fused blocks.
    `indices` (one full permutation of range(dim) per block) plays the part of the reference's `inds_path` file
    (src/apla/apla_vit.py:20-24) for callers that hold the permutations in memory, e.g. a loaded state dict's `inds`."""
    from .apla.apla_block import fuse_apla_blocks
    from .apla.apla_vit import build_apla
    a = ARCHS[arch] if isinstance(arch, str) else arch
    vit = HostDinoViT(a, img_size=img_size, patch_size=patch_size, layerscale=layerscale)
    vit = build_apla(apla_config, vit, attn_class)
    if indices is not None:
        if len(indices) != len(vit.blocks):
            raise ValueError("one index permutation per block expected")
        for blk, inds in zip(vit.blocks, indices):
            old = blk.attn
            inds = torch.as_tensor(inds, dtype=torch.long).cpu()
            if sorted(inds.tolist()) != list(range(old.dim)):
                raise ValueError("indices must be a permutation of range(dim)")
            new = type(old)(config=apla_config, dim=old.dim, indices=inds, num_heads=old.num_heads,
                            qkv_bias=old.qkv.bias is not None, qk_scale=old.scale, attn_drop=old.attn_drop.p,
                            proj_drop=old.proj_drop.p)
            with torch.no_grad():                      # same content, re-split by the given permutation
                new.qkv.load_state_dict(old.qkv.state_dict())
                w = torch.empty(old.dim, old.dim)
                b = torch.empty(old.dim)
                w[old.trainable_inds], w[old.freezed_inds] = old.proj_weight1.data, old.proj_weight2.data
                b[old.trainable_inds], b[old.freezed_inds] = old.proj_bias1.data, old.proj_bias2.data
                new.proj_weight1.data, new.proj_weight2.data = w[new.trainable_inds], w[new.freezed_inds]
                new.proj_bias1.data, new.proj_bias2.data = b[new.trainable_inds], b[new.freezed_inds]
            blk.attn = new
    if fuse:
        from .apla.patch_embed import fuse_patch_embed
        fuse_patch_embed(fuse_apla_blocks(vit))       # frozen stem: patch extraction + tcgen05 GEMM instead of cuDNN
    return vit


class SSLMetaArch(nn.Module):
    """Student / teacher pair and the objective of one step, wired as DINOv2.forward wires them (models.py:207-433)."""

    def __init__(self, student_backbone: nn.Module, teacher_backbone: nn.Module, student_head: nn.Module,
                 teacher_head: nn.Module, out_dim: int, *, n_global_crops: int = 2, n_local_crops: int = 8,
                 dino_loss_weight: float = 1.0, koleo_loss_weight: float = 0.1, ibot_loss_weight: float = 1.0,
                 loss_classes=None, fused_objective: bool = False, centering: str = "centering"):
        super().__init__()
        if n_global_crops != 2:
            raise AssertionError("the objective is written for two global crops (models.py:214)")
        if loss_classes is None:
            from .dinov2 import DINOLoss, KoLeoLoss, iBOTPatchLoss
            loss_classes = (DINOLoss, iBOTPatchLoss, KoLeoLoss)
        self.student = nn.ModuleDict(dict(backbone=student_backbone, dino_head=student_head))
        self.teacher = nn.ModuleDict(dict(backbone=teacher_backbone, dino_head=teacher_head))
        for p in self.teacher.parameters():
            p.requires_grad = False                                                      # models.py:140-141
        self.dino_loss = loss_classes[0](out_dim)
        self.ibot_patch_loss = loss_classes[1](out_dim)
        self.koleo_loss = loss_classes[2]()
        self.n_global_crops, self.n_local_crops = n_global_crops, n_local_crops
        self.dino_loss_weight, self.koleo_loss_weight = dino_loss_weight, koleo_loss_weight
        self.ibot_loss_weight = ibot_loss_weight
        # True: student head + the three cross-entropy terms as ONE autograd node over `apla_ssl_objective`
        # (DINOHead.forward_with_objective); False: the reference's call structure, one loss-class call per term
        self.fused_objective = fused_objective
        if centering not in ("centering", "sinkhorn_knopp"):
            raise NotImplementedError(centering)                                         # models.py:317-318
        if centering == "sinkhorn_knopp" and fused_objective:
            raise ValueError("the fused objective implements softmax-centre targets; use the per-term form with "
                             "centering='sinkhorn_knopp'")
        self.centering = centering

    def forward(self, images: Dict[str, torch.Tensor], teacher_temp: float):
        dev = next(self.student.parameters()).device
        ng, nl = self.n_global_crops, self.n_local_crops
        global_crops = images["collated_global_crops"].to(dev, non_blocking=True)
        local_crops = images["collated_local_crops"].to(dev, non_blocking=True)
        masks = images["collated_masks"].to(dev, non_blocking=True)
        mask_indices = images["mask_indices_list"].to(dev, non_blocking=True)
        masks_weight = images["masks_weight"].to(dev, non_blocking=True)
        n_masked = mask_indices.shape[0]
        n_local_terms = max(nl * ng, 1)                                                  # models.py:227-228
        n_global_terms = (ng - 1) * ng
        ibot_loss_scale = 1.0 / ng                                                       # :234

        # ---- teacher (:237-318): unmasked global crops, CLS rows of the two crops swapped, head over [cls | masked]
        with torch.no_grad():
            t = self.teacher["backbone"](global_crops, is_training=True)
            a, b = t["x_norm_clstoken"].chunk(ng)
            t_cls = torch.cat((b, a))
            n_cls = t_cls.shape[0]
            t_patch = t["x_norm_patchtokens"].flatten(0, 1).index_select(0, mask_indices)
            t_out = self.teacher["dino_head"](torch.cat((t_cls, t_patch)))
            t_cls_out, t_patch_out = t_out[:n_cls], t_out[n_cls:n_cls + n_masked]
        if self.fused_objective:
            return self._forward_fused(t_out, global_crops, local_crops, masks, mask_indices, masks_weight, teacher_temp)
        with torch.no_grad():
            if self.centering == "sinkhorn_knopp":                                        # models.py:303-315
                t_dino = self.dino_loss.sinkhorn_knopp_teacher(t_cls_out, teacher_temp=teacher_temp) \
                    .view(ng, -1, t_cls_out.shape[-1])
                t_ibot = self.ibot_patch_loss.sinkhorn_knopp_teacher(
                    t_patch_out, teacher_temp=teacher_temp,
                    n_masked_patches_tensor=images["n_masked_patches"].to(dev).clone())
            else:                                                                         # "centering", models.py:288-301
                t_dino = self.dino_loss.softmax_center_teacher(t_cls_out, teacher_temp=teacher_temp) \
                    .view(ng, -1, t_cls_out.shape[-1])
                self.dino_loss.update_center(t_cls_out)
                t_ibot = self.ibot_patch_loss.softmax_center_teacher(t_patch_out.unsqueeze(0),
                                                                     teacher_temp=teacher_temp).squeeze(0)
                self.ibot_patch_loss.update_center(t_patch_out.unsqueeze(0))

        # ---- student (:322-371): [masked global | local] crops in one packed pass, one head pass over all rows
        s_glob, s_loc = self.student["backbone"]([global_crops, local_crops], masks=[masks, None], is_training=True)
        s_patch = s_glob["x_norm_patchtokens"].flatten(0, 1).index_select(0, mask_indices)
        n_l, n_g = s_loc["x_norm_clstoken"].shape[0], s_glob["x_norm_clstoken"].shape[0]
        s_out = self.student["dino_head"](torch.cat((s_loc["x_norm_clstoken"], s_glob["x_norm_clstoken"], s_patch)))
        s_local, s_global, s_masked = s_out[:n_l], s_out[n_l:n_l + n_g], s_out[n_l + n_g:]

        loss_dict, total = {}, 0
        denom = n_global_terms + n_local_terms
        if nl > 0:                                                                       # :374-386
            l = self.dino_loss(s_local.chunk(nl), list(t_dino)) / denom
            loss_dict["dino_local_crops_loss"] = l
            total = total + self.dino_loss_weight * l
        loss_scales = 2                                                                  # :389
        g = self.dino_loss([s_global], [t_dino.flatten(0, 1)]) * loss_scales / denom      # :392-404
        loss_dict["dino_global_crops_loss"] = g
        total = total + self.dino_loss_weight * g
        if self.koleo_loss_weight > 0:                                                   # :412-420
            k = self.koleo_loss_weight * sum(self.koleo_loss(p) for p in s_glob["x_norm_clstoken"].chunk(2))
            loss_dict["koleo_loss"] = k / loss_scales
            total = total + k
        i = self.ibot_patch_loss.forward_masked(s_masked, t_ibot, student_masks_flat=masks, n_masked_patches=n_masked,
                                                masks_weight=masks_weight) * loss_scales * ibot_loss_scale   # :423-433
        loss_dict["ibot_loss"] = i / 2
        total = total + self.ibot_loss_weight * i
        return total, loss_dict

    def _forward_fused(self, t_out, global_crops, local_crops, masks, mask_indices, masks_weight, teacher_temp):
        ng, nl = self.n_global_crops, self.n_local_crops
        with torch.no_grad():                       # the centres to use now = the ones updated with the previous batch
            self.dino_loss.apply_center_update()
            self.ibot_patch_loss.apply_center_update()
        s_glob, s_loc = self.student["backbone"]([global_crops, local_crops], masks=[masks, None], is_training=True)
        s_patch = s_glob["x_norm_patchtokens"].flatten(0, 1).index_select(0, mask_indices)
        rows = torch.cat((s_loc["x_norm_clstoken"], s_glob["x_norm_clstoken"], s_patch))
        B = s_glob["x_norm_clstoken"].shape[0] // ng
        ce, losses, dsum, imean = self.student["dino_head"].forward_with_objective(
            rows, t_scores=t_out, dino_center=self.dino_loss.center, ibot_center=self.ibot_patch_loss.center,
            masks_weight=masks_weight.float().contiguous(), B=B, n_local=nl, teacher_temp=teacher_temp,
            student_temp=self.dino_loss.student_temp, dino_weight=self.dino_loss_weight,
            ibot_weight=self.ibot_loss_weight)
        self.dino_loss.register_center_stat(dsum, ng * B)
        self.ibot_patch_loss.register_center_stat(imean, 1)
        loss_dict = {"dino_global_crops_loss": losses[1], "ibot_loss": losses[2] / 2}
        if nl > 0:
            loss_dict["dino_local_crops_loss"] = losses[0]
        total = ce
        if self.koleo_loss_weight > 0:
            cls = s_glob["x_norm_clstoken"]
            k = self.koleo_loss_weight * (self.koleo_loss.forward_chunks(cls, 2) if hasattr(self.koleo_loss, "forward_chunks")
                                          else sum(self.koleo_loss(p) for p in cls.chunk(2)))
            loss_dict["koleo_loss"] = k / 2
            total = total + k
        return total, loss_dict

    @torch.no_grad()
    def update_teacher(self, m: float, ema_fn=None):
        """models.py:437-447: every teacher parameter <- m * teacher + (1 - m) * student."""
        if ema_fn is None:
            from .dinov2 import update_teacher as ema_fn
        for k in self.student.keys():
            ema_fn(list(self.student[k].parameters()), list(self.teacher[k].parameters()), m)

    def cancel_last_layer_grads(self) -> int:
        """Drop the gradients of the student head's last (weight-normalised prototype) layer: what the reference does after
        backward while `epoch <= freeze_last_layer_epochs` (`possibly_cancel_last_layer_grads`, dinov2/trainer.py:86-91 ->
        `cancel_gradients`, utils/_utils.py:416-419: `p.grad = None` for every parameter whose name contains the
        attribute).  Call between `backward()` and `optimizer.step()`; -> number of tensors whose gradient was dropped."""
        n = 0
        for name, p in self.named_parameters():
            if "student.dino_head.last_layer" in name or "student.ibot_head.last_layer" in name:
                if p.grad is not None:
                    n += 1
                p.grad = None
        return n
