"""`FusedPatchEmbed`: the frozen patch embedding of a ViT (PatchEmbed.forward, src/utils/transformers/vit.py:291-307;
dinov2 layers/patch_embed.py) as the library's im2col-free patch extraction + tcgen05 GEMM (`apla_patchify` +
`apla_gemm_bias_fwd`) instead of a cuDNN convolution -- the same two kernels the step engine uses for its stem.
Under APLA the patch embedding is frozen (build_apla freezes everything but the projection rows), so there is no
backward; a patch embedding that requires gradients is refused.  State-dict keys are unchanged (`patch_embed.proj.*`)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .._lib import LIB, ptr, require_device, stream


class FusedPatchEmbed(nn.Module):
    def __init__(self, patch_embed: nn.Module):
        super().__init__()
        conv = patch_embed.proj
        if not isinstance(conv, nn.Conv2d) or conv.kernel_size != conv.stride or conv.kernel_size[0] != conv.kernel_size[1] \
                or conv.in_channels != 3 or conv.padding != (0, 0):
            raise TypeError("FusedPatchEmbed needs a Conv2d(3, D, kernel = stride = patch) patch projection")
        for name, child in patch_embed.named_children():
            self.add_module(name, child)
        for k in ("img_size", "patch_size", "num_patches"):
            if hasattr(patch_embed, k):
                setattr(self, k, getattr(patch_embed, k))
        self._w = None
        self._w_key = None

    def _weights(self, device):
        conv = self.proj
        key = (conv.weight.data_ptr(), conv.weight._version, str(device))
        if self._w is None or self._w_key != key:
            p = conv.kernel_size[0]
            D = conv.out_channels
            k = 3 * p * p
            kpad = (k + 63) // 64 * 64
            with torch.no_grad():
                w = torch.zeros(D, kpad, device=device, dtype=torch.float32)
                w[:, :k] = conv.weight.detach().to(device=device, dtype=torch.float32).reshape(D, k)
                b = (conv.bias.detach().to(device=device, dtype=torch.float32).contiguous() if conv.bias is not None
                     else torch.zeros(D, device=device))
            self._w, self._w_key = (w.to(torch.bfloat16).contiguous(), b, p, kpad), key
        return self._w

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """images fp32 [B, 3, S, S] -> patch tokens fp32 [B, (S/p)^2, D]"""
        if not x.is_cuda:
            raise RuntimeError("apla_b200.FusedPatchEmbed runs on CUDA (sm_100a) only; there is no CPU fallback")
        require_device()
        if self.proj.weight.requires_grad or (self.proj.bias is not None and self.proj.bias.requires_grad):
            raise RuntimeError("the fused patch embedding has no backward: its parameters must be frozen (APLA's freeze policy)")
        w, b, p, kpad = self._weights(x.device)
        B, C, S, S2 = x.shape
        if C != 3 or S != S2 or S % p:
            raise RuntimeError(f"images must be [B, 3, S, S] with S a multiple of the patch size {p}, got {tuple(x.shape)}")
        P = (S // p) ** 2
        xi = x.detach().to(torch.float32).contiguous()
        patches = torch.empty(B * P, kpad, device=x.device, dtype=torch.bfloat16)
        LIB.call("apla_patchify", ptr(xi), ptr(patches), B, S, p, kpad, stream())
        y = ops.gemm_bias(patches, w, b)
        return y.view(B, P, -1).to(torch.float32)


def fuse_patch_embed(model: nn.Module) -> nn.Module:
    """Swap `model.patch_embed` (a backbone, or a Classifier's `.backbone`) for a `FusedPatchEmbed` in place."""
    bb = model if hasattr(model, "patch_embed") else getattr(model, "backbone", None)
    if bb is None or not hasattr(bb, "patch_embed"):
        raise AttributeError("model exposes no .patch_embed")
    if not isinstance(bb.patch_embed, FusedPatchEmbed):
        bb.patch_embed = FusedPatchEmbed(bb.patch_embed)
    return model


def cache_pos_encoding(model: nn.Module) -> nn.Module:
    """Memoise the backbone's position-table resize (`interpolate_pos_encoding`, src/utils/transformers/vit.py:421-437 and
    dinov2_vits.py:176-208; `pos_for` of the host stand-in): the reference recomputes the bicubic interpolation of a
    FROZEN table every forward (1-2 ms per call at ViT-B/L).  The result depends on the arguments only
    through their shapes / dtypes and on the table, so it is cached per (table storage, version, argument signature);
    a table that requires gradients is never cached."""
    bb = model if hasattr(model, "pos_embed") else getattr(model, "backbone", None)
    if bb is None or not hasattr(bb, "pos_embed"):
        raise AttributeError("model exposes no .pos_embed")
    for name in ("interpolate_pos_encoding", "pos_for"):
        fn = getattr(bb, name, None)
        if fn is None or getattr(fn, "_apla_cached", False):
            continue
        cache = {}

        def cached(*args, _fn=fn, _cache=cache, _bb=bb):
            table = _bb.pos_embed
            if table.requires_grad and torch.is_grad_enabled():
                return _fn(*args)
            sig = tuple((tuple(a.shape), a.dtype, str(a.device)) if torch.is_tensor(a) else a for a in args)
            key = (table.data_ptr(), table._version, sig)
            hit = _cache.get(key)
            if hit is None:
                if len(_cache) > 8:
                    _cache.clear()
                with torch.no_grad():
                    hit = _cache[key] = _fn(*args).detach()
            return hit

        cached._apla_cached = True
        setattr(bb, name, cached)
    return model
