"""CPU oracle for the APLA fine-tune step  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under ``apla_b200/`` imports it
and the product path never routes through it.

It is a plain fp32 PyTorch-on-CPU *restatement* of the reference algorithm (the reference is
100 % Python/PyTorch, so fp32 torch ops are the reference arithmetic):

* weights live in a flat ``dict`` keyed exactly like the reference ``state_dict``
  (``backbone.blocks.3.attn.proj_weight1`` ...), no ``nn.Module`` tree;
* the random-number stream (``torch.manual_seed`` -> constructor order) is replayed so the
  weights and the APLA index draws are bit-identical to the reference's;
* the forward is written as explicit functional math; the backward of the trainable tensors
  is taken with autograd over that fp32 math and, for the APLA projection, also restated in
  closed form (``proj_wgrad_closed_form``) so kernels can be checked piecewise.

Pinning: ``tests/golden/make_golden.py`` imports the UNMODIFIED reference from
``/root/reference/src`` in the build container and records weights digests, indices, logits,
loss, gradients and post-step parameters; ``tests/test_oracle_golden.py`` replays them against
this file (bit-exact indices / weights, <=1e-5 relative on floats).  Parity is therefore
*pinned* for configs C1, C2, C3, C5 shapes; the varlen (xformers) path of C4 is restated from
the call sites (xformers is not available) and pinned MODULO A PLAIN-TORCH XFORMERS
SHIM by tests/golden/make_golden_ssl_step.py, which runs the reference's packed multi-crop blocks
through it (tests/test_ssl_oracle.py::test_ssl_step_matches_reference).

Reference citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class VitCfg:
    """Shape of the host ViT.  Factories: src/utils/transformers/vit.py:511-596."""
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    patch_size: int = 14
    img_size: int = 518           # size the pos-embed table is built for (vit.py:322-327)
    mlp_ratio: float = 4.0
    qkv_bias: bool = True
    layerscale: Optional[float] = 1.0   # block_conf.layerscale_init_values, None = Identity
    ln_eps: float = 1e-6          # partial(nn.LayerNorm, eps=1e-6)  vit.py:519,536,554,571
    n_classes: int = 555          # Classifier.fc  src/defaults/models.py:65
    partial_size: object = 8      # int, or 'full' (multi-GPU only, apla_vit.py:65-75)
    is_multi_gpu: bool = False
    inds: Optional[Dict[str, List[int]]] = None   # content of an inds_path json (apla_vit.py:20-24)

    @property
    def hidden(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)

    @property
    def num_patches_table(self) -> int:
        return (self.img_size // self.patch_size) ** 2


VIT_S16 = dict(embed_dim=384, depth=12, num_heads=6, patch_size=16, img_size=224)
VIT_B14 = dict(embed_dim=768, depth=12, num_heads=12, patch_size=14, img_size=518)
VIT_L14 = dict(embed_dim=1024, depth=24, num_heads=16, patch_size=14, img_size=518)


# --------------------------------------------------------------------------------------
# weight construction: replays the reference constructor order on the global CPU RNG
# --------------------------------------------------------------------------------------
def _trunc_normal_(t: Tensor, std: float = 0.02, mean: float = 0.0, a: float = -2.0, b: float = 2.0) -> Tensor:
    """Reference's own truncated normal (vit.py:34-69): uniform -> erfinv -> scale -> clamp."""
    def norm_cdf(x):
        return (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0
    with torch.no_grad():
        lo = norm_cdf((a - mean) / std)
        up = norm_cdf((b - mean) / std)
        t.uniform_(2 * lo - 1, 2 * up - 1)
        t.erfinv_()
        t.mul_(std * math.sqrt(2.0))
        t.add_(mean)
        t.clamp_(min=a, max=b)
    return t


def _consume_linear(in_f: int, out_f: int, bias: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    """nn.Linear's default reset_parameters draws (kaiming_uniform_ + bias uniform_)."""
    m = nn.Linear(in_f, out_f, bias=bias)
    return m.weight.detach().clone(), (m.bias.detach().clone() if bias else None)


def build_state(cfg: VitCfg, seed: Optional[int] = 0) -> Dict[str, Tensor]:
    """Weights of Classifier(backbone=vit+APLA, fc) in reference construction order.

    Order replayed (each step consumes the global CPU generator exactly as the reference):
      1. VisionTransformer.__init__  vit.py:312-341: PatchEmbed conv (default init), blocks'
         Linears (default init; overwritten later but the draws still advance the stream),
         trunc_normal_(pos_embed), trunc_normal_(cls_token), then ``self.apply(_init_weights)``
         which re-draws every nn.Linear weight with trunc_normal_(std=.02) in module order.
      2. build_apla / replace_attn_with_apla  apla_vit.py:63-101,11-60: per block
         ``torch.randperm(dim)`` (appla_attn.py:26) THEN ``nn.Linear(dim, 3*dim)`` (appla_attn.py:37).
      3. Classifier.fc = nn.Linear(D, n_classes)  models.py:65 (default init, not re-drawn).
    """
    if seed is not None:
        torch.manual_seed(seed)
    D, L, Hd = cfg.embed_dim, cfg.depth, cfg.hidden
    p = cfg.patch_size
    sd: Dict[str, Tensor] = {}
    pre = "backbone."

    # -- 1. VisionTransformer.__init__ -------------------------------------------------
    conv = nn.Conv2d(3, D, kernel_size=p, stride=p, bias=True)          # vit.py:302
    sd[pre + "patch_embed.proj.weight"] = conv.weight.detach().clone()
    sd[pre + "patch_embed.proj.bias"] = conv.bias.detach().clone()
    sd[pre + "cls_token"] = torch.zeros(1, 1, D)
    sd[pre + "pos_embed"] = torch.zeros(1, cfg.num_patches_table + 1, D)
    for i in range(L):                                                   # vit.py:330-335
        _consume_linear(D, 3 * D, cfg.qkv_bias)                          # Attention.qkv  vit.py:179
        _consume_linear(D, D)                                            # Attention.proj vit.py:181
        _consume_linear(D, Hd)                                           # Mlp.fc1        vit.py:157
        _consume_linear(Hd, D)                                           # Mlp.fc2        vit.py:159
    _trunc_normal_(sd[pre + "pos_embed"], std=.02)                       # vit.py:340
    _trunc_normal_(sd[pre + "cls_token"], std=.02)                       # vit.py:341
    # self.apply(_init_weights)  vit.py:342-351 : children first, registration order
    full_proj_w, full_proj_b = [], []
    for i in range(L):
        b = f"{pre}blocks.{i}."
        sd[b + "norm1.weight"] = torch.ones(D)
        sd[b + "norm1.bias"] = torch.zeros(D)
        sd[b + "attn.qkv.weight"] = _trunc_normal_(torch.empty(3 * D, D), std=.02)
        if cfg.qkv_bias:
            sd[b + "attn.qkv.bias"] = torch.zeros(3 * D)
        full_proj_w.append(_trunc_normal_(torch.empty(D, D), std=.02))
        full_proj_b.append(torch.zeros(D))
        sd[b + "norm2.weight"] = torch.ones(D)
        sd[b + "norm2.bias"] = torch.zeros(D)
        sd[b + "mlp.fc1.weight"] = _trunc_normal_(torch.empty(Hd, D), std=.02)
        sd[b + "mlp.fc1.bias"] = torch.zeros(Hd)
        sd[b + "mlp.fc2.weight"] = _trunc_normal_(torch.empty(D, Hd), std=.02)
        sd[b + "mlp.fc2.bias"] = torch.zeros(D)
        if cfg.layerscale is not None:                                   # vit.py:267-273
            sd[b + "ls1.gamma"] = cfg.layerscale * torch.ones(D)
            sd[b + "ls2.gamma"] = cfg.layerscale * torch.ones(D)
    sd[pre + "norm.weight"] = torch.ones(D)
    sd[pre + "norm.bias"] = torch.zeros(D)

    # -- 2. build_apla ------------------------------------------------------------------
    if cfg.is_multi_gpu and cfg.partial_size == "full":
        # apla_vit.py:65-75: stock Attention kept, attn.proj.* trainable
        for i in range(L):
            b = f"{pre}blocks.{i}.attn."
            sd[b + "proj.weight"] = full_proj_w[i]
            sd[b + "proj.bias"] = full_proj_b[i]
    else:
        if cfg.is_multi_gpu:
            assert cfg.inds is not None, '"inds_path" should be present with multi-gpu training with random sampling'
        r = int(cfg.partial_size)
        for i in range(L):
            b = f"{pre}blocks.{i}.attn."
            if cfg.inds is not None:                                     # apla_vit.py:20-24
                tr = list(cfg.inds[f"block_{i}"])
                fr = [j for j in range(D) if j not in tr]
                inds = torch.tensor(tr + fr)
            else:
                inds = torch.randperm(D)                                 # appla_attn.py:26
            _consume_linear(D, 3 * D, cfg.qkv_bias)                      # appla_attn.py:37
            sd[b + "inds"] = inds
            sd[b + "proj_weight1"] = full_proj_w[i][inds[:r], :].clone()  # apla_vit.py:51-52
            sd[b + "proj_weight2"] = full_proj_w[i][inds[r:], :].clone()
            sd[b + "proj_bias1"] = full_proj_b[i][inds[:r]].clone()      # apla_vit.py:55-56
            sd[b + "proj_bias2"] = full_proj_b[i][inds[r:]].clone()

    # -- 3. Classifier head -------------------------------------------------------------
    w, bb = _consume_linear(D, cfg.n_classes)
    sd["fc.weight"], sd["fc.bias"] = w, bb
    return sd


def trainable_keys(cfg: VitCfg, sd: Dict[str, Tensor]) -> List[str]:
    """Invariant I5: proj_weight1/proj_bias1 per block + fc, in named_parameters order."""
    keys = []
    for i in range(cfg.depth):
        b = f"backbone.blocks.{i}.attn."
        if b + "proj_weight1" in sd:
            keys += [b + "proj_weight1", b + "proj_bias1"]
        else:
            keys += [b + "proj.weight", b + "proj.bias"]
    return keys + ["fc.weight", "fc.bias"]


def perturb_state(sd: Dict[str, Tensor], seed: int = 7, scale: float = 0.05) -> None:
    """Make biases / LN / LayerScale non-trivial so parity tests exercise every term.

    Applied identically to the imported reference model in make_golden.py (same key order,
    same generator).  Skips integer buffers."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        t = sd[k]
        if not t.is_floating_point():
            continue
        t.add_(torch.randn(t.shape, generator=g) * scale * (0.2 if t.dim() > 1 else 1.0))


# --------------------------------------------------------------------------------------
# functional forward
# --------------------------------------------------------------------------------------
def interpolate_pos_encoding(pos_embed: Tensor, npatch: int) -> Tensor:
    """vit.py:421-437 (bicubic, align_corners=False, scale_factor = sqrt(npatch/N))."""
    N = pos_embed.shape[1] - 1
    if npatch == N:
        return pos_embed
    dim = pos_embed.shape[-1]
    class_emb = pos_embed[:, 0]
    grid = pos_embed[:, 1:]
    s = int(math.sqrt(N))
    grid = F.interpolate(grid.reshape(1, s, s, dim).permute(0, 3, 1, 2),
                         scale_factor=math.sqrt(npatch / N), mode="bicubic",
                         align_corners=False, recompute_scale_factor=False)
    grid = grid.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((class_emb.unsqueeze(0), grid), dim=1)


def softmax_attention(qkv: Tensor, B: int, N: int, H: int, scale: float) -> Tuple[Tensor, Tensor]:
    """appla_attn.py:53-60.  qkv is [B,N,3*D] laid out (3,H,hd) along the last dim."""
    C = qkv.shape[-1] // 3
    t = qkv.reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    attn = (q @ k.transpose(-2, -1)) * scale
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return out, attn


def varlen_attention(qkv: Tensor, seqlens: Sequence[int], H: int, scale: float) -> Tensor:
    """Restatement of xformers memory_efficient_attention with a BlockDiagonalMask
    (appla_attn_mem_eff.py:37-43; dinov2/layers/block.py:191-217): independent softmax
    attention per original sequence of the packed [1, sum(N), 3D] tensor.  Pinned modulo the
    xformers shim (module docstring): xformers 0.0.18 itself is absent from this image and /root/reference."""
    outs, o = [], 0
    for n in seqlens:
        out, _ = softmax_attention(qkv[:, o:o + n], 1, n, H, scale)
        outs.append(out)
        o += n
    return torch.cat(outs, dim=1)


def apla_proj(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, inds: Tensor) -> Tensor:
    """appla_attn.py:64-79: two linears, outputs interleaved by two scatter_ along the feature dim."""
    r = w1.shape[0]
    t_out = F.linear(x, w1, b1)
    f_out = F.linear(x, w2, b2)
    out = torch.empty(x.shape, dtype=t_out.dtype)
    out.scatter_(-1, inds[:r].view(1, 1, -1).expand(x.size(0), x.size(1), -1), t_out)
    out.scatter_(-1, inds[r:].view(1, 1, -1).expand(x.size(0), x.size(1), -1), f_out)
    return out


def proj_wgrad_closed_form(dy: Tensor, x: Tensor, inds: Tensor, r: int) -> Tuple[Tensor, Tensor]:
    """What autograd yields for proj_weight1 / proj_bias1 (scatter_ backward = gather):
    dW1[r,D] = dY[:, idx]^T . X ,  db1 = sum_t dY[t, idx]."""
    idx = inds[:r]
    dy2 = dy.reshape(-1, dy.shape[-1])[:, idx]
    return dy2.t() @ x.reshape(-1, x.shape[-1]), dy2.sum(0)


def attention_module(sd: Dict[str, Tensor], b: str, x: Tensor, H: int,
                     seqlens: Optional[Sequence[int]] = None) -> Tensor:
    """APLA_Attention.forward (appla_attn.py:50-83) or, for multi-GPU 'full', stock Attention
    (vit.py:184-196).  ``b`` is the key prefix '...blocks.i.attn.'."""
    B, N, C = x.shape
    scale = (C // H) ** -0.5
    qkv = F.linear(x, sd[b + "qkv.weight"], sd.get(b + "qkv.bias"))
    if seqlens is None:
        a, _ = softmax_attention(qkv, B, N, H, scale)
    else:
        a = varlen_attention(qkv, seqlens, H, scale)
    if b + "proj_weight1" in sd:
        return apla_proj(a, sd[b + "proj_weight1"], sd[b + "proj_bias1"],
                         sd[b + "proj_weight2"], sd[b + "proj_bias2"], sd[b + "inds"])
    return F.linear(a, sd[b + "proj.weight"], sd[b + "proj.bias"])


def block_forward(sd: Dict[str, Tensor], b: str, x: Tensor, H: int, eps: float,
                  seqlens: Optional[Sequence[int]] = None) -> Tensor:
    """Block.forward vit.py:279-288 (drop_path = identity at rate 0); Mlp.forward vit.py:162-168
    (exact erf GELU, vit.py:153); LayerScale vit.py:243-244."""
    D = x.shape[-1]
    y = attention_module(sd, b + "attn.", F.layer_norm(x, (D,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], eps), H, seqlens)
    if b + "ls1.gamma" in sd:
        y = y * sd[b + "ls1.gamma"]
    x = x + y
    h = F.layer_norm(x, (D,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], eps)
    h = F.linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
    if b + "ls2.gamma" in sd:
        h = h * sd[b + "ls2.gamma"]
    return x + h


def embed(sd: Dict[str, Tensor], cfg: VitCfg, images: Tensor) -> Tensor:
    """PatchEmbed.forward vit.py:304-307 + forward_features vit.py:389-396."""
    pre = "backbone."
    B = images.shape[0]
    x = F.conv2d(images, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"],
                 stride=cfg.patch_size).flatten(2).transpose(1, 2)
    x = torch.cat((sd[pre + "cls_token"].expand(B, -1, -1), x), dim=1)
    return x + interpolate_pos_encoding(sd[pre + "pos_embed"], x.shape[1] - 1)


def forward_tokens(sd: Dict[str, Tensor], cfg: VitCfg, images: Tensor) -> Tensor:
    """All tokens after the last block and the final LayerNorm (vit.py:414-417)."""
    x = embed(sd, cfg, images)
    for i in range(cfg.depth):
        x = block_forward(sd, f"backbone.blocks.{i}.", x, cfg.num_heads, cfg.ln_eps)
    D = x.shape[-1]
    return F.layer_norm(x, (D,), sd["backbone.norm.weight"], sd["backbone.norm.bias"], cfg.ln_eps)


def forward_logits(sd: Dict[str, Tensor], cfg: VitCfg, images: Tensor) -> Tensor:
    """Classifier.forward models.py:81-92 (CLS token, vit.py:419, then fc)."""
    return F.linear(forward_tokens(sd, cfg, images)[:, 0], sd["fc.weight"], sd["fc.bias"])


# --------------------------------------------------------------------------------------
# the step: Trainer.global_step  src/defaults/trainer.py:106-138
# --------------------------------------------------------------------------------------
@dataclass
class StepResult:
    logits: Tensor
    loss: Tensor
    grads: Dict[str, Tensor]
    grad_norm: Optional[Tensor] = None
    new_params: Dict[str, Tensor] = field(default_factory=dict)


def loss_and_grads(sd: Dict[str, Tensor], cfg: VitCfg, images: Tensor, labels: Tensor) -> StepResult:
    """forward -> CrossEntropyLoss(mean) (wrappers.py:314) -> backward for the trainable tensors."""
    keys = trainable_keys(cfg, sd)
    leaves = {}
    work = dict(sd)
    for k in keys:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    logits = forward_logits(work, cfg, images)
    loss = F.cross_entropy(logits, labels)
    gs = torch.autograd.grad(loss, [leaves[k] for k in keys])
    return StepResult(logits.detach(), loss.detach(), {k: g for k, g in zip(keys, gs)})


def clip_grad_norm_(grads: Dict[str, Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ (trainer.py:136): total L2 norm, coef = min(1, max/(norm+1e-6))."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads.values()]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads.values():
        g.mul_(coef)
    return total


@dataclass
class AdamWState:
    step: int = 0
    exp_avg: Dict[str, Tensor] = field(default_factory=dict)
    exp_avg_sq: Dict[str, Tensor] = field(default_factory=dict)


def adamw_step(sd: Dict[str, Tensor], grads: Dict[str, Tensor], st: AdamWState, lr: float = 3e-5,
               weight_decay: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """torch.optim.AdamW with the reference's two groups (wrappers.py:205-221): tensors whose
    name ends in '.bias' or that are 1-D get weight_decay 0."""
    st.step += 1
    b1, b2 = betas
    for k, g in grads.items():
        p = sd[k]
        wd = 0.0 if (k.endswith(".bias") or p.dim() == 1) else weight_decay
        if k not in st.exp_avg:
            st.exp_avg[k] = torch.zeros_like(p)
            st.exp_avg_sq[k] = torch.zeros_like(p)
        m, v = st.exp_avg[k], st.exp_avg_sq[k]
        p.mul_(1 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** st.step
        bc2 = 1 - b2 ** st.step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def fine_tune_step(sd: Dict[str, Tensor], cfg: VitCfg, images: Tensor, labels: Tensor, st: AdamWState,
                   lr: float = 3e-5, weight_decay: float = 1e-5, clip: float = 1.0,
                   world_grads: Optional[List[Dict[str, Tensor]]] = None) -> StepResult:
    """One Trainer.global_step in fp32: fwd, CE, bwd, (DDP mean over ranks), clip 1.0, AdamW.
    ``world_grads``: gradients of the other data-parallel ranks to average with (DDP semantics)."""
    res = loss_and_grads(sd, cfg, images, labels)
    grads = {k: g.clone() for k, g in res.grads.items()}
    if world_grads:
        n = 1 + len(world_grads)
        for k in grads:
            for other in world_grads:
                grads[k] += other[k]
            grads[k] /= n
    res.grad_norm = clip_grad_norm_(grads, clip) if clip else None
    adamw_step(sd, grads, st, lr=lr, weight_decay=weight_decay)
    res.new_params = {k: sd[k].clone() for k in grads}
    return res


def synthetic_batch(batch: int, img: int, n_classes: int, seed: int = 1234, rank: int = 0) -> Tuple[Tensor, Tensor]:
    """g = Generator().manual_seed(1234 + rank); randn images, randint labels."""
    g = torch.Generator().manual_seed(seed + rank)
    images = torch.randn(batch, 3, img, img, generator=g)
    labels = torch.randint(0, n_classes, (batch,), generator=g)
    return images, labels


def flops_per_image(cfg: VitCfg, img: int, r: Optional[int] = None) -> Tuple[float, float]:
    """SURVEY Appendix B FLOP model -> (forward, backward) algorithmic FLOPs per image."""
    D, L, p = cfg.embed_dim, cfg.depth, cfg.patch_size
    P = (img // p) ** 2
    N = P + 1
    r = D if (r is None and cfg.partial_size == "full") else (int(cfg.partial_size) if r is None else r)
    lin = 24 * D * D
    att = 4 * N * D
    fwd = 2 * 3 * p * p * D * P + L * N * (lin + att) + 2 * D * cfg.n_classes
    bwd = (L - 1) * N * (lin + 2.5 * att + 2 * r * D) + N * (16 * D * D + 2 * r * D) + 2 * (2 * D * cfg.n_classes)
    return float(fwd), float(bwd)
