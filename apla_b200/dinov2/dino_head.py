"""`DINOHead` of the reference (src/self_supervised/dinov2/layers/dino_head.py:12-58) with every parameter TRAINABLE, as
ONE autograd node over the library's kernels: bf16 tcgen05 GEMMs with the bias / GELU epilogues for the MLP
(`apla_gemm_bias_gelu_fwd`, `apla_gemm_dgrad_gelu_bwd`, split-K `apla_proj_wgrad_gather` for the weight gradients,
`apla_colsum` for the bias gradients), the row kernels of csrc/ssl.cu for `F.normalize` and `weight_norm`, fp32 prototype
scores out.  Same constructor, same state-dict keys (`mlp.{0,2,4}.weight/bias`, `last_layer.weight_g/weight_v`), same
initialisation draws (the parameter containers ARE torch's `nn.Linear` / `weight_norm`; only `forward` differs).

Differences from the reference: BatchNorm (`use_bn=True`) raises -- no shipped config sets it; the MLP runs in bf16 with
fp32 accumulation (the reference runs it in fp16 under autocast); input rows must be CUDA tensors.
v0 of this node trades HBM traffic for reuse of kernels that are already validated: the fp32 `[rows, out_dim]` scores are
zero-filled and then written through the residual epilogue, and the incoming score gradient is cast to bf16 in a separate
pass.  DESIGN.md section 9 lists what replaces both (cross-entropy in the prototype GEMM's epilogue).
STATUS: composition of validated GEMM kernels and the not-yet-run ssl.cu kernels; first hardware run pending."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.init import trunc_normal_
from torch.nn.utils import weight_norm

from .. import ops as G
from . import ops as R

BF16, F32 = torch.bfloat16, torch.float32


def _b(t):
    return t.detach().to(BF16).contiguous()


class _DinoHeadFn(torch.autograd.Function):
    """(x [n, in], weight_g, weight_v, W_0, b_0, W_1, b_1, ...) -> scores [n, out] f32; b_i may be None."""

    @staticmethod
    def forward(ctx, x, g, v, *wb):
        weights, biases = wb[0::2], wb[1::2]
        a = x.detach().to(BF16).contiguous()
        acts, pre = [a], []
        for i, (W, b) in enumerate(zip(weights, biases)):
            bias = None if b is None else b.detach().float().contiguous()
            if i + 1 < len(weights):
                h, a = G.gemm_bias_gelu(a, _b(W), bias)              # h = pre-activation (kept for GELU'), a = GELU(h)
                pre.append(h)
                acts.append(a)
            else:
                z = G.gemm_bias(a, _b(W), bias)                      # bottleneck rows, bf16
        eps = 1e-12                                                  # dino_head.py:38 (fp16 inputs do not occur here)
        zn = R.l2norm_fwd(z, eps, BF16)
        wn = R.weightnorm_fwd(g.detach().contiguous(), v.detach().contiguous(), BF16)
        scores = torch.zeros(zn.shape[0], wn.shape[0], device=x.device, dtype=F32)
        G.gemm_bias_ls_residual(zn, wn, None, None, scores, out=scores)
        ctx.save_for_backward(g, v, z, zn, wn, *weights, *acts, *pre)
        ctx.n_layers = len(weights)
        ctx.has_bias = [b is not None for b in biases]
        ctx.x_dtype = x.dtype
        return scores

    @staticmethod
    def backward(ctx, dscores):
        L = ctx.n_layers
        saved = ctx.saved_tensors
        g, v, z, zn, wn = saved[:5]
        weights, acts, pre = saved[5:5 + L], saved[5 + L:5 + 2 * L], saved[5 + 2 * L:]
        need = ctx.needs_input_grad
        dl = G.ls_cast(dscores.contiguous()) if dscores.dtype == F32 else dscores.contiguous()
        # prototype layer: dW = dl^T zn (fp32, split-K), then through the weight normalisation
        dwn = torch.zeros(wn.shape, device=dl.device, dtype=F32)
        G.proj_wgrad(dl, zn, dwn, wn.shape[0])
        dg, dv = R.weightnorm_bwd(g.detach().contiguous(), v.detach().contiguous(), dwn, need_dg=need[1],
                                  need_dv=need[2])
        d = G.gemm_dgrad(dl, wn.t().contiguous())                    # d zn
        d = R.l2norm_bwd(z, d, 1e-12)                                # d z
        grads = [None] * (2 * L)
        for i in range(L - 1, -1, -1):
            W = weights[i]
            if need[3 + 2 * i]:
                dW = torch.zeros(W.shape, device=d.device, dtype=F32)
                G.proj_wgrad(d, acts[i], dW, W.shape[0])
                grads[2 * i] = dW.to(W.dtype)
            if ctx.has_bias[i] and need[4 + 2 * i]:
                db = torch.zeros(W.shape[0], device=d.device, dtype=F32)
                G.colsum(d, db, W.shape[0])
                grads[2 * i + 1] = db
            wt = _b(W).t().contiguous()                              # [in, out]: the dgrad GEMM's operand layout
            if i > 0:
                d = G.gemm_dgrad_gelu_bwd(d, wt, pre[i - 1])         # (d W) * GELU'(pre-activation of layer i-1)
            elif need[0]:
                d = G.gemm_dgrad(d, wt)
        dx = d.to(ctx.x_dtype) if need[0] else None
        return (dx, dg, dv, *grads)


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, nlayers=3, hidden_dim=2048, bottleneck_dim=256, mlp_bias=True):
        super().__init__()
        if use_bn:
            raise NotImplementedError("DINOHead(use_bn=True) has no B200 path (no shipped config uses it)")
        nlayers = max(nlayers, 1)
        dims = [in_dim] + [hidden_dim] * (nlayers - 1) + [bottleneck_dim]
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim, bias=mlp_bias)          # key `mlp.weight` (dino_head.py:45-46)
        else:
            layers = []
            for i in range(nlayers):
                layers.append(nn.Linear(dims[i], dims[i + 1], bias=mlp_bias))
                if i + 1 < nlayers:
                    layers.append(nn.GELU())                                     # placeholder: keeps the keys mlp.0/2/4
            self.mlp = nn.Sequential(*layers)
        for m in self.modules():                                                 # self.apply(_init_weights), :26,30-34
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self.last_layer = weight_norm(nn.Linear(bottleneck_dim, out_dim, bias=False))
        self.last_layer.weight_g.data.fill_(1)

    def _linears(self):
        return [self.mlp] if isinstance(self.mlp, nn.Linear) else [m for m in self.mlp if isinstance(m, nn.Linear)]

    def forward(self, x):
        lin = self._linears()
        for m in lin + [self.last_layer]:
            for d in (m.in_features, m.out_features):
                if d % 64:
                    raise RuntimeError(f"DINOHead: layer widths must be multiples of 64 for the GEMM tiles, got {d}")
        shape = x.shape
        wb = []
        for m in lin:
            wb += [m.weight, m.bias]
        y = _DinoHeadFn.apply(x.reshape(-1, shape[-1]), self.last_layer.weight_g, self.last_layer.weight_v, *wb)
        return y.view(*shape[:-1], y.shape[-1])
