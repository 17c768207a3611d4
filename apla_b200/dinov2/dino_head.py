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
Validated on the B200 in round 2 (tests/test_ssl_gpu.py::test_dino_head, whole-step cases)."""
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


def head_forward(x, g, v, weights, biases):
    """-> (scores [n, out] f32, saved tuple for `head_backward`).  x [n, in] any float dtype; b_i may be None."""
    a = x.detach().to(BF16).contiguous()
    acts, pre = [a], []
    for i, (W, b) in enumerate(zip(weights, biases)):
        bias = None if b is None else b.detach().float().contiguous()
        if i + 1 < len(weights):
            h, a = G.gemm_bias_gelu(a, _b(W), bias)                  # h = pre-activation (kept for GELU'), a = GELU(h)
            pre.append(h)
            acts.append(a)
        else:
            z = G.gemm_bias(a, _b(W), bias)                          # bottleneck rows, bf16
    zn = R.l2norm_fwd(z, 1e-12, BF16)                                # dino_head.py:38 (fp16 inputs do not occur here)
    wn = R.weightnorm_fwd(g.detach().contiguous(), v.detach().contiguous(), BF16)
    scores = torch.zeros(zn.shape[0], wn.shape[0], device=x.device, dtype=F32)
    G.gemm_bias_ls_residual(zn, wn, None, None, scores, out=scores)
    return scores, (g, v, z, zn, wn, tuple(weights), tuple(acts), tuple(pre), tuple(b is not None for b in biases))


def head_backward(saved, dscores, need_x=True, need_g=True, need_v=True, need_w=None):
    """Gradients of everything `head_forward` read, from dscores [n, out] (f32, or bf16 as `apla_soft_ce_bwd` /
    `apla_ssl_objective` write it).  -> (dx bf16 or None, dg, dv, [dW_0, db_0, dW_1, db_1, ...])."""
    g, v, z, zn, wn, weights, acts, pre, has_bias = saved
    L = len(weights)
    need_w = [True] * (2 * L) if need_w is None else list(need_w)
    dl = G.ls_cast(dscores.contiguous()) if dscores.dtype == F32 else dscores.contiguous()
    # prototype layer: dW = dl^T zn (fp32, split-K), then through the weight normalisation
    dwn = torch.zeros(wn.shape, device=dl.device, dtype=F32)
    G.proj_wgrad(dl, zn, dwn, wn.shape[0])
    dg, dv = R.weightnorm_bwd(g.detach().contiguous(), v.detach().contiguous(), dwn, need_dg=need_g, need_dv=need_v)
    d = G.gemm_dgrad(dl, wn.t().contiguous())                        # d zn
    d = R.l2norm_bwd(z, d, 1e-12)                                    # d z
    grads = [None] * (2 * L)
    for i in range(L - 1, -1, -1):
        W = weights[i]
        if need_w[2 * i]:
            dW = torch.zeros(W.shape, device=d.device, dtype=F32)
            G.proj_wgrad(d, acts[i], dW, W.shape[0])
            grads[2 * i] = dW.to(W.dtype)
        if has_bias[i] and need_w[2 * i + 1]:
            db = torch.zeros(W.shape[0], device=d.device, dtype=F32)
            G.colsum(d, db, W.shape[0])
            grads[2 * i + 1] = db
        wt = _b(W).t().contiguous()                                  # [in, out]: the dgrad GEMM's operand layout
        if i > 0:
            d = G.gemm_dgrad_gelu_bwd(d, wt, pre[i - 1])             # (d W) * GELU'(pre-activation of layer i-1)
        elif need_x:
            d = G.gemm_dgrad(d, wt)
    return (d if need_x else None), dg, dv, grads


def _pack(saved):
    g, v, z, zn, wn, weights, acts, pre, has_bias = saved
    return (g, v, z, zn, wn, *weights, *acts, *pre), (len(weights), has_bias)


def _unpack(tensors, meta):
    L, has_bias = meta
    return (*tensors[:5], tuple(tensors[5:5 + L]), tuple(tensors[5 + L:5 + 2 * L]), tuple(tensors[5 + 2 * L:]), has_bias)


class _DinoHeadFn(torch.autograd.Function):
    """(x [n, in], weight_g, weight_v, W_0, b_0, W_1, b_1, ...) -> scores [n, out] f32; b_i may be None."""

    @staticmethod
    def forward(ctx, x, g, v, *wb):
        scores, saved = head_forward(x, g, v, wb[0::2], wb[1::2])
        tensors, ctx.meta = _pack(saved)
        ctx.save_for_backward(*tensors)
        ctx.x_dtype = x.dtype
        return scores

    @staticmethod
    def backward(ctx, dscores):
        need = ctx.needs_input_grad
        dx, dg, dv, grads = head_backward(_unpack(ctx.saved_tensors, ctx.meta), dscores, need[0], need[1], need[2],
                                          need[3:])
        return (dx.to(ctx.x_dtype) if dx is not None else None, dg, dv, *grads)


class _HeadObjectiveFn(torch.autograd.Function):
    """Student head + the whole objective as ONE autograd node: scores -> `apla_ssl_objective` (forward and backward of the
    three cross-entropy terms in one native launch sequence, `ds` written once, in bf16) -> head backward.  The fp32
    `[rows, out]` score gradient and its bf16 cast pass of the separate path never exist.  Returns
    (dino_weight (local + global) + ibot_weight ibot,  losses [3] for reporting,  dino_batch_sum,  ibot_batch_mean)."""

    @staticmethod
    def forward(ctx, x, g, v, cfg, *wb):
        scores, saved = head_forward(x, g, v, wb[0::2], wb[1::2])
        res = R.ssl_objective(scores, cfg["t_scores"], cfg["dino_center"], cfg["ibot_center"], cfg["masks_weight"],
                              cfg["B"], cfg["n_local"], cfg["teacher_temp"], cfg["student_temp"], cfg["dino_weight"],
                              cfg["ibot_weight"], None, BF16, True)
        del scores
        tensors, ctx.meta = _pack(saved)
        ctx.save_for_backward(res["ds"], *tensors)
        ctx.x_dtype = x.dtype
        losses = res["losses"]
        total = cfg["dino_weight"] * (losses[0] + losses[1]) + cfg["ibot_weight"] * losses[2]
        ctx.mark_non_differentiable(losses, res["dino_batch_sum"], res["ibot_batch_mean"])
        return total, losses, res["dino_batch_sum"], res["ibot_batch_mean"]

    @staticmethod
    def backward(ctx, g_total, *_):
        need = ctx.needs_input_grad
        ds, *tensors = ctx.saved_tensors
        dx, dg, dv, grads = head_backward(_unpack(tensors, ctx.meta), ds, need[0], need[1], need[2], need[4:])
        # everything downstream of ds is linear in it: the upstream gradient scales the (small) results, not ds
        sc = lambda t: None if t is None else t * g_total.to(t.dtype)               # noqa: E731
        return (sc(dx.to(ctx.x_dtype)) if dx is not None else None, sc(dg), sc(dv), None, *[sc(t) for t in grads])


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

    def _check_widths(self):
        for m in self._linears() + [self.last_layer]:
            for d in (m.in_features, m.out_features):
                if d % 64:
                    raise RuntimeError(f"DINOHead: layer widths must be multiples of 64 for the GEMM tiles, got {d}")

    def forward(self, x):
        lin = self._linears()
        self._check_widths()
        shape = x.shape
        wb = []
        for m in lin:
            wb += [m.weight, m.bias]
        y = _DinoHeadFn.apply(x.reshape(-1, shape[-1]), self.last_layer.weight_g, self.last_layer.weight_v, *wb)
        return y.view(*shape[:-1], y.shape[-1])

    def forward_with_objective(self, x, *, t_scores, dino_center, ibot_center, masks_weight, B, n_local, teacher_temp,
                               student_temp=0.1, dino_weight=1.0, ibot_weight=1.0):
        """x = [local CLS rows | global CLS rows | masked patch rows] (the order DINOv2.forward concatenates them in,
        models.py:335-359).  -> (weighted loss of the three cross-entropy terms, losses [3] = dino_local, dino_global,
        2 * ibot_loss, dino_batch_sum, ibot_batch_mean); see `_HeadObjectiveFn`."""
        self._check_widths()
        wb = []
        for m in self._linears():
            wb += [m.weight, m.bias]
        cfg = dict(t_scores=t_scores, dino_center=dino_center, ibot_center=ibot_center, masks_weight=masks_weight, B=B,
                   n_local=n_local, teacher_temp=teacher_temp, student_temp=student_temp, dino_weight=dino_weight,
                   ibot_weight=ibot_weight)
        return _HeadObjectiveFn.apply(x, self.last_layer.weight_g, self.last_layer.weight_v, cfg, *wb)
