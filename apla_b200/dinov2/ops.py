"""Tensor-level wrappers over the self-supervised row kernels of the C ABI (include/apla_b200.h, section "DINOv2
self-supervised objective"; kernels in csrc/ssl.cu).  Same rules as apla_b200/ops.py: torch supplies device memory and
the stream, every wrapper validates and raises, nothing falls back to PyTorch math.

STATUS: the kernels are built for sm_100a but have not run on hardware yet (round 1's GPU budget was spent before they
were written); tests/test_ssl_gpu.py holds them to oracle/ssl_oracle.py and is the first GPU job of round 2.  Until then
tests/test_ssl_emu.py runs the same tests on the kernel source executed by a CPU SIMT emulator."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .._lib import LIB, ptr, require_device, stream

F32, BF16, I32 = torch.float32, torch.bfloat16, torch.int32


def _rows(t: torch.Tensor, name: str, dtype=F32) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (apla_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name} must be 2-D with contiguous rows, got shape {tuple(t.shape)} strides {t.stride()}")
    return t


def softmax_center(teacher_output: torch.Tensor, center: torch.Tensor, teacher_temp: float,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax((teacher_output - center) / teacher_temp, dim=-1); teacher_output [n, K] f32, center K floats."""
    require_device()
    t = _rows(teacher_output, "teacher_output")
    n, K = t.shape
    if center.numel() != K or center.dtype != F32 or not center.is_contiguous():
        raise RuntimeError(f"center must hold {K} contiguous floats")
    if out is None:
        out = torch.empty(n, K, device=t.device, dtype=F32)
    LIB.call("apla_softmax_center", ptr(t), t.stride(0), ptr(center), 1.0 / float(teacher_temp), n, K, ptr(out),
             out.stride(0), stream())
    return out


def colsum(a: torch.Tensor, scale: float = 1.0, splits: Optional[int] = None) -> torch.Tensor:
    """[1, K] = scale * column sums of a [n, K] f32 matrix, in a fixed order."""
    require_device()
    a = _rows(a, "a")
    n, K = a.shape
    if splits is None:
        splits = max(1, min(32, n // 64))
    ws = torch.empty(splits, K, device=a.device, dtype=F32)
    out = torch.empty(1, K, device=a.device, dtype=F32)
    LIB.call("apla_colsum_f32", ptr(a), a.stride(0), n, K, ptr(ws), splits, float(scale), ptr(out), stream())
    return out


def center_ema_(center: torch.Tensor, batch_sum: torch.Tensor, count: float, momentum: float) -> torch.Tensor:
    """center <- center * momentum + batch_sum / count * (1 - momentum), in place."""
    require_device()
    K = center.numel()
    if batch_sum.numel() != K or center.dtype != F32 or batch_sum.dtype != F32:
        raise RuntimeError("center and batch_sum must be f32 of the same length")
    if not (center.is_contiguous() and batch_sum.is_contiguous()):
        raise RuntimeError("center and batch_sum must be contiguous")
    LIB.call("apla_center_ema", ptr(center), ptr(batch_sum), K, 1.0 / float(count), float(momentum), stream())
    return center


def sinkhorn_knopp(teacher_output: torch.Tensor, teacher_temp: float, n_iterations: int, n_samples_world: int,
                   all_reduce=None) -> torch.Tensor:
    """Sinkhorn-Knopp targets [n, K] (rows sum to 1) from teacher scores [n, K]; `all_reduce(t)` sums a tensor over the
    ranks in place (None = single process); `n_samples_world` = B of the reference (samples over all ranks)."""
    require_device()
    t = _rows(teacher_output, "teacher_output")
    n, K = t.shape
    P = torch.empty(n, K, device=t.device, dtype=F32)
    LIB.call("apla_sk_exp", ptr(t), t.stride(0), 1.0 / float(teacher_temp), n, K, ptr(P), P.stride(0), stream())
    for it in range(n_iterations):
        cs = colsum(P)
        if all_reduce is not None:
            all_reduce(cs)
        row_scale = 1.0 if it + 1 == n_iterations else 1.0 / float(n_samples_world)
        LIB.call("apla_sk_normalize", ptr(P), P.stride(0), n, K, ptr(cs), 1.0 / K, row_scale, stream())
    return P


def soft_ce_fwd(s, t0, t1, t_rows, w_row, w_uniform, inv_temp) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (loss scalar, lse [rows], mass [rows])."""
    require_device()
    s = _rows(s, "student scores")
    t0 = _rows(t0, "teacher targets")
    rows, K = s.shape
    if t0.shape[1] != K or t0.shape[0] < min(t_rows, max(rows, 1)):
        raise RuntimeError(f"teacher targets {tuple(t0.shape)} do not cover {t_rows} rows of width {K}")
    if t1 is not None:
        t1 = _rows(t1, "second teacher targets")
        if t1.shape != t0.shape or t1.stride(0) != t0.stride(0):
            raise RuntimeError("the two teacher target matrices must have the same shape and row stride")
    if w_row is not None and (w_row.dtype != F32 or w_row.numel() < rows or not w_row.is_contiguous()):
        raise RuntimeError("row weights must be contiguous f32, one per student row")
    row_loss = torch.empty(max(rows, 1), device=s.device, dtype=F32)
    lse = torch.empty_like(row_loss)
    mass = torch.empty_like(row_loss)
    loss = torch.empty((), device=s.device, dtype=F32)
    LIB.call("apla_soft_ce_fwd", ptr(s), s.stride(0), rows, K, ptr(t0), ptr(t1), t0.stride(0), int(t_rows), ptr(w_row),
             float(w_uniform), float(inv_temp), ptr(row_loss), ptr(lse), ptr(mass), stream())
    LIB.call("apla_sum_f32", ptr(row_loss), rows, 1.0, ptr(loss), stream())
    return loss, lse, mass


def soft_ce_bwd(s, t0, t1, t_rows, w_row, w_uniform, inv_temp, lse, mass, gscale: Optional[torch.Tensor],
                out_dtype=F32) -> torch.Tensor:
    """Gradient of the loss of `soft_ce_fwd` with respect to s, times the device scalar `gscale`."""
    require_device()
    rows, K = s.shape
    if out_dtype not in (F32, BF16):
        raise RuntimeError("ds must be f32 or bf16")
    ds = torch.empty(rows, K, device=s.device, dtype=out_dtype)
    if gscale is not None and (gscale.dtype != F32 or gscale.numel() != 1 or not gscale.is_cuda):
        raise RuntimeError("gscale must be a one-element f32 CUDA tensor")
    LIB.call("apla_soft_ce_bwd", ptr(s), s.stride(0), rows, K, ptr(t0), ptr(t1), t0.stride(0), int(t_rows), ptr(w_row),
             float(w_uniform), float(inv_temp), ptr(lse), ptr(mass), ptr(gscale), ptr(ds), ds.stride(0),
             int(out_dtype == BF16), stream())
    return ds


def l2norm_fwd(x: torch.Tensor, eps: float, out_dtype=F32) -> torch.Tensor:
    """F.normalize(x, dim=-1, p=2, eps=eps) for x [n, d] f32 or bf16."""
    require_device()
    x = _rows(x, "x", x.dtype)
    if x.dtype not in (F32, BF16) or out_dtype not in (F32, BF16):
        raise RuntimeError("l2norm: f32 or bf16 only")
    n, d = x.shape
    y = torch.empty(n, d, device=x.device, dtype=out_dtype)
    LIB.call("apla_l2norm_fwd", ptr(x), x.stride(0), int(x.dtype == F32), n, d, float(eps),
             ptr(y) if out_dtype == BF16 else None, ptr(y) if out_dtype == F32 else None, y.stride(0), stream())
    return y


def l2norm_bwd(x: torch.Tensor, dy: torch.Tensor, eps: float) -> torch.Tensor:
    require_device()
    x = _rows(x, "x", x.dtype)
    dy = _rows(dy, "dy", dy.dtype)
    if x.dtype not in (F32, BF16) or dy.dtype not in (F32, BF16) or dy.shape != x.shape:
        raise RuntimeError("l2norm_bwd: x and dy must be f32 or bf16 of one shape")
    n, d = x.shape
    dx = torch.empty(n, d, device=x.device, dtype=dy.dtype)
    LIB.call("apla_l2norm_bwd", ptr(x), x.stride(0), int(x.dtype == F32), ptr(dy), dy.stride(0), int(dy.dtype == F32), n,
             d, float(eps), ptr(dx), dx.stride(0), stream())
    return dx


def weightnorm_fwd(g: torch.Tensor, v: torch.Tensor, out_dtype=BF16) -> torch.Tensor:
    """W = g * v / ||v||_row  (torch.nn.utils.weight_norm, dim=0) for g [K, 1] or [K], v [K, d]."""
    require_device()
    v = _rows(v, "weight_v")
    K, d = v.shape
    if not v.is_contiguous() or g.numel() != K or g.dtype != F32 or not g.is_contiguous() or not g.is_cuda:
        raise RuntimeError("weight_g must hold K contiguous floats and weight_v must be contiguous")
    w = torch.empty(K, d, device=v.device, dtype=out_dtype)
    LIB.call("apla_weightnorm_fwd", ptr(g), ptr(v), K, d, ptr(w) if out_dtype == BF16 else None,
             ptr(w) if out_dtype == F32 else None, stream())
    return w


def weightnorm_bwd(g: torch.Tensor, v: torch.Tensor, dW: torch.Tensor, need_dg: bool = True, need_dv: bool = True):
    """-> (dg shaped like g or None, dv [K, d] or None) from dW [K, d] f32."""
    require_device()
    v = _rows(v, "weight_v")
    dW = _rows(dW, "dW")
    K, d = v.shape
    if dW.shape != v.shape or not v.is_contiguous():
        raise RuntimeError("dW must have the shape of weight_v")
    dg = torch.empty_like(g) if need_dg else None
    dv = torch.empty_like(v) if need_dv else None
    LIB.call("apla_weightnorm_bwd", ptr(g), ptr(v), ptr(dW), dW.stride(0), K, d, ptr(dg), ptr(dv), stream())
    return dg, dv


def koleo_fwd(x: torch.Tensor, eps: float, groups: int = 1, weight: float = 1.0):
    """x [groups * n, D] f32 -> (loss scalar summed over groups, xn, nn, dist)."""
    require_device()
    x = _rows(x, "student_output")
    if not x.is_contiguous() or x.shape[0] % groups:
        raise RuntimeError("koleo: x must be contiguous with groups * n rows")
    n, D = x.shape[0] // groups, x.shape[1]
    xn = l2norm_fwd(x, eps, F32)
    nn = torch.empty(groups * n, device=x.device, dtype=I32)
    dist = torch.empty(groups * n, device=x.device, dtype=F32)
    row_loss = torch.empty(groups * n, device=x.device, dtype=F32)
    loss = torch.empty((), device=x.device, dtype=F32)
    LIB.call("apla_koleo_fwd", ptr(xn), groups, n, D, float(eps), float(weight), ptr(nn), ptr(dist), ptr(row_loss),
             stream())
    LIB.call("apla_sum_f32", ptr(row_loss), groups * n, 1.0, ptr(loss), stream())
    return loss, xn, nn, dist


def koleo_bwd(x, xn, nn, dist, eps: float, gscale: Optional[torch.Tensor], groups: int = 1, weight: float = 1.0):
    require_device()
    n, D = x.shape[0] // groups, x.shape[1]
    dx = torch.empty_like(x)
    LIB.call("apla_koleo_bwd", ptr(x), ptr(xn), groups, n, D, float(eps), float(eps), float(weight), ptr(nn), ptr(dist),
             ptr(gscale), ptr(dx), stream())
    return dx


def ema_check(teacher: torch.Tensor, student: torch.Tensor) -> None:
    """The argument checks of `ema_update_` alone (a pair the caller skips must still be one the kernel would accept:
    no silent CPU path)."""
    require_device()
    for t, name in ((teacher, "teacher"), (student, "student")):
        if not t.is_cuda or t.dtype != F32 or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous f32 CUDA tensor")
    if teacher.numel() != student.numel():
        raise RuntimeError("teacher and student differ in size")


def ema_update_(teacher: torch.Tensor, student: torch.Tensor, m: float) -> torch.Tensor:
    """teacher <- m * teacher + (1 - m) * student, in place (f32, contiguous, same number of elements)."""
    ema_check(teacher, student)
    LIB.call("apla_ema_update", ptr(teacher), ptr(student), teacher.numel(), float(m), stream())
    return teacher


def ssl_objective(s_scores: torch.Tensor, t_scores: torch.Tensor, dino_center: torch.Tensor, ibot_center: torch.Tensor,
                  masks_weight: Optional[torch.Tensor], B: int, n_local: int, teacher_temp: float,
                  student_temp: float = 0.1, dino_weight: float = 1.0, ibot_weight: float = 1.0,
                  gscale: Optional[torch.Tensor] = None, ds_dtype=BF16, need_grad: bool = True):
    """`apla_ssl_objective`: the whole objective of one step on given head outputs in one native launch sequence.
    -> dict(losses [3] = dino_local, dino_global, 2 * ibot_loss; ds [rows, K] or None; t_probs; dino_batch_sum [1, K];
    ibot_batch_mean [1, 1, K]).  Row layout of s_scores / t_scores: include/apla_b200.h."""
    require_device()
    s, t = _rows(s_scores, "s_scores"), _rows(t_scores, "t_scores")
    rows, K = s.shape
    n_masked = t.shape[0] - 2 * B
    if n_masked < 0 or rows != n_local * B + 2 * B + n_masked or t.shape[1] != K:
        raise RuntimeError(f"ssl_objective: {tuple(s.shape)} student rows / {tuple(t.shape)} teacher rows do not fit "
                           f"B={B}, n_local={n_local}")
    for c, name in ((dino_center, "dino_center"), (ibot_center, "ibot_center")):
        if c.numel() != K or c.dtype != F32 or not c.is_contiguous() or not c.is_cuda:
            raise RuntimeError(f"{name} must hold {K} contiguous f32 values on the device")
    if n_masked > 0 and (masks_weight is None or masks_weight.numel() != n_masked or masks_weight.dtype != F32
                         or not masks_weight.is_contiguous()):
        raise RuntimeError("masks_weight must hold one contiguous f32 weight per masked patch")
    dev = s.device
    splits = max(1, min(32, max(2 * B, n_masked) // 64))
    t_probs = torch.empty_like(t)
    row_ws = torch.empty(3 * rows, device=dev, dtype=F32)
    col_ws = torch.empty(splits, K, device=dev, dtype=F32)
    losses = torch.empty(3, device=dev, dtype=F32)
    dsum, imean = torch.empty(1, K, device=dev, dtype=F32), torch.empty(1, 1, K, device=dev, dtype=F32)
    ds = torch.empty(rows, K, device=dev, dtype=ds_dtype) if need_grad else None
    LIB.call("apla_ssl_objective", ptr(s), s.stride(0), ptr(t), t.stride(0), ptr(t_probs), t_probs.stride(0),
             ptr(dino_center), ptr(ibot_center), ptr(masks_weight), B, n_local, n_masked, K, float(teacher_temp),
             float(student_temp), float(dino_weight), float(ibot_weight), ptr(row_ws), ptr(col_ws), splits, ptr(ds),
             K if ds is None else ds.stride(0), int(ds_dtype == BF16), ptr(gscale), ptr(losses), ptr(dsum), ptr(imean),
             stream())
    return dict(losses=losses, ds=ds, t_probs=t_probs, dino_batch_sum=dsum, ibot_batch_mean=imean)
