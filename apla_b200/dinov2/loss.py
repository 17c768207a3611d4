"""Host-side mirror of the reference's self-supervised losses (src/self_supervised/dinov2/loss/*.py) on the sm_100a row
kernels of csrc/ssl.cu: same class names, constructor arguments, attributes, buffers and method names, so that
`DINOv2.__init__` / `DINOv2.forward` (src/self_supervised/dinov2/models.py:107-124, 207-433) can take them unchanged.

  DINOLoss       loss/dino_clstoken_loss.py:12-98    softmax_center_teacher, forward, update_center, apply_center_update
  iBOTPatchLoss  loss/ibot_patch_loss.py:28-145      softmax_center_teacher, forward, forward_masked, update_center, ...
  KoLeoLoss      loss/koleo_loss.py:17-45            forward
  update_teacher models.py:437-447                   EMA of the teacher

Differences from the reference, all deliberate:
  * inputs must be fp32 CUDA tensors with contiguous rows (the reference also accepts fp16 under autocast); anything else
    raises -- there is no PyTorch fallback;
  * `center` is updated IN PLACE (the reference rebinds the buffer to a new tensor every step);
  * the [rows, K] log-softmax and the teacher / student product are never materialised: one kernel per loss call reads the
    scores once for the forward (per-row log-sum-exp kept) and once for the backward;
  * `sinkhorn_knopp_teacher` skips the reference's first division by the total mass (it cancels in the first column
    normalisation); `int(B)` of the iBOT variant is one host read of a one-element tensor, as in the reference's `Q /= B`.
Tested against oracle/ssl_oracle.py (itself pinned to the reference) in tests/test_ssl_gpu.py.
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn

from . import ops


class _SoftCE(torch.autograd.Function):
    """loss = -sum_rows w_row * sum_k q_k log_softmax(s * inv_temp)_k,  q = t0[row % t_rows] (+ t1[row % t_rows])."""

    @staticmethod
    def forward(ctx, s, t0, t1, w_row, w_uniform, inv_temp, t_rows):
        loss, lse, mass = ops.soft_ce_fwd(s, t0, t1, t_rows, w_row, w_uniform, inv_temp)
        ctx.save_for_backward(s, t0, t1, w_row, lse, mass)
        ctx.cfg = (w_uniform, inv_temp, t_rows)
        return loss

    @staticmethod
    def backward(ctx, g):
        s, t0, t1, w_row, lse, mass = ctx.saved_tensors
        w_uniform, inv_temp, t_rows = ctx.cfg
        ds = ops.soft_ce_bwd(s, t0, t1, t_rows, w_row, w_uniform, inv_temp, lse, mass, g.contiguous())
        return ds, None, None, None, None, None, None


def _as_one_matrix(chunks: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
    """If `chunks` are consecutive equal-height row blocks of one buffer (what `x.chunk(n)` yields), the [n*b, K] matrix
    over all of them as a VIEW (autograd-connected through torch.cat otherwise would copy 65 536-wide rows)."""
    base = getattr(chunks[0], "_base", None)
    if base is None or base.dim() != 2 or base.stride(1) != 1:
        return None
    b, K = chunks[0].shape
    off = chunks[0].storage_offset()
    for i, c in enumerate(chunks):
        if getattr(c, "_base", None) is not base or c.shape != (b, K) or c.stride() != base.stride() \
                or c.storage_offset() != off + i * b * base.stride(0):
            return None
    first = (off - base.storage_offset()) // base.stride(0)
    return base[first:first + len(chunks) * b]


class DINOLoss(nn.Module):
    def __init__(self, out_dim, student_temp=0.1, center_momentum=0.9):
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.updated = True
        self.reduce_handle = None
        self.len_teacher_output = None
        self.async_batch_center = None

    @torch.no_grad()
    def softmax_center_teacher(self, teacher_output, teacher_temp):
        self.apply_center_update()
        return ops.softmax_center(teacher_output, self.center, teacher_temp)

    @torch.no_grad()
    def sinkhorn_knopp_teacher(self, teacher_output, teacher_temp, n_iterations=3):
        world = dist.get_world_size() if dist.is_initialized() else 1
        return ops.sinkhorn_knopp(teacher_output, teacher_temp, n_iterations, teacher_output.shape[0] * world,
                                  dist.all_reduce if dist.is_initialized() else None)

    def forward(self, student_output_list, teacher_out_softmaxed_centered_list):
        """- sum over (student crop, teacher crop) pairs of mean_b sum_k t log_softmax(s / student_temp)."""
        students = list(student_output_list)
        teachers = list(teacher_out_softmaxed_centered_list)
        inv_temp = 1.0 / self.student_temp
        merged = _as_one_matrix(students) if len(students) > 1 else None
        groups = [merged] if merged is not None else students
        total = None
        for s in groups:
            for j in range(0, len(teachers), 2):                      # two teacher crops per launch
                t0 = teachers[j]
                t1 = teachers[j + 1] if j + 1 < len(teachers) else None
                b = t0.shape[0]
                if s.shape[0] % b:
                    raise RuntimeError(f"student rows {s.shape[0]} are not a multiple of the teacher's {b}")
                if t1 is not None and t1.stride(0) != t0.stride(0):
                    t1 = t1.contiguous()
                    t0 = t0.contiguous()
                term = _SoftCE.apply(s, t0, t1, None, 1.0 / b, inv_temp, b)
                total = term if total is None else total + term
        return total

    @torch.no_grad()
    def update_center(self, teacher_output):
        self.reduce_center_update(teacher_output)

    @torch.no_grad()
    def reduce_center_update(self, teacher_output):
        self.updated = False
        self.len_teacher_output = len(teacher_output)
        self.async_batch_center = ops.colsum(teacher_output)
        if dist.is_initialized():
            self.reduce_handle = dist.all_reduce(self.async_batch_center, async_op=True)

    @torch.no_grad()
    def register_center_stat(self, batch_sum, n_rows):
        """`reduce_center_update` for a statistic that is already computed (`apla_ssl_objective` returns the column sums of
        the teacher's CLS scores): same pending-update protocol, same asynchronous all-reduce."""
        self.updated = False
        self.len_teacher_output = n_rows
        self.async_batch_center = batch_sum.view(1, -1)
        if dist.is_initialized():
            self.reduce_handle = dist.all_reduce(self.async_batch_center, async_op=True)

    @torch.no_grad()
    def apply_center_update(self):
        if self.updated is False:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
            if self.reduce_handle is not None:
                self.reduce_handle.wait()
            ops.center_ema_(self.center, self.async_batch_center, self.len_teacher_output * world_size,
                            self.center_momentum)
            self.updated = True


class iBOTPatchLoss(nn.Module):
    def __init__(self, patch_out_dim, student_temp=0.1, center_momentum=0.9):
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.register_buffer("center", torch.zeros(1, 1, patch_out_dim))
        self.updated = True
        self.reduce_handle = None
        self.len_teacher_patch_tokens = None
        self.async_batch_center = None

    @torch.no_grad()
    def softmax_center_teacher(self, teacher_patch_tokens, teacher_temp):
        self.apply_center_update()
        K = teacher_patch_tokens.shape[-1]
        out = ops.softmax_center(teacher_patch_tokens.reshape(-1, K), self.center, teacher_temp)
        return out.view(teacher_patch_tokens.shape)

    @torch.no_grad()
    def sinkhorn_knopp_teacher(self, teacher_output, teacher_temp, n_masked_patches_tensor, n_iterations=3):
        """B = the number of masked patches over all ranks (ibot_patch_loss.py:58-59; the reference all-reduces it even
        outside a process group, which raises -- here a single process is world size 1)."""
        B = n_masked_patches_tensor
        if dist.is_initialized():
            dist.all_reduce(B)
        return ops.sinkhorn_knopp(teacher_output, teacher_temp, n_iterations, int(B),
                                  dist.all_reduce if dist.is_initialized() else None)

    def forward(self, student_patch_tokens, teacher_patch_tokens, student_masks_flat):
        """Dense form (B, N, K): - mean_b sum_n mask_bn sum_k t log_softmax(s / T) / max(sum_n mask_bn, 1)."""
        B, N, K = student_patch_tokens.shape
        m = student_masks_flat.float()
        w = (m / m.sum(dim=-1, keepdim=True).clamp(min=1.0)).reshape(-1).contiguous()
        return _SoftCE.apply(student_patch_tokens.reshape(B * N, K), teacher_patch_tokens.reshape(B * N, K), None, w,
                             1.0 / B, 1.0 / self.student_temp, B * N)

    def forward_masked(self, student_patch_tokens_masked, teacher_patch_tokens_masked, student_masks_flat,
                       n_masked_patches=None, masks_weight=None):
        s, t = student_patch_tokens_masked, teacher_patch_tokens_masked
        if masks_weight is None:
            masks_weight = (1 / student_masks_flat.sum(-1).clamp(min=1.0)).unsqueeze(-1) \
                .expand_as(student_masks_flat)[student_masks_flat]
        n = s.shape[0] if n_masked_patches is None else int(n_masked_patches)
        w = masks_weight.float().contiguous()
        if w.numel() != n:
            raise RuntimeError(f"masks_weight has {w.numel()} entries for {n} masked patches")
        if n == 0:                      # no masked patch in the batch: the reference's sum over an empty tensor
            return s[:0].sum() * 0.0   # (ibot_patch_loss.py:119-121); keeps the graph, zero gradient
        return _SoftCE.apply(s[:n], t[:n], None, w, 1.0 / student_masks_flat.shape[0], 1.0 / self.student_temp,
                             max(n, 1))

    @torch.no_grad()
    def update_center(self, teacher_patch_tokens):
        self.reduce_center_update(teacher_patch_tokens)

    @torch.no_grad()
    def reduce_center_update(self, teacher_patch_tokens):
        """teacher_patch_tokens [b, n, K]: the statistic is sum_b mean_n = the column sums over all b * n rows / n (taken
        BEFORE the all-reduce: ranks hold different numbers of masked patches and the reference averages per-rank means)."""
        self.updated = False
        b, n, K = teacher_patch_tokens.shape
        self.len_teacher_patch_tokens = b
        self.async_batch_center = ops.colsum(teacher_patch_tokens.reshape(b * n, K), 1.0 / max(n, 1)).view(1, 1, K)
        if dist.is_initialized():
            self.reduce_handle = dist.all_reduce(self.async_batch_center, async_op=True)

    @torch.no_grad()
    def register_center_stat(self, batch_mean, n_items=1):
        """`reduce_center_update` for an already computed statistic (`apla_ssl_objective`'s mean over the masked rows)."""
        self.updated = False
        self.len_teacher_patch_tokens = n_items
        self.async_batch_center = batch_mean.view(1, 1, -1)
        if dist.is_initialized():
            self.reduce_handle = dist.all_reduce(self.async_batch_center, async_op=True)

    @torch.no_grad()
    def apply_center_update(self):
        if self.updated is False:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
            if self.reduce_handle is not None:
                self.reduce_handle.wait()
            ops.center_ema_(self.center, self.async_batch_center,
                            self.len_teacher_patch_tokens * world_size, self.center_momentum)
            self.updated = True


class _KoLeo(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, groups):
        loss, xn, nn_idx, d = ops.koleo_fwd(x, eps, groups)
        ctx.save_for_backward(x, xn, nn_idx, d)
        ctx.eps, ctx.groups = eps, groups
        return loss

    @staticmethod
    def backward(ctx, g):
        x, xn, nn_idx, d = ctx.saved_tensors
        return ops.koleo_bwd(x, xn, nn_idx, d, ctx.eps, g.contiguous(), ctx.groups), None, None


class KoLeoLoss(nn.Module):
    """Kozachenko-Leonenko regulariser (loss/koleo_loss.py:17-45): -mean_i log(min_j ||x^_i - x^_j|| + eps)."""

    def __init__(self):
        super().__init__()

    def forward(self, student_output, eps=1e-8):
        return _KoLeo.apply(student_output.contiguous(), eps, 1)

    def forward_chunks(self, student_output, chunks, eps=1e-8):
        """sum(self(p) for p in student_output.chunk(chunks)) -- what DINOv2.forward computes over the two global crops
        (models.py:414-416) -- in one set of launches: every chunk is an independent group of rows."""
        if student_output.shape[0] % chunks:
            raise RuntimeError("forward_chunks needs equally sized chunks")
        return _KoLeo.apply(student_output.contiguous(), eps, chunks)


# id(student tensor) -> (weak references to that tensor and to its teacher twin, both versions at the time of the check,
# verdict).  Validated through the OBJECTS, not through addresses: the caching allocator hands the addresses of a freed
# model to the next one.  (Not a WeakKeyDictionary: it compares keys with ==, which is element-wise for tensors.)
_FROZEN_PAIR_IS_IDENTICAL = {}


def _frozen_and_identical(s: torch.Tensor, t: torch.Tensor) -> bool:
    """A student parameter that does not train never changes; if the teacher's copy holds the same values (it is loaded
    from the student, models.py:138), `m * t + (1 - m) * s` leaves it where it is, up to the rounding of that expression.
    Checked once per (student tensor, teacher tensor) pair and again whenever either version moved."""
    if s.requires_grad or s.shape != t.shape:
        return False
    key = id(s)
    hit = _FROZEN_PAIR_IS_IDENTICAL.get(key)
    if hit is not None and hit[0]() is s and hit[1]() is t and hit[2] == s._version and hit[3] == t._version:
        return hit[4]
    same = bool(torch.equal(s.data, t.data))
    _FROZEN_PAIR_IS_IDENTICAL[key] = (weakref.ref(s, lambda _, k=key: _FROZEN_PAIR_IS_IDENTICAL.pop(k, None)),
                                      weakref.ref(t), s._version, t._version, same)
    return same


@torch.no_grad()
def update_teacher(student_params: Sequence[torch.Tensor], teacher_params: Sequence[torch.Tensor], m: float) -> None:
    """models.py:437-447: teacher <- m * teacher + (1 - m) * student, tensor by tensor (one launch each).

    Under APLA adaptation all but the projection rows and the head are frozen in the student and identical in the teacher:
    those pairs are skipped (the reference multiplies and adds them back to the same values every step).  Skipping is not
    only 250 of ~300 launches and 85 % of the EMA's bytes at ViT-L: an untouched teacher tensor keeps its version, so the
    teacher's cached bf16 weight copies and its cached position table stay valid instead of being rebuilt every step
    (measured at ViT-L, 16 images: 2.6 ms of bicubic resize + 2 ms of weight re-casting per step).  A frozen student tensor
    whose teacher copy DIFFERS is still averaged, as in the reference."""
    student_params, teacher_params = list(student_params), list(teacher_params)
    if len(student_params) != len(teacher_params):
        raise RuntimeError("student and teacher parameter lists differ in length")
    for s, t in zip(student_params, teacher_params):
        if _frozen_and_identical(s, t):
            check = getattr(ops, "ema_check", None)
            if check is not None:
                check(t.data, s.data)
            continue
        ops.ema_update_(t.data, s.data, m)
        # the kernel writes through a raw pointer: tell autograd (and the bf16 working-set caches of APLA_Attention /
        # FusedAplaBlock, which key on (data_ptr, _version)) that the tensor changed
        torch.autograd.graph.increment_version(t)
