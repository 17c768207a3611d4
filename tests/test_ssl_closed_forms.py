"""The closed-form gradients that csrc/ssl.cu implements (soft-target cross-entropy, L2 normalisation, weight
normalisation, KoLeo), written out in numpy-style torch exactly as the kernels compute them and checked against autograd
over oracle/ssl_oracle.py.  CPU only: this pins the MATH of the kernels before they meet hardware; tests/test_ssl_gpu.py
pins the kernels themselves."""
import torch
import torch.nn.functional as F

from oracle import ssl_oracle as S

torch.manual_seed(0)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def soft_ce_closed_form(s, t0, t1, t_rows, w_row, w_uniform, inv_temp, g=1.0):
    """soft_ce_fwd_kernel / soft_ce_bwd_kernel."""
    rows = s.shape[0]
    tr = torch.arange(rows) % t_rows
    q = t0[tr] + (t1[tr] if t1 is not None else 0)
    z = s * inv_temp
    lse = torch.logsumexp(z, dim=-1)
    mass = q.sum(-1)
    w = w_uniform * (w_row if w_row is not None else torch.ones(rows))
    row_loss = -w * ((q * z).sum(-1) - mass * lse)
    ds = (-w * inv_temp * g)[:, None] * (q - mass[:, None] * torch.exp(z - lse[:, None]))
    return row_loss.sum(), ds


def test_soft_ce_is_the_dino_loss_over_all_crop_pairs():
    B, K, n_local = 3, 32, 4
    s = torch.randn(n_local * B, K, requires_grad=True)
    t = F.softmax(torch.randn(2, B, K) * 3, dim=-1)
    ref = S.dino_loss(s.chunk(n_local), list(t), student_temp=0.1)
    ref.backward()
    loss, ds = soft_ce_closed_form(s.detach(), t[0], t[1], B, None, 1.0 / B, 10.0)
    assert rel(loss, ref.detach()) < 1e-6 and rel(ds, s.grad) < 1e-5
    # global crops: one student matrix against the flattened (swapped) teacher targets
    s2 = torch.randn(2 * B, K, requires_grad=True)
    ref2 = S.dino_loss([s2], [t.flatten(0, 1)], student_temp=0.1)
    ref2.backward()
    loss2, ds2 = soft_ce_closed_form(s2.detach(), t.flatten(0, 1), None, 2 * B, None, 1.0 / (2 * B), 10.0, g=1.0)
    assert rel(loss2, ref2.detach()) < 1e-6 and rel(ds2, s2.grad) < 1e-5


def test_soft_ce_is_the_masked_ibot_loss():
    nimg, P, K = 4, 9, 32
    masks = torch.rand(nimg, P) < 0.4
    masks[1] = False
    n = int(masks.sum())
    s = torch.randn(n + 3, K, requires_grad=True)                    # padded to the collate's upper bound, sliced
                                                                     # to n before the loss (models.py:371)
    t = F.softmax(torch.randn(n, K) * 3, dim=-1)
    mw = S.masks_weight_of(masks)
    ref = S.ibot_loss_masked(s[:n], t, masks, n_masked_patches=n, masks_weight=mw, student_temp=0.1) * 0.7
    ref.backward()
    loss, ds = soft_ce_closed_form(s.detach()[:n], t, None, max(n, 1), mw, 1.0 / nimg, 10.0, g=0.7)
    assert rel(loss * 0.7, ref.detach()) < 1e-6
    assert rel(ds, s.grad[:n]) < 1e-5 and float(s.grad[n:].abs().sum()) == 0.0


def test_l2norm_backward_closed_form():
    x = torch.randn(5, 16, requires_grad=True)
    dy = torch.randn(5, 16)
    eps = 1e-12
    F.normalize(x, dim=-1, p=2, eps=eps).backward(dy)
    xd = x.detach()
    nrm = xd.norm(dim=-1, keepdim=True)
    inv = 1 / nrm.clamp(min=eps)
    coef = (xd * dy).sum(-1, keepdim=True) * inv ** 3                 # l2norm_bwd_kernel
    assert rel(dy * inv - xd * coef, x.grad) < 1e-5


def test_weightnorm_backward_closed_form():
    K, d = 7, 16
    g = (1 + 0.1 * torch.randn(K, 1)).requires_grad_(True)
    v = torch.randn(K, d, requires_grad=True)
    dW = torch.randn(K, d)
    S.weight_norm_weight(g, v).backward(dW)
    vd, gd = v.detach(), g.detach()
    inv = 1 / vd.norm(dim=1, keepdim=True)
    vdot = (vd * dW).sum(1, keepdim=True)
    assert rel(vdot * inv, g.grad) < 1e-5                              # weightnorm_bwd_kernel: dg
    assert rel(gd * inv * (dW - vd * vdot * inv * inv), v.grad) < 1e-5   # dv


def test_koleo_closed_form():
    n, D, eps = 9, 12, 1e-8
    x = torch.randn(n, D, requires_grad=True)
    ref = S.koleo_loss(x, eps) * 0.3
    ref.backward()
    xd = x.detach()
    nrm = xd.norm(dim=-1, keepdim=True)
    inv = 1 / nrm.clamp(min=eps)
    xn = xd * inv
    dots = (xn @ xn.t()).masked_fill(torch.eye(n, dtype=torch.bool), -1.0)
    nn_idx = dots.argmax(1)
    u = xn - xn[nn_idx] + 1e-8
    dist = u.norm(dim=-1)                                              # koleo_nn_kernel
    loss = (-torch.log(dist + eps) / n).sum()
    assert rel(loss * 0.3, ref.detach()) < 1e-6
    up = -(1.0 / n) * 0.3                                              # koleo_bwd_kernel
    c = up / ((dist + eps) * dist)
    gi = c[:, None] * u
    for j in range(n):
        gi[nn_idx[j]] -= c[j] * (xn[j] - xn[nn_idx[j]] + 1e-8)
    coef = (xd * gi).sum(-1, keepdim=True) * inv ** 3
    assert rel(gi * inv - xd * coef, x.grad) < 1e-4


def test_center_updates_closed_form():
    K = 16
    t = torch.randn(6, K)
    c = torch.randn(1, K)
    assert rel(c * 0.9 + t.sum(0, keepdim=True) * (1 / (6 * 2)) * 0.1, S.dino_center_update(c, t, 0.9, world_size=2)) < 1e-6
    tp = torch.randn(1, 5, K)
    ci = torch.randn(1, 1, K)
    stat = tp.reshape(5, K).sum(0) * (1 / 5)                           # colsum with scale 1/n, then count = len * world
    assert rel(ci * 0.9 + stat * (1 / 1) * 0.1, S.ibot_center_update(ci, tp, 0.9)) < 1e-6
