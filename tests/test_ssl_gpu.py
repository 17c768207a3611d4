"""Parity of the self-supervised row kernels (csrc/ssl.cu) and of the loss classes built on them
(apla_b200/dinov2/loss.py) against oracle/ssl_oracle.py, which is pinned to the reference.  Everything is fp32; the bars
(1e-4 forward, 1e-3 gradients unless stated) cover the fast-math exp / log of the kernels.

Every test here is a plain (strict) GPU test: a failure fails the run."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import ssl_oracle as S

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

DEV = "cuda"


def _dinov2():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import apla_b200.dinov2 as D
    from apla_b200.dinov2 import ops
    return D, ops


def rel(a, b):
    a = a.detach().double().flatten().cpu(); b = b.detach().double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize("n,K,temp", [(37, 4096, 0.07), (5, 65536, 0.04), (1, 64, 0.04), (3, 4, 0.05), (2, 2052, 0.05)])
def test_softmax_center(n, K, temp):
    D, ops = _dinov2()
    t, c = gen(n, K, seed=1), gen(1, K, seed=2, scale=0.3)
    out = ops.softmax_center(t.to(DEV), c.to(DEV), temp)
    ref = S.softmax_center_teacher(t, c, temp)
    assert rel(out, ref) < 1e-4
    assert float((out.sum(-1) - 1).abs().max()) < 1e-4


@pytest.mark.parametrize("n,K", [(6, 64), (128, 4096), (1000, 65536)])
def test_center_update(n, K):
    D, ops = _dinov2()
    t, c = gen(n, K, seed=3), gen(1, K, seed=4)
    cs = ops.colsum(t.to(DEV))
    assert rel(cs, t.double().sum(0, keepdim=True)) < 1e-5
    cd = c.to(DEV).clone()
    ops.center_ema_(cd, cs, n * 2, 0.9)
    assert rel(cd, S.dino_center_update(c, t, 0.9, world_size=2)) < 1e-5
    ci = c.view(1, 1, K).to(DEV).clone()
    ops.center_ema_(ci, ops.colsum(t.to(DEV), 1.0 / n), 1, 0.9)
    assert rel(ci, S.ibot_center_update(c.view(1, 1, K), t.view(1, n, K), 0.9)) < 1e-5


@pytest.mark.parametrize("B,K,n_local,merged", [(3, 64, 4, True), (3, 64, 4, False), (8, 65536, 8, True), (1, 4, 2, True),
                                                 (2, 2052, 3, False)])
def test_dino_loss_local_and_global(B, K, n_local, merged):
    D, ops = _dinov2()
    loss_mod = D.DINOLoss(K).to(DEV)
    t = F.softmax(gen(2, B, K, seed=5) * 3, dim=-1)
    s_local, s_global = gen(n_local * B, K, seed=6), gen(2 * B, K, seed=7)
    # oracle
    a = s_local.clone().requires_grad_(True); b = s_global.clone().requires_grad_(True)
    ref = S.dino_loss(a.chunk(n_local), list(t), 0.1) * 0.25 + S.dino_loss([b], [t.flatten(0, 1)], 0.1) * 0.5
    ref.backward()
    # kernels
    x = s_local.to(DEV).requires_grad_(True); y = s_global.to(DEV).requires_grad_(True)
    td = t.to(DEV)
    chunks = x.chunk(n_local) if merged else [c.clone() for c in x.chunk(n_local)]
    out = loss_mod(chunks, list(td)) * 0.25 + loss_mod([y], [td.flatten(0, 1)]) * 0.5
    out.backward()
    assert rel(out, ref) < 1e-4
    assert rel(x.grad, a.grad) < 1e-3 and rel(y.grad, b.grad) < 1e-3


@pytest.mark.parametrize("nimg,P,K,pad", [(4, 16, 64, 3), (6, 256, 65536, 5), (4, 16, 64, 0), (2, 4, 4, 1), (3, 9, 2052, 0)])
def test_ibot_forward_masked(nimg, P, K, pad):
    D, ops = _dinov2()
    loss_mod = D.iBOTPatchLoss(K).to(DEV)
    g = torch.Generator().manual_seed(8)
    masks = torch.rand(nimg, P, generator=g) < 0.1
    masks[0, :2] = True
    masks[1] = False
    n = int(masks.sum())
    s, t = gen(n + pad, K, seed=9), F.softmax(gen(n, K, seed=10) * 3, dim=-1)
    mw = S.masks_weight_of(masks)
    a = s.clone().requires_grad_(True)
    ref = S.ibot_loss_masked(a[:n], t, masks, n_masked_patches=n, masks_weight=mw)
    ref.backward()
    x = s.to(DEV).requires_grad_(True)
    out = loss_mod.forward_masked(x[:n], t.to(DEV), student_masks_flat=masks.to(DEV), n_masked_patches=n,
                                  masks_weight=mw.to(DEV))
    out.backward()
    assert rel(out, ref) < 1e-4 and rel(x.grad, a.grad) < 1e-3
    # masks_weight derived from the masks when it is not passed (ibot_patch_loss.py:113-118)
    out2 = loss_mod.forward_masked(x.detach()[:n], t.to(DEV), student_masks_flat=masks.to(DEV))
    assert rel(out2, ref) < 1e-4


def test_ibot_no_masked_patches():
    D, ops = _dinov2()
    loss_mod = D.iBOTPatchLoss(64).to(DEV)
    masks = torch.zeros(4, 16, dtype=torch.bool, device=DEV)
    x = torch.randn(3, 64, device=DEV, requires_grad=True)
    out = loss_mod.forward_masked(x[:0], torch.zeros(0, 64, device=DEV), masks, n_masked_patches=0,
                                  masks_weight=torch.zeros(0, device=DEV))
    out.backward()
    assert float(out) == 0.0 and float(x.grad.abs().sum()) == 0.0


def test_ibot_dense_forward():
    D, ops = _dinov2()
    B, N, K = 3, 10, 128
    loss_mod = D.iBOTPatchLoss(K).to(DEV)
    s, t = gen(B, N, K, seed=11), F.softmax(gen(B, N, K, seed=12) * 3, dim=-1)
    masks = torch.rand(B, N, generator=torch.Generator().manual_seed(13)) < 0.4
    a = s.clone().requires_grad_(True)
    l = torch.sum(t * F.log_softmax(a / 0.1, dim=-1), dim=-1)                      # ibot_patch_loss.py:84-100
    ref = -(torch.sum(l * masks.float(), dim=-1) / masks.sum(dim=-1).clamp(min=1.0)).mean()
    ref.backward()
    x = s.to(DEV).requires_grad_(True)
    out = loss_mod(x, t.to(DEV), masks.to(DEV))
    out.backward()
    assert rel(out, ref) < 1e-4 and rel(x.grad, a.grad) < 1e-3


@pytest.mark.parametrize("n,Dm", [(9, 12), (64, 1024), (2, 64), (1, 64)])
def test_koleo(n, Dm):
    D, ops = _dinov2()
    x = gen(n, Dm, seed=14)
    a = x.clone().requires_grad_(True)
    ref = S.koleo_loss(a) * 0.1
    ref.backward()
    xd = x.to(DEV).requires_grad_(True)
    out = D.KoLeoLoss()(xd) * 0.1
    out.backward()
    assert rel(out, ref) < 1e-4
    if n == 1:                                  # a lone row is its own neighbour (koleo_loss.py:30-33): constant loss, zero gradient
        assert float(a.grad.abs().max()) == 0.0 and float(xd.grad.abs().max()) < 1e-6
    else:
        assert rel(xd.grad, a.grad) < 1e-3


@pytest.mark.parametrize("n,K,temp,iters", [(12, 128, 0.04, 3), (37, 512, 0.07, 1), (5, 65536, 0.04, 3), (2, 4, 0.05, 2),
                                            (3, 2052, 0.05, 3)])
def test_sinkhorn_knopp_teacher(n, K, temp, iters):
    """DINOLoss / iBOTPatchLoss.sinkhorn_knopp_teacher against the oracle (pinned to the reference by ssl_sk_small)."""
    D, ops = _dinov2()
    t = gen(n, K, seed=80, scale=0.3)
    ref = S.sinkhorn_knopp(t, temp, iters)
    out = D.DINOLoss(K).to(DEV).sinkhorn_knopp_teacher(t.to(DEV), temp, n_iterations=iters)
    assert rel(out, ref) < 1e-4 and float((out.sum(-1) - 1).abs().max()) < 1e-4
    out_i = D.iBOTPatchLoss(K).to(DEV).sinkhorn_knopp_teacher(
        t.to(DEV), temp, n_masked_patches_tensor=torch.full((1,), n, dtype=torch.long), n_iterations=iters)
    assert rel(out_i, S.sinkhorn_knopp(t, temp, iters, n_samples_world=n)) < 1e-4


def test_koleo_chunks():
    """KoLeoLoss.forward_chunks(x, 2) == sum of the loss over x.chunk(2) (models.py:414-416), gradients included."""
    D, ops = _dinov2()
    x = gen(2 * 7, 40, seed=22)
    a = x.clone().requires_grad_(True)
    ref = sum(S.koleo_loss(p) for p in a.chunk(2)) * 0.1
    ref.backward()
    xd = x.to(DEV).requires_grad_(True)
    out = D.KoLeoLoss().forward_chunks(xd, 2) * 0.1
    out.backward()
    assert rel(out, ref) < 1e-4 and rel(xd.grad, a.grad) < 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_l2norm(dtype):
    D, ops = _dinov2()
    x = gen(100, 256, seed=15).to(dtype)
    dy = gen(100, 256, seed=16).to(dtype)
    a = x.float().requires_grad_(True)
    y_ref = F.normalize(a, dim=-1, p=2, eps=1e-12)
    y_ref.backward(dy.float())
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    assert rel(ops.l2norm_fwd(x.to(DEV), 1e-12, torch.float32), y_ref) < 1e-5
    assert rel(ops.l2norm_fwd(x.to(DEV), 1e-12, torch.bfloat16).float(), y_ref) < 6e-3
    assert rel(ops.l2norm_bwd(x.to(DEV), dy.to(DEV), 1e-12).float(), a.grad) < tol


def test_weightnorm():
    D, ops = _dinov2()
    K, d = 4096, 256
    g = (1 + 0.1 * gen(K, 1, seed=17)); v = gen(K, d, seed=18); dW = gen(K, d, seed=19)
    a, b = g.clone().requires_grad_(True), v.clone().requires_grad_(True)
    w_ref = S.weight_norm_weight(a, b)
    w_ref.backward(dW)
    assert rel(ops.weightnorm_fwd(g.to(DEV), v.to(DEV), torch.float32), w_ref) < 1e-5
    assert rel(ops.weightnorm_fwd(g.to(DEV), v.to(DEV), torch.bfloat16).float(), w_ref) < 6e-3
    dg, dv = ops.weightnorm_bwd(g.to(DEV), v.to(DEV), dW.to(DEV))
    assert dg.shape == g.shape and rel(dg, a.grad) < 1e-4 and rel(dv, b.grad) < 1e-4
    dg2, dv2 = ops.weightnorm_bwd(g.to(DEV), v.to(DEV), dW.to(DEV), need_dg=False)
    assert dg2 is None and torch.equal(dv2, dv)


@pytest.mark.parametrize("n", [1, 7, 4096, 1 << 20])
def test_update_teacher(n):
    D, ops = _dinov2()
    s, t = gen(n, seed=20), gen(n, seed=21)
    sd, td = {"w": s.clone()}, {"w": t.clone()}
    S.ema_update(td, sd, 0.994)
    sp, tp = torch.nn.Parameter(s.to(DEV)), torch.nn.Parameter(t.to(DEV), requires_grad=False)
    D.update_teacher([sp], [tp], 0.994)
    assert rel(tp, td["w"]) < 1e-6 and torch.equal(sp.detach().cpu(), s)


def test_update_teacher_refreshes_cached_weights():
    """The EMA kernel writes the teacher through raw pointers; the bf16 working sets of APLA_Attention / FusedAplaBlock are
    keyed on (data_ptr, _version).  After update_teacher the teacher backbone must compute with the NEW projection rows:
    its output has to change and to equal that of a freshly built (un-cached) copy holding the same parameters."""
    D, ops = _dinov2()
    from apla_b200.config import AplaConfig
    from apla_b200.hostdino import build_dino_backbone
    from apla_b200.hostvit import VitArch
    arch = VitArch(128, 2, 2)
    torch.manual_seed(11)
    student = build_dino_backbone(arch, img_size=56, patch_size=14, apla_config=AplaConfig(16))
    inds = [b.attn.inds.clone() for b in student.blocks]
    teacher = build_dino_backbone(arch, img_size=56, patch_size=14, apla_config=AplaConfig(16), indices=inds)
    teacher.load_state_dict(student.state_dict())
    with torch.no_grad():                                   # a student that has moved away from the teacher
        for n, p in student.named_parameters():
            if p.requires_grad:
                p.add_(0.05 * torch.randn_like(p))
    student, teacher = student.to(DEV), teacher.to(DEV)
    x = gen(3, 3, 56, 56, seed=90).to(DEV)
    with torch.no_grad():
        before = teacher(x, is_training=True)["x_norm_clstoken"].clone()
        D.update_teacher(list(student.parameters()), list(teacher.parameters()), 0.5)
        after = teacher(x, is_training=True)["x_norm_clstoken"].clone()
        fresh = build_dino_backbone(arch, img_size=56, patch_size=14, apla_config=AplaConfig(16), indices=inds)
        fresh.load_state_dict({k: v.cpu() for k, v in teacher.state_dict().items()})
        want = fresh.to(DEV)(x, is_training=True)["x_norm_clstoken"]
    assert rel(after, before) > 1e-3, "teacher output did not move: stale cached weights"
    assert rel(after, want) < 1e-5


def test_centre_protocol_over_two_steps():
    """softmax_center_teacher applies the update left pending by the previous step (dino_clstoken_loss.py:28-31,88-98)."""
    D, ops = _dinov2()
    K, n = 256, 12
    dl, il = D.DINOLoss(K).to(DEV), D.iBOTPatchLoss(K).to(DEV)
    dc, ic = torch.zeros(1, K), torch.zeros(1, 1, K)
    for step in range(2):
        t_cls, t_patch = gen(n, K, seed=30 + step), gen(1, 2 * n, K, seed=40 + step)
        got_d = dl.softmax_center_teacher(t_cls.to(DEV), 0.05)
        dl.update_center(t_cls.to(DEV))
        got_i = il.softmax_center_teacher(t_patch.to(DEV), 0.05)
        il.update_center(t_patch.to(DEV))
        assert rel(got_d, S.softmax_center_teacher(t_cls, dc, 0.05)) < 1e-4
        assert rel(got_i, S.softmax_center_teacher(t_patch, ic, 0.05)) < 1e-4
        dc, ic = S.dino_center_update(dc, t_cls), S.ibot_center_update(ic, t_patch)
    dl.apply_center_update(); il.apply_center_update()
    assert rel(dl.center, dc) < 1e-5 and rel(il.center, ic) < 1e-5
    assert dl.center.shape == (1, K) and il.center.shape == (1, 1, K)


def test_objective_matches_oracle_assembly():
    """The loss classes wired exactly as DINOv2.forward wires them (models.py:374-433) against ssl_objective, on given
    head outputs (the head GEMMs are not part of this file)."""
    D, ops = _dinov2()
    B, K, P, n_local, Dm = 4, 512, 16, 8, 64
    g = torch.Generator().manual_seed(50)
    masks = torch.rand(2 * B, P, generator=g) < 0.3
    masks[0, 0] = True
    idx, mw = S.mask_indices_of(masks), S.masks_weight_of(masks)
    n = idx.shape[0]
    s_local, s_global, s_patch = gen(n_local * B, K, seed=51), gen(2 * B, K, seed=52), gen(n, K, seed=53)
    t_cls, t_patch = gen(2 * B, K, seed=54), gen(n, K, seed=55)
    cls_feat = gen(2 * B, Dm, seed=56)
    terms = 2 + n_local * 2

    def run(dev, dino, ibot, koleo, sm_d, sm_i):
        # (fresh leaves per run: `.to("cpu")` would hand back the shared CPU tensor itself)
        a, b, c = (x.detach().clone().to(dev).requires_grad_(True) for x in (s_local, s_global, s_patch))
        f = cls_feat.detach().clone().to(dev).requires_grad_(True)
        t_d = sm_d(t_cls.to(dev)).view(2, B, K)
        t_i = sm_i(t_patch.to(dev).unsqueeze(0)).squeeze(0)
        total = dino(a.chunk(n_local), list(t_d)) / terms
        total = total + dino([b], [t_d.flatten(0, 1)]) * 2 / terms
        total = total + 0.1 * sum(koleo(p) for p in f.chunk(2))
        total = total + ibot(c, t_i, masks.to(dev), n, mw.to(dev)) * 2 * 0.5
        total.backward()
        return total, a.grad, b.grad, c.grad, f.grad

    ref = run("cpu", lambda s, t: S.dino_loss(s, t), lambda s, t, m, n_, w: S.ibot_loss_masked(s, t, m, n_, w),
              S.koleo_loss, lambda t: S.softmax_center_teacher(t, torch.zeros(1, K), 0.05),
              lambda t: S.softmax_center_teacher(t, torch.zeros(1, 1, K), 0.05))
    dl, il, kl = D.DINOLoss(K).to(DEV), D.iBOTPatchLoss(K).to(DEV), D.KoLeoLoss()
    got = run(DEV, dl, lambda s, t, m, n_, w: il.forward_masked(s, t, m, n_masked_patches=n_, masks_weight=w), kl,
              lambda t: dl.softmax_center_teacher(t, 0.05), lambda t: il.softmax_center_teacher(t, 0.05))
    assert rel(got[0], ref[0]) < 1e-4
    for gg, rr in zip(got[1:], ref[1:]):
        assert rel(gg, rr) < 1e-3


@pytest.mark.parametrize("n,in_dim,hidden,bott,K,nlayers,bias", [(70, 64, 128, 64, 256, 3, True), (33, 128, 128, 64, 512, 1, True),
                                                              (200, 1024, 2048, 256, 4096, 3, True),
                                                              (70, 64, 128, 64, 256, 2, False)])
def test_dino_head(n, in_dim, hidden, bott, K, nlayers, bias):
    """DINOHead as one autograd node against the oracle's fp32 head on the same weights.  bf16 GEMM operands: relative
    error <= 1e-2 on the scores and cosine >= 0.999 / relative error <= 3e-2 on every gradient."""
    D, ops = _dinov2()
    torch.manual_seed(3)
    head = D.DINOHead(in_dim, K, nlayers=nlayers, hidden_dim=hidden, bottleneck_dim=bott, mlp_bias=bias)
    with torch.no_grad():                                           # trunc_normal(0.02) leaves the head nearly linear
        for name, p in head.named_parameters():
            if name.endswith("weight"):
                p.mul_(4.0)
            elif name.endswith("bias"):
                p.normal_(0, 0.1)
            elif name.endswith("weight_g"):
                p.add_(0.1 * torch.randn_like(p))
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in head.state_dict().items()}
    x = gen(n, in_dim, seed=60)
    dy = gen(n, K, seed=61) * 0.01
    xr = x.clone().requires_grad_(True)
    ref = S.dino_head_forward(sd, xr)
    ref.backward(dy)
    head = head.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    out = head(xd)
    out.backward(dy.to(DEV))
    assert out.dtype == torch.float32 and rel(out, ref) < 1e-2
    got = dict(head.named_parameters())
    for k, v in sd.items():
        gk = got[k].grad
        assert gk is not None and gk.shape == v.shape, k
        cos = float(torch.nn.functional.cosine_similarity(gk.detach().double().flatten().cpu(), v.grad.double().flatten(), dim=0))
        assert cos > 0.999 and rel(gk, v.grad) < 3e-2, (k, cos, rel(gk, v.grad))
    assert rel(xd.grad, xr.grad) < 3e-2
    assert list(head.state_dict().keys()) == list(sd.keys())


@pytest.mark.parametrize("B,K,n_local,P,empty", [(3, 64, 4, 16, False), (4, 512, 8, 16, False), (2, 64, 0, 16, False),
                                                 (3, 64, 2, 16, True), (1, 4, 1, 4, False), (2, 2052, 2, 9, False)])
def test_native_objective_sequence(B, K, n_local, P, empty):
    """`apla_ssl_objective` (one native launch sequence) against the oracle's `ssl_objective` on given head outputs: the
    three loss terms, the gradient with respect to every student score and both centre updates."""
    D, ops = _dinov2()
    g = torch.Generator().manual_seed(70)
    masks = torch.zeros(2 * B, P, dtype=torch.bool) if empty else torch.rand(2 * B, P, generator=g) < 0.3
    if not empty:
        masks[0, 0] = True
    mw = S.masks_weight_of(masks)
    n = int(masks.sum())
    s_all = gen(n_local * B + 2 * B + n, K, seed=71)
    t_all = gen(2 * B + n, K, seed=72)
    dc, ic = gen(1, K, seed=73, scale=0.2), gen(1, 1, K, seed=74, scale=0.2)
    dino_w, ibot_w, temp = 0.8, 1.3, 0.05
    # oracle: the same arithmetic as ssl_objective from the head outputs on (KoLeo acts on backbone features, not here)
    a = s_all.clone().requires_grad_(True)
    s_local, s_global, s_patch = a[:n_local * B], a[n_local * B:n_local * B + 2 * B], a[n_local * B + 2 * B:]
    t_d = S.softmax_center_teacher(t_all[:2 * B], dc, temp).view(2, B, K)
    t_i = S.softmax_center_teacher(t_all[2 * B:].unsqueeze(0), ic, temp).squeeze(0)
    terms = 2 + max(2 * n_local, 1)
    l = S.dino_loss(s_local.chunk(n_local), list(t_d)) / terms if n_local else torch.zeros(())
    gl = S.dino_loss([s_global], [t_d.flatten(0, 1)]) * 2 / terms
    ib = S.ibot_loss_masked(s_patch, t_i, masks, n_masked_patches=n, masks_weight=mw) * 2 * 0.5
    (dino_w * (l + gl) + ibot_w * ib).backward()
    up = torch.tensor(0.5)
    res = ops.ssl_objective(s_all.to(DEV), t_all.to(DEV), dc.to(DEV), ic.to(DEV), mw.to(DEV), B, n_local, temp,
                            dino_weight=dino_w, ibot_weight=ibot_w, gscale=up.to(DEV), ds_dtype=torch.float32)
    want = torch.stack([l.detach(), gl.detach(), ib.detach()])
    assert rel(res["losses"], want) < 1e-4, (res["losses"], want)
    assert rel(res["ds"], a.grad * 0.5) < 1e-3
    assert rel(res["t_probs"][:2 * B], t_d.flatten(0, 1)) < 1e-4
    new_dc = dc.to(DEV).clone(); new_ic = ic.to(DEV).clone()
    ops.center_ema_(new_dc, res["dino_batch_sum"], 2 * B, 0.9)
    ops.center_ema_(new_ic, res["ibot_batch_mean"], 1, 0.9)
    assert rel(new_dc, S.dino_center_update(dc, t_all[:2 * B], 0.9)) < 1e-5
    if n:
        assert rel(new_ic, S.ibot_center_update(ic, t_all[2 * B:].unsqueeze(0), 0.9)) < 1e-5
    bf = ops.ssl_objective(s_all.to(DEV), t_all.to(DEV), dc.to(DEV), ic.to(DEV), mw.to(DEV), B, n_local, temp,
                           dino_weight=dino_w, ibot_weight=ibot_w)
    assert bf["ds"].dtype == torch.bfloat16 and rel(bf["ds"].float(), a.grad) < 6e-3


@pytest.mark.parametrize("fused", [False, True], ids=["per-term-losses", "fused-head-objective"])
def test_ssl_step_against_reference_vectors(fused):
    """Two whole self-supervised steps -- fused multi-crop student / teacher backbones (apla_b200.apla), DINOHead, the
    three losses, teacher EMA, centre updates -- against the vectors recorded from the reference's unmodified DINOv2
    meta-architecture (tests/golden/make_golden_ssl_step.py).  Bars: tests/helpers.py SSL_BF16_BARS."""
    D, ops = _dinov2()
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import helpers
    from apla_b200.config import AplaConfig
    from apla_b200.hostdino import SSLMetaArch, build_dino_backbone
    from apla_b200.hostvit import VitArch
    cfg, student, teacher, trainable, batch, arr = helpers.ssl_step_case()
    arch = VitArch(cfg["embed_dim"], cfg["depth"], cfg["num_heads"])

    def make(sd):
        inds = [sd[f"backbone.blocks.{i}.attn.inds"] for i in range(cfg["depth"])]
        bb = build_dino_backbone(arch, img_size=cfg["global_px"], patch_size=cfg["patch"],
                                 apla_config=AplaConfig(cfg["partial_size"]), indices=inds)
        bb.load_state_dict({k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}, strict=True)
        head = D.DINOHead(cfg["embed_dim"], cfg["K"], nlayers=3, hidden_dim=cfg["head_hidden"],
                          bottleneck_dim=cfg["head_bottleneck"])
        head.load_state_dict({k[len("dino_head."):]: v for k, v in sd.items() if k.startswith("dino_head.")})
        return bb, head

    (sb, sh), (tb, th) = make(student), make(teacher)
    model = SSLMetaArch(sb, tb, sh, th, cfg["K"], n_global_crops=cfg["n_global"], n_local_crops=cfg["n_local"],
                        dino_loss_weight=cfg["dino_w"], koleo_loss_weight=cfg["koleo_w"],
                        ibot_loss_weight=cfg["ibot_w"], fused_objective=fused).to(DEV)
    assert sorted(n for n, p in model.student.named_parameters() if p.requires_grad) == sorted(trainable)
    helpers.run_ssl_meta_steps(model, cfg, trainable, batch, arr)


def test_rejects_what_it_cannot_run():
    D, ops = _dinov2()
    with pytest.raises(RuntimeError):
        ops.softmax_center(torch.randn(4, 64, device=DEV, dtype=torch.float16), torch.zeros(1, 64, device=DEV), 0.05)
    with pytest.raises(RuntimeError):
        ops.softmax_center(torch.randn(4, 66, device=DEV), torch.zeros(1, 66, device=DEV), 0.05)   # K % 4
    with pytest.raises(NotImplementedError):
        D.DINOHead(64, 256, use_bn=True)
    with pytest.raises(RuntimeError, match="multiples of 64"):
        D.DINOHead(48, 256, hidden_dim=128, bottleneck_dim=64)(torch.zeros(2, 48, device=DEV))
