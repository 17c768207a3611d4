"""Host logic of apla_b200/dinov2/loss.py on the CPU: the kernels of csrc/ssl.cu are replaced -- IN THIS TEST ONLY, by
monkeypatching -- with torch restatements of what each C-ABI call computes (the closed forms of
tests/test_ssl_closed_forms.py), and the GPU parity tests of tests/test_ssl_gpu.py are then run unchanged on the CPU.
This pins everything around the kernels (crop-pair batching, the `chunk` view detection, row weights, the two-step centre
protocol, autograd wiring, error paths) and keeps the GPU test file itself exercised while no GPU is at hand.  The
product has no such fallback: without the patch every call below raises (test_no_fallback)."""
import importlib.util
import os

import pytest
import torch


def rel(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


HERE = os.path.dirname(os.path.abspath(__file__))


class FakeOps:
    """What the kernels compute, in torch, with the wrappers' signatures (apla_b200/dinov2/ops.py)."""
    F32, BF16 = torch.float32, torch.bfloat16

    @staticmethod
    def _rows(t, name, dtype=torch.float32):
        if t.dtype != dtype or t.dim() != 2 or t.stride(1) != 1:
            raise RuntimeError(f"{name}: bad dtype / layout")
        return t

    @classmethod
    def softmax_center(cls, t, center, temp, out=None):
        cls._rows(t, "teacher_output")
        if t.shape[1] % 4:
            raise RuntimeError("K must be a multiple of 4")
        return torch.softmax((t - center.reshape(1, -1)) * (1.0 / temp), dim=-1)

    @classmethod
    def colsum(cls, a, scale=1.0, splits=None):
        return cls._rows(a, "a").sum(0, keepdim=True) * scale

    @staticmethod
    def center_ema_(center, batch_sum, count, momentum):
        center.copy_(center * momentum + batch_sum.reshape(center.shape) * (1.0 / count) * (1 - momentum))
        return center

    @classmethod
    def sinkhorn_knopp(cls, t, temp, n_iterations, n_samples_world, all_reduce=None):
        P = torch.exp(cls._rows(t, "teacher_output") * (1.0 / temp))
        K = P.shape[1]
        for it in range(n_iterations):
            cs = cls.colsum(P)
            if all_reduce is not None:
                all_reduce(cs)
            P = P * ((1.0 / K) / cs)
            P = P * ((1.0 if it + 1 == n_iterations else 1.0 / n_samples_world) / P.sum(-1, keepdim=True))
        return P

    @staticmethod
    def _q(s, t0, t1, t_rows):
        tr = torch.arange(s.shape[0]) % t_rows
        return t0[tr] + (t1[tr] if t1 is not None else 0)

    @classmethod
    def soft_ce_fwd(cls, s, t0, t1, t_rows, w_row, w_uniform, inv_temp):
        cls._rows(s, "s"); cls._rows(t0, "t0")
        q = cls._q(s, t0, t1, t_rows)
        z = s * inv_temp
        lse, mass = torch.logsumexp(z, -1), q.sum(-1)
        w = w_uniform * (w_row[:s.shape[0]] if w_row is not None else 1.0)
        return (-w * ((q * z).sum(-1) - mass * lse)).sum(), lse, mass

    @classmethod
    def soft_ce_bwd(cls, s, t0, t1, t_rows, w_row, w_uniform, inv_temp, lse, mass, gscale, out_dtype=torch.float32):
        q = cls._q(s, t0, t1, t_rows)
        w = w_uniform * (w_row[:s.shape[0]] if w_row is not None else torch.ones(s.shape[0]))
        c = -w * inv_temp * (gscale.reshape(()) if gscale is not None else 1.0)
        return (c[:, None] * (q - mass[:, None] * torch.exp(s * inv_temp - lse[:, None]))).to(out_dtype)

    @staticmethod
    def l2norm_fwd(x, eps, out_dtype=torch.float32):
        return torch.nn.functional.normalize(x.float(), dim=-1, eps=eps).to(out_dtype)

    @staticmethod
    def l2norm_bwd(x, dy, eps):
        xf, g = x.float(), dy.float()
        inv = 1 / xf.norm(dim=-1, keepdim=True).clamp(min=eps)
        return (g * inv - xf * (xf * g).sum(-1, keepdim=True) * inv ** 3).to(dy.dtype)

    @staticmethod
    def weightnorm_fwd(g, v, out_dtype=torch.bfloat16):
        return (v * (g.reshape(-1, 1) / v.norm(dim=1, keepdim=True))).to(out_dtype)

    @staticmethod
    def weightnorm_bwd(g, v, dW, need_dg=True, need_dv=True):
        inv = 1 / v.norm(dim=1, keepdim=True)
        vd = (v * dW).sum(1, keepdim=True)
        dg = (vd * inv).reshape(g.shape) if need_dg else None
        dv = g.reshape(-1, 1) * inv * (dW - v * vd * inv * inv) if need_dv else None
        return dg, dv

    @classmethod
    def koleo_fwd(cls, x, eps, groups=1, weight=1.0):
        n = x.shape[0] // groups
        xn = cls.l2norm_fwd(x, eps)
        nn_idx, dist = [], []
        for g in range(groups):
            b = xn[g * n:(g + 1) * n]
            dots = (b @ b.t()).masked_fill(torch.eye(n, dtype=torch.bool), -1.0)
            i = dots.argmax(1)
            nn_idx.append(i.to(torch.int32)); dist.append((b - b[i] + 1e-8).norm(dim=-1))
        nn_idx, dist = torch.cat(nn_idx), torch.cat(dist)
        return (-torch.log(dist + eps) * weight / n).sum(), xn, nn_idx, dist

    @staticmethod
    def koleo_bwd(x, xn, nn_idx, dist, eps, gscale, groups=1, weight=1.0):
        n = x.shape[0] // groups
        up = -(weight / n) * (gscale.reshape(()) if gscale is not None else 1.0)
        c = up / ((dist + eps) * dist)
        gi = torch.zeros_like(xn)
        for g in range(groups):
            o = g * n
            for i in range(n):
                j = o + int(nn_idx[o + i])
                u = xn[o + i] - xn[j] + 1e-8
                gi[o + i] += c[o + i] * u
                gi[j] -= c[o + i] * u
        inv = 1 / x.norm(dim=-1, keepdim=True).clamp(min=eps)
        return gi * inv - x * (x * gi).sum(-1, keepdim=True) * inv ** 3

    @staticmethod
    def ema_update_(teacher, student, m):
        teacher.mul_(m).add_(student, alpha=1 - m)
        return teacher

    @classmethod
    def ssl_objective(cls, s, t, dino_center, ibot_center, masks_weight, B, n_local, teacher_temp, student_temp=0.1,
                      dino_weight=1.0, ibot_weight=1.0, gscale=None, ds_dtype=torch.bfloat16, need_grad=True):
        """The launch sequence of csrc/ssl.cu:ssl_objective, call for call."""
        K = s.shape[1]
        n_g, n_l = 2 * B, n_local * B
        n_m = t.shape[0] - n_g
        t_probs = torch.cat((cls.softmax_center(t[:n_g], dino_center, teacher_temp),
                             cls.softmax_center(t[n_g:], ibot_center, teacher_temp) if n_m else t[n_g:]))
        dsum = cls.colsum(t[:n_g])
        imean = (cls.colsum(t[n_g:], 1.0 / n_m) if n_m else torch.zeros(1, K)).view(1, 1, K)
        terms = 2 + max(2 * n_local, 1)
        spec = [(0, n_l, t_probs[:B], t_probs[B:n_g], B, None, 1.0 / (B * terms), dino_weight),
                (n_l, n_g, t_probs[:n_g], None, n_g, None, 2.0 / (n_g * terms), dino_weight),
                (n_l + n_g, n_m, t_probs[n_g:], None, max(n_m, 1), masks_weight, 1.0 / n_g, ibot_weight)]
        losses, ds = [], []
        for r0, n, t0, t1, t_rows, w, scale, weight in spec:
            sp = s[r0:r0 + n]
            loss, lse, mass = cls.soft_ce_fwd(sp, t0, t1, t_rows, w, scale, 1.0 / student_temp) if n else \
                (torch.zeros(()), None, None)
            losses.append(loss)
            if need_grad:
                ds.append(cls.soft_ce_bwd(sp, t0, t1, t_rows, w, scale * weight, 1.0 / student_temp, lse, mass, gscale,
                                          ds_dtype) if n else torch.zeros(0, K, dtype=ds_dtype))
        return dict(losses=torch.stack(losses), ds=torch.cat(ds) if need_grad else None, t_probs=t_probs,
                    dino_batch_sum=dsum, ibot_batch_mean=imean)


class FakeGemm:
    """The GEMM wrappers of apla_b200/ops.py that DINOHead composes, restated in torch with the kernels' rounding points
    (bf16 operands and outputs, fp32 accumulation)."""

    @staticmethod
    def _chk(*ts):
        for t in ts:
            if t.dtype != torch.bfloat16 or t.dim() != 2:
                raise RuntimeError("bf16 2-D operands expected")

    @classmethod
    def gemm_bias(cls, a, w, bias=None, out=None):
        cls._chk(a, w)
        y = a.float() @ w.float().t()
        return (y + bias if bias is not None else y).bfloat16()

    @classmethod
    def gemm_bias_gelu(cls, a, w, bias=None, h=None, g=None):
        cls._chk(a, w)
        y = a.float() @ w.float().t()
        y = y + bias if bias is not None else y
        return y.bfloat16(), torch.nn.functional.gelu(y).bfloat16()

    @classmethod
    def gemm_bias_ls_residual(cls, a, w, bias, gamma, resid, out=None):
        cls._chk(a, w)
        y = a.float() @ w.float().t()
        if bias is not None:
            y = y + bias
        if gamma is not None:
            y = y * gamma
        out.copy_(resid + y)
        return out

    @staticmethod
    def ls_cast(x, gamma=None, out=None):
        assert x.dtype == torch.float32
        return (x if gamma is None else x * gamma).bfloat16()

    @classmethod
    def proj_wgrad(cls, dysub, x, dw1, r, rowmap=None):
        cls._chk(dysub, x)
        assert dw1.dtype == torch.float32 and dw1.shape == (r, x.shape[1]) and rowmap is None
        dw1 += dysub.float().t()[:r] @ x.float()
        return dw1

    @classmethod
    def gemm_dgrad(cls, dy, wt, out=None):
        cls._chk(dy, wt)
        assert wt.shape[1] == dy.shape[1]
        return (dy.float() @ wt.float().t()).bfloat16()

    @classmethod
    def gemm_dgrad_gelu_bwd(cls, dy, wt, h, out=None):
        cls._chk(dy, wt, h)
        hf = h.float().requires_grad_(True)
        with torch.enable_grad():
            torch.nn.functional.gelu(hf).sum().backward()
        return ((dy.float() @ wt.float().t()) * hf.grad).bfloat16()

    @staticmethod
    def colsum(dy, db, n, rowmap=None):
        db += dy.float().sum(0)[:n]
        return db


def _load_gpu_tests():
    spec = importlib.util.spec_from_file_location("_ssl_gpu_tests", os.path.join(HERE, "test_ssl_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


G = _load_gpu_tests()
CASES = []
for _name in sorted(dir(G)):
    _fn = getattr(G, _name)
    if not _name.startswith("test_") or not callable(_fn):
        continue
    if _name in ("test_ssl_step_against_reference_vectors",     # needs the real fused backbone; its CPU counterpart is
                 "test_update_teacher_refreshes_cached_weights"):   # test_meta_arch_two_steps_against_reference_vectors;
        continue                                                # the second checks caches that live in the CUDA modules
    _params = [m for m in getattr(_fn, "pytestmark", []) if m.name == "parametrize"]
    if not _params:
        CASES.append(pytest.param(_name, {}, id=_name))
        continue
    _names = [a.strip() for a in _params[0].args[0].split(",")]
    for _vals in _params[0].args[1]:
        _vals = _vals if isinstance(_vals, (tuple, list)) else (_vals,)
        _kw = dict(zip(_names, _vals))
        if _kw.get("K", 0) >= 65536 or _kw.get("n", 0) >= 1 << 20:
            _kw = {k: (4096 if k == "K" else v) for k, v in _kw.items()}   # keep the CPU run short
            if _kw.get("n", 0) >= 1 << 20:
                _kw["n"] = 1 << 14
        CASES.append(pytest.param(_name, _kw, id=f"{_name}-{'-'.join(str(v) for v in _kw.values())}"))


@pytest.mark.parametrize("name,kwargs", CASES)
def test_gpu_suite_on_cpu_with_emulated_kernels(name, kwargs, monkeypatch):
    import apla_b200.dinov2 as D
    from apla_b200.dinov2 import dino_head, loss
    monkeypatch.setattr(loss, "ops", FakeOps)
    monkeypatch.setattr(dino_head, "R", FakeOps)
    monkeypatch.setattr(dino_head, "G", FakeGemm)
    monkeypatch.setattr(G, "DEV", "cpu")
    monkeypatch.setattr(G, "_dinov2", lambda: (D, FakeOps))
    getattr(G, name)(**kwargs)


def test_no_fallback():
    """Unpatched, the wrappers refuse to run without the device: nothing routes through torch math."""
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    import apla_b200.dinov2 as D
    from apla_b200.dinov2 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.softmax_center(torch.zeros(2, 64), torch.zeros(1, 64), 0.05)
    with pytest.raises(RuntimeError):
        D.DINOLoss(64)([torch.zeros(2, 64, requires_grad=True)], [torch.zeros(2, 64)])
    with pytest.raises(RuntimeError):
        D.KoLeoLoss()(torch.randn(4, 8))
    with pytest.raises(RuntimeError):
        D.update_teacher([torch.zeros(4)], [torch.zeros(4)], 0.99)
    with pytest.raises(RuntimeError):
        D.DINOHead(64, 256, hidden_dim=128, bottleneck_dim=64)(torch.zeros(2, 64))


# ---------------------------------------------------------------------------------------------------------------------
# the whole step: SSLMetaArch (apla_b200/hostdino.py) + DINOHead + the loss classes against the vectors recorded from the
# reference's DINOv2 meta-architecture.  On the CPU the backbone is the oracle's (the APLA attention has no CPU path) and
# the kernels under the head / losses are the emulations above; tests/test_ssl_gpu.py runs the same check with the fused
# multi-crop backbone and the real kernels.
# ---------------------------------------------------------------------------------------------------------------------
class OracleDinoBackbone(torch.nn.Module):
    """TEST-ONLY backbone: parameters registered under the reference's names, arithmetic by oracle/ssl_oracle.py."""

    def __init__(self, sd, trainable, cfg):
        super().__init__()
        self.cfg = cfg
        for k, v in sd.items():                                 # nested containers -> the reference's dotted names
            *path, leaf = k.split(".")
            mod = self
            for part in path:
                if not hasattr(mod, part):
                    mod.add_module(part, torch.nn.Module())
                mod = getattr(mod, part)
            if v.is_floating_point():
                mod.register_parameter(leaf, torch.nn.Parameter(v.clone(), requires_grad=k in trainable))
            else:
                mod.register_buffer(leaf, v.clone())

    def forward(self, x, masks=None, is_training=True):
        from oracle import ssl_oracle as S
        sd = {"backbone." + k: v for k, v in list(self.named_parameters()) + list(self.named_buffers())}
        kw = dict(patch=self.cfg["patch"], depth=self.cfg["depth"], num_heads=self.cfg["num_heads"])
        if isinstance(x, (list, tuple)):
            return [dict(x_norm_clstoken=o["cls"], x_norm_patchtokens=o["patch"]) for o in S.dinov2_backbone(sd, x, masks, **kw)]
        o = S.dinov2_backbone(sd, x, masks, **kw)
        return dict(x_norm_clstoken=o["cls"], x_norm_patchtokens=o["patch"])


class ExactGemm(FakeGemm):
    """The same wrappers with NO rounding (used with dino_head.BF16 patched to float32): isolates the wiring."""

    @staticmethod
    def _chk(*ts):
        pass

    @staticmethod
    def gemm_bias(a, w, bias=None, out=None):
        y = a @ w.t()
        return y + bias if bias is not None else y

    @staticmethod
    def gemm_bias_gelu(a, w, bias=None, h=None, g=None):
        y = a @ w.t()
        y = y + bias if bias is not None else y
        return y, torch.nn.functional.gelu(y)

    @staticmethod
    def ls_cast(x, gamma=None, out=None):
        return x if gamma is None else x * gamma

    @staticmethod
    def gemm_dgrad(dy, wt, out=None):
        return dy @ wt.t()

    @staticmethod
    def gemm_dgrad_gelu_bwd(dy, wt, h, out=None):
        hf = h.clone().requires_grad_(True)
        with torch.enable_grad():
            torch.nn.functional.gelu(hf).sum().backward()
        return (dy @ wt.t()) * hf.grad


class ExactOps(FakeOps):
    @staticmethod
    def l2norm_fwd(x, eps, out_dtype=torch.float32):
        return torch.nn.functional.normalize(x.float(), dim=-1, eps=eps)

    @staticmethod
    def weightnorm_fwd(g, v, out_dtype=torch.float32):
        return v * (g.reshape(-1, 1) / v.norm(dim=1, keepdim=True))


@pytest.mark.parametrize("fused", [False, True], ids=["per-term-losses", "fused-head-objective"])
@pytest.mark.parametrize("exact", [True, False], ids=["exact-arithmetic", "bf16-rounding-points"])
def test_meta_arch_two_steps_against_reference_vectors(monkeypatch, exact, fused):
    import sys
    sys.path.insert(0, HERE)
    import helpers
    import apla_b200.dinov2 as D
    from apla_b200.dinov2 import dino_head, loss
    from apla_b200.hostdino import SSLMetaArch
    monkeypatch.setattr(loss, "ops", FakeOps)
    monkeypatch.setattr(dino_head, "R", ExactOps if exact else FakeOps)
    monkeypatch.setattr(dino_head, "G", ExactGemm if exact else FakeGemm)
    if exact:
        monkeypatch.setattr(dino_head, "BF16", torch.float32)
    cfg, student, teacher, trainable, batch, arr = helpers.ssl_step_case()

    def split(sd):
        bb = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
        hd = {k[len("dino_head."):]: v for k, v in sd.items() if k.startswith("dino_head.")}
        return bb, hd

    def head_of(hd):
        h = D.DINOHead(cfg["embed_dim"], cfg["K"], nlayers=3, hidden_dim=cfg["head_hidden"],
                       bottleneck_dim=cfg["head_bottleneck"])
        h.load_state_dict(hd)
        return h

    (sbb, shd), (tbb, thd) = split(student), split(teacher)
    bb_train = {n[len("backbone."):] for n in trainable if n.startswith("backbone.")}
    model = SSLMetaArch(OracleDinoBackbone(sbb, bb_train, cfg), OracleDinoBackbone(tbb, set(), cfg), head_of(shd),
                        head_of(thd), cfg["K"], n_global_crops=cfg["n_global"], n_local_crops=cfg["n_local"],
                        dino_loss_weight=cfg["dino_w"], koleo_loss_weight=cfg["koleo_w"], ibot_loss_weight=cfg["ibot_w"],
                        fused_objective=fused)
    assert sorted(n for n, p in model.student.named_parameters() if p.requires_grad) == sorted(trainable)
    assert not any(p.requires_grad for p in model.teacher.parameters())
    ema = lambda s, t, m: [FakeOps.ema_update_(b.data, a.data, m) for a, b in zip(s, t)]      # noqa: E731
    bars = ((1e-5, 0.999999, 1e-5),) * 2 if exact else helpers.SSL_BF16_BARS
    helpers.run_ssl_meta_steps(model, cfg, trainable, batch, arr, ema_fn=ema, bars=bars)


def test_c_abi_argument_checks_without_a_gpu():
    """Every self-supervised entry point refuses malformed sizes BEFORE touching the device: exercises the ctypes
    marshalling of each signature (argument count, order and types as parsed from include/apla_b200.h) on the CPU."""
    from apla_b200._lib import LIB
    dll = LIB.load()
    bad = {
        "apla_softmax_center": (None, 0, None, 1.0, 1, 6, None, 0, None),                       # K % 4
        "apla_colsum_f32": (None, 0, 1, 0, None, 1, 1.0, None, None),                           # K = 0
        "apla_center_ema": (None, None, 0, 1.0, 0.9, None),                                     # K = 0
        "apla_soft_ce_fwd": (None, 0, 1, 6, None, None, 0, 1, None, 1.0, 10.0, None, None, None, None),
        "apla_soft_ce_bwd": (None, 0, 1, 8, None, None, 0, 0, None, 1.0, 10.0, None, None, None, None, 0, 0, None),
        "apla_soft_ce_fwd_bwd": (None, 0, 1, 8, None, None, 0, 1, None, 1.0, 1.0, 10.0, None, None, None, 0, 0, None),  # no ds
        "apla_sk_exp": (None, 0, 1.0, 1, 6, None, 0, None),
        "apla_sk_normalize": (None, 0, 1, 6, None, 1.0, 1.0, None),
        "apla_sum_f32": (None, -1, 1.0, None, None),
        "apla_l2norm_fwd": (None, 0, 1, 1, 0, 1e-12, None, None, 0, None),                      # d = 0
        "apla_l2norm_bwd": (None, 0, 1, None, 0, 1, 1, 0, 1e-12, None, 0, None),
        "apla_weightnorm_fwd": (None, None, 0, 4, None, None, None),
        "apla_weightnorm_bwd": (None, None, None, 0, 0, 4, None, None, None),
        "apla_koleo_fwd": (None, 1, 0, 8, 1e-8, 1.0, None, None, None, None),                   # n < 1
        "apla_koleo_bwd": (None, None, 1, 0, 8, 1e-8, 1e-8, 1.0, None, None, None, None, None),
        "apla_ema_update": (None, None, -1, 0.99, None),
        "apla_ssl_objective": (None, 0, None, 0, None, 0, None, None, None, 0, 8, 0, 64, 0.05, 0.1, 1.0, 1.0, None, None, 1,
                               None, 0, 1, None, None, None, None, None),                         # B = 0
    }
    for name, args in bad.items():
        assert len(args) == len(LIB.protos[name][1]), name
        rc = getattr(dll, name)(*args)
        assert rc == 1, (name, rc)
        assert dll.apla_last_error(), name
    # a wide row buffer is refused rather than launched with too much shared memory
    assert dll.apla_koleo_fwd(None, 1, 4, 20000, 1e-8, 1.0, None, None, None, None) == 1
    assert b"too wide" in dll.apla_last_error()


def test_meta_arch_sinkhorn_knopp_centering(monkeypatch):
    """`centering: sinkhorn_knopp` (models.py:303-315): the step's loss with the teacher targets swapped for the
    Sinkhorn-Knopp ones, against the same assembly done by hand over the oracle (exact arithmetic under the wiring)."""
    import sys
    sys.path.insert(0, HERE)
    import helpers
    import apla_b200.dinov2 as D
    from apla_b200.dinov2 import dino_head, loss
    from apla_b200.hostdino import SSLMetaArch
    from oracle import ssl_oracle as S
    monkeypatch.setattr(loss, "ops", FakeOps)
    monkeypatch.setattr(dino_head, "R", ExactOps)
    monkeypatch.setattr(dino_head, "G", ExactGemm)
    monkeypatch.setattr(dino_head, "BF16", torch.float32)
    cfg, student, teacher, trainable, batch, arr = helpers.ssl_step_case()
    split = lambda sd, pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}          # noqa: E731

    def head_of(sd):
        h = D.DINOHead(cfg["embed_dim"], cfg["K"], nlayers=3, hidden_dim=cfg["head_hidden"],
                       bottleneck_dim=cfg["head_bottleneck"])
        h.load_state_dict(split(sd, "dino_head."))
        return h

    model = SSLMetaArch(OracleDinoBackbone(split(student, "backbone."), set(), cfg),
                        OracleDinoBackbone(split(teacher, "backbone."), set(), cfg), head_of(student), head_of(teacher),
                        cfg["K"], n_local_crops=cfg["n_local"], koleo_loss_weight=cfg["koleo_w"],
                        centering="sinkhorn_knopp")
    loss_val, parts = model(batch, teacher_temp=cfg["teacher_temp"])
    # by hand over the oracle
    kw = dict(patch=cfg["patch"], depth=cfg["depth"], num_heads=cfg["num_heads"])
    masks, B, nl = batch["collated_masks"], cfg["B"], cfg["n_local"]
    idx, mw = S.mask_indices_of(masks), S.masks_weight_of(masks)
    with torch.no_grad():
        t = S.dinov2_backbone(teacher, batch["collated_global_crops"], None, **kw)
        a, b = t["cls"].chunk(2)
        t_out = S.dino_head_forward(split(teacher, "dino_head."), torch.cat((b, a, t["patch"].flatten(0, 1)[idx])))
        t_d = S.sinkhorn_knopp(t_out[:2 * B], cfg["teacher_temp"]).view(2, B, -1)
        t_i = S.sinkhorn_knopp(t_out[2 * B:], cfg["teacher_temp"], n_samples_world=idx.shape[0])
        sg, sl = S.dinov2_backbone(student, [batch["collated_global_crops"], batch["collated_local_crops"]],
                                   [masks, None], **kw)
        s_out = S.dino_head_forward(split(student, "dino_head."),
                                    torch.cat((sl["cls"], sg["cls"], sg["patch"].flatten(0, 1)[idx])))
        s_l, s_g, s_p = s_out[:nl * B], s_out[nl * B:nl * B + 2 * B], s_out[nl * B + 2 * B:]
        terms = 2 + 2 * nl
        want = (S.dino_loss(s_l.chunk(nl), list(t_d)) / terms + S.dino_loss([s_g], [t_d.flatten(0, 1)]) * 2 / terms
                + cfg["koleo_w"] * sum(S.koleo_loss(p) for p in sg["cls"].chunk(2))
                + S.ibot_loss_masked(s_p, t_i, masks, n_masked_patches=idx.shape[0], masks_weight=mw))
    assert rel(loss_val.detach(), want) < 1e-5
    assert model.dino_loss.updated is True                       # no centre update is pending in this mode
    with pytest.raises(ValueError):
        SSLMetaArch(torch.nn.Identity(), torch.nn.Identity(), torch.nn.Identity(), torch.nn.Identity(), 8,
                    centering="sinkhorn_knopp", fused_objective=True)


def test_cancel_last_layer_grads_drops_only_the_student_prototype_layer():
    """possibly_cancel_last_layer_grads (dinov2/trainer.py:86-91): while the last layer is frozen its gradients are set to
    None after backward -- the student head's `last_layer.weight_g / weight_v` and nothing else."""
    import apla_b200.dinov2 as D
    from apla_b200.hostdino import SSLMetaArch
    head = lambda: D.DINOHead(16, 32, nlayers=2, hidden_dim=16, bottleneck_dim=8)          # noqa: E731
    model = SSLMetaArch(torch.nn.Linear(4, 4), torch.nn.Linear(4, 4), head(), head(), 32)
    for p in model.parameters():
        p.grad = torch.ones_like(p)
    dropped = model.cancel_last_layer_grads()
    names = {n for n, p in model.named_parameters() if p.grad is None}
    assert dropped == 2 and names == {"student.dino_head.last_layer.weight_g", "student.dino_head.last_layer.weight_v"}
    assert model.cancel_last_layer_grads() == 0


def test_update_teacher_skips_frozen_identical_pairs(monkeypatch):
    """Frozen student tensors that the teacher holds bit-identically are left alone (no launch, no version bump -> the
    teacher's cached weight copies / position table stay valid); trainable pairs and frozen-but-different pairs are averaged."""
    from apla_b200.dinov2 import loss
    calls = []

    class CountingOps:
        @staticmethod
        def ema_update_(t, s, m):
            calls.append(t.data_ptr())
            t.mul_(m).add_(s, alpha=1 - m)
            return t
    monkeypatch.setattr(loss, "ops", CountingOps)
    s_frozen, s_train, s_frozen2 = torch.randn(8), torch.randn(8, requires_grad=True), torch.randn(8)
    t_same, t_train, t_diff = s_frozen.clone(), s_train.detach().clone() + 1.0, s_frozen2.clone() + 1.0
    v_same = t_same._version
    for _ in range(3):
        loss.update_teacher([s_frozen, s_train, s_frozen2], [t_same, t_train, t_diff], 0.5)
    assert calls.count(t_same.data_ptr()) == 0 and t_same._version == v_same and torch.equal(t_same, s_frozen)
    assert calls.count(t_train.data_ptr()) == 3 and calls.count(t_diff.data_ptr()) == 3
    assert torch.allclose(t_diff, s_frozen2 + 0.125) and torch.allclose(t_train, s_train.detach() + 0.125)
    with torch.no_grad():
        s_frozen.add_(1.0)                               # the frozen student tensor changed after all: averaged again
    loss.update_teacher([s_frozen], [t_same], 0.5)
    assert calls.count(t_same.data_ptr()) == 1 and torch.allclose(t_same, s_frozen - 0.5)
