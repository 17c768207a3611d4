"""The kernels of apla_b200/csrc/ssl.cu executed ON THE CPU by a small SIMT emulator (tests/emu/cuda_emu.h: one OS thread
per CUDA thread, real barriers, warp shuffles through a guarded buffer) and held to the same parity tests as on the GPU
(tests/test_ssl_gpu.py, run here unchanged with device "cpu" and smaller K).  The kernel source is compiled as it is --
tests/emu/build_emu.py rewrites only the launch syntax -- so this pins indexing, reductions, barrier placement and the fp32
arithmetic of every kernel before its first hardware run; it says nothing about speed, coalescing or fast-math rounding.
TEST INFRASTRUCTURE: nothing in apla_b200/ can reach the emulator."""
import ctypes
import importlib.util
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import build_emu  # noqa: E402

from apla_b200._lib import parse_header  # noqa: E402

F32, BF16, I32 = torch.float32, torch.bfloat16, torch.int32


@pytest.fixture(scope="module")
def emu():
    dll = ctypes.CDLL(build_emu.build())
    protos = parse_header()
    for name in build_emu.ENTRY:
        fn = getattr(dll, name)
        fn.restype, fn.argtypes = protos[name]
    dll.emu_last_error.restype = ctypes.c_char_p
    return dll


def _p(t):
    return None if t is None else t.data_ptr()


GUARD, SENTINEL = 64, 12345.678


class _Guarded:
    """Output buffers with sentinel-filled guard bands on both sides: an out-of-bounds WRITE of a kernel is caught after
    the call (the emulator itself would let it land in the heap silently)."""
    live = []

    @classmethod
    def empty(cls, *shape, dtype=F32):
        n = 1
        for d in shape:
            n *= d
        flat = torch.full((n + 2 * GUARD,), SENTINEL).to(dtype)
        cls.live.append(flat)
        return flat[GUARD:GUARD + n].view(*shape) if shape else flat[GUARD:GUARD + 1].view(())

    @classmethod
    def check(cls, where):
        for flat in cls.live:
            want = torch.tensor(SENTINEL).to(flat.dtype)
            assert bool((flat[:GUARD] == want).all()) and bool((flat[-GUARD:] == want).all()), \
                f"{where} wrote outside its output buffer"
        cls.live.clear()


class EmuOps:
    """apla_b200/dinov2/ops.py's wrappers over the emulated library, for CPU tensors."""
    dll = None

    @classmethod
    def call(cls, name, *args):
        rc = getattr(cls.dll, name)(*args, None)
        if rc != 0:
            raise RuntimeError(f"{name} failed (rc={rc}): {cls.dll.emu_last_error().decode()}")
        _Guarded.check(name)

    @staticmethod
    def _rows(t, name, dtype=F32):
        if t.dtype != dtype or t.dim() != 2 or t.stride(1) != 1:
            raise RuntimeError(f"{name}: bad dtype / layout")
        return t

    @classmethod
    def softmax_center(cls, t, center, temp, out=None):
        cls._rows(t, "teacher_output")
        n, K = t.shape
        out = _Guarded.empty(n, K)
        cls.call("apla_softmax_center", _p(t), t.stride(0), _p(center), 1.0 / temp, n, K, _p(out), out.stride(0))
        return out

    @classmethod
    def colsum(cls, a, scale=1.0, splits=None):
        cls._rows(a, "a")
        n, K = a.shape
        splits = splits or max(1, min(32, n // 64)) if n >= 64 else 3          # exercise ragged splits on small inputs
        ws, out = _Guarded.empty(splits, K), _Guarded.empty(1, K)
        cls.call("apla_colsum_f32", _p(a), a.stride(0), n, K, _p(ws), splits, float(scale), _p(out))
        return out

    @classmethod
    def center_ema_(cls, center, batch_sum, count, momentum):
        cls.call("apla_center_ema", _p(center), _p(batch_sum), center.numel(), 1.0 / count, momentum)
        return center

    @classmethod
    def sinkhorn_knopp(cls, t, temp, n_iterations, n_samples_world, all_reduce=None):
        cls._rows(t, "teacher_output")
        n, K = t.shape
        P = _Guarded.empty(n, K)
        keep = list(_Guarded.live)
        cls.call("apla_sk_exp", _p(t), t.stride(0), 1.0 / temp, n, K, _p(P), P.stride(0))
        for it in range(n_iterations):
            cs = cls.colsum(P)
            _Guarded.live.extend(keep)                       # P's guard bands are re-checked after the in-place pass
            cls.call("apla_sk_normalize", _p(P), P.stride(0), n, K, _p(cs), 1.0 / K,
                     1.0 if it + 1 == n_iterations else 1.0 / n_samples_world)
        return P

    @classmethod
    def soft_ce_fwd(cls, s, t0, t1, t_rows, w_row, w_uniform, inv_temp):
        cls._rows(s, "s"); cls._rows(t0, "t0")
        rows, K = s.shape
        row_loss, lse, mass = (_Guarded.empty(max(rows, 1)) for _ in range(3))
        loss = _Guarded.empty()
        cls.call("apla_soft_ce_fwd", _p(s), s.stride(0), rows, K, _p(t0), _p(t1), t0.stride(0), int(t_rows), _p(w_row),
                 float(w_uniform), float(inv_temp), _p(row_loss), _p(lse), _p(mass))
        cls.call("apla_sum_f32", _p(row_loss), rows, 1.0, _p(loss))
        return loss, lse, mass

    @classmethod
    def soft_ce_bwd(cls, s, t0, t1, t_rows, w_row, w_uniform, inv_temp, lse, mass, gscale, out_dtype=F32):
        rows, K = s.shape
        ds = _Guarded.empty(rows, K, dtype=out_dtype)
        cls.call("apla_soft_ce_bwd", _p(s), s.stride(0), rows, K, _p(t0), _p(t1), t0.stride(0), int(t_rows), _p(w_row),
                 float(w_uniform), float(inv_temp), _p(lse), _p(mass), _p(gscale), _p(ds), ds.stride(0),
                 int(out_dtype == BF16))
        return ds

    @classmethod
    def l2norm_fwd(cls, x, eps, out_dtype=F32):
        n, d = x.shape
        y = _Guarded.empty(n, d, dtype=out_dtype)
        cls.call("apla_l2norm_fwd", _p(x), x.stride(0), int(x.dtype == F32), n, d, float(eps),
                 _p(y) if out_dtype == BF16 else None, _p(y) if out_dtype == F32 else None, y.stride(0))
        return y

    @classmethod
    def l2norm_bwd(cls, x, dy, eps):
        n, d = x.shape
        dx = _Guarded.empty(n, d, dtype=dy.dtype)
        cls.call("apla_l2norm_bwd", _p(x), x.stride(0), int(x.dtype == F32), _p(dy), dy.stride(0), int(dy.dtype == F32),
                 n, d, float(eps), _p(dx), dx.stride(0))
        return dx

    @classmethod
    def weightnorm_fwd(cls, g, v, out_dtype=BF16):
        K, d = v.shape
        w = _Guarded.empty(K, d, dtype=out_dtype)
        cls.call("apla_weightnorm_fwd", _p(g), _p(v), K, d, _p(w) if out_dtype == BF16 else None,
                 _p(w) if out_dtype == F32 else None)
        return w

    @classmethod
    def weightnorm_bwd(cls, g, v, dW, need_dg=True, need_dv=True):
        K, d = v.shape
        dg = _Guarded.empty(*g.shape) if need_dg else None
        dv = _Guarded.empty(*v.shape) if need_dv else None
        cls.call("apla_weightnorm_bwd", _p(g), _p(v), _p(dW), dW.stride(0), K, d, _p(dg), _p(dv))
        return dg, dv

    @classmethod
    def koleo_fwd(cls, x, eps, groups=1, weight=1.0):
        n, D = x.shape[0] // groups, x.shape[1]
        xn = cls.l2norm_fwd(x, eps, F32)
        nn = _Guarded.empty(groups * n, dtype=I32); dist = _Guarded.empty(groups * n)
        row_loss, loss = _Guarded.empty(groups * n), _Guarded.empty()
        cls.call("apla_koleo_fwd", _p(xn), groups, n, D, float(eps), float(weight), _p(nn), _p(dist), _p(row_loss))
        cls.call("apla_sum_f32", _p(row_loss), groups * n, 1.0, _p(loss))
        return loss, xn, nn, dist

    @classmethod
    def koleo_bwd(cls, x, xn, nn, dist, eps, gscale, groups=1, weight=1.0):
        n, D = x.shape[0] // groups, x.shape[1]
        dx = _Guarded.empty(*x.shape)
        cls.call("apla_koleo_bwd", _p(x), _p(xn), groups, n, D, float(eps), float(eps), float(weight), _p(nn), _p(dist),
                 _p(gscale), _p(dx))
        return dx

    @classmethod
    def ema_update_(cls, teacher, student, m):
        cls.call("apla_ema_update", _p(teacher), _p(student), teacher.numel(), float(m))
        return teacher

    @classmethod
    def ssl_objective(cls, s, t, dino_center, ibot_center, masks_weight, B, n_local, teacher_temp, student_temp=0.1,
                      dino_weight=1.0, ibot_weight=1.0, gscale=None, ds_dtype=BF16, need_grad=True):
        rows, K = s.shape
        n_masked = t.shape[0] - 2 * B
        splits = 3
        t_probs, row_ws, col_ws = _Guarded.empty(*t.shape), _Guarded.empty(3 * rows), _Guarded.empty(splits, K)
        losses, dsum, imean = _Guarded.empty(3), _Guarded.empty(1, K), _Guarded.empty(1, 1, K)
        ds = _Guarded.empty(rows, K, dtype=ds_dtype) if need_grad else None
        live = list(_Guarded.live)                      # the inner launchers are not wrapped: check once at the end
        cls.call("apla_ssl_objective", _p(s), s.stride(0), _p(t), t.stride(0), _p(t_probs), t_probs.stride(0),
                 _p(dino_center), _p(ibot_center), _p(masks_weight), B, n_local, n_masked, K, float(teacher_temp),
                 float(student_temp), float(dino_weight), float(ibot_weight), _p(row_ws), _p(col_ws), splits, _p(ds),
                 K if ds is None else ds.stride(0), int(ds_dtype == BF16), _p(gscale), _p(losses), _p(dsum), _p(imean))
        del live
        return dict(losses=losses, ds=ds, t_probs=t_probs, dino_batch_sum=dsum, ibot_batch_mean=imean)


def _load(name):
    spec = importlib.util.spec_from_file_location("_" + name, os.path.join(HERE, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


G = _load("test_ssl_gpu")
HOST = _load("test_ssl_host")
SHRINK = dict(K=512, n=300, P=32, Dm=96)           # upper bounds for the emulated run (one OS thread per CUDA thread)
SKIP = {"test_ssl_step_against_reference_vectors",   # needs the tensor-core backbone
        "test_update_teacher_refreshes_cached_weights",   # ditto (the caches it checks live in the CUDA modules)
        "test_rejects_what_it_cannot_run"}           # dtype / K % 4 refusals live in the real wrappers (test_ssl_host)
CASES = []
for _name, _kw in ((c.values[0], c.values[1]) for c in HOST.CASES):
    if _name in SKIP:
        continue
    _kw = {k: (min(v, SHRINK[k]) if k in SHRINK and isinstance(v, int) and not (k == "K" and v < 4096) else v)
           for k, v in _kw.items()}                      # K below 4096 is kept as is (2052: more float4s than threads)
    if _name == "test_dino_head":
        _kw = dict(_kw, n=min(_kw["n"], 40), in_dim=min(_kw["in_dim"], 128), hidden=min(_kw["hidden"], 128),
                   bott=min(_kw["bott"], 64), K=min(_kw["K"], 256))
    _id = f"{_name}-{'-'.join(str(v) for v in _kw.values())}"
    if _id not in [c.id for c in CASES]:
        CASES.append(pytest.param(_name, _kw, id=_id))


@pytest.mark.parametrize("name,kwargs", CASES)
def test_gpu_suite_on_the_emulated_kernels(name, kwargs, emu, monkeypatch):
    import apla_b200.dinov2 as D
    from apla_b200.dinov2 import dino_head, loss
    EmuOps.dll = emu
    monkeypatch.setattr(loss, "ops", EmuOps)
    monkeypatch.setattr(dino_head, "R", EmuOps)
    monkeypatch.setattr(dino_head, "G", HOST.FakeGemm)          # the GEMMs are not in ssl.cu
    monkeypatch.setattr(G, "DEV", "cpu")
    monkeypatch.setattr(G, "_dinov2", lambda: (D, EmuOps))
    getattr(G, name)(**kwargs)


@pytest.mark.parametrize("fused", [False, True], ids=["per-term-losses", "fused-head-objective"])
def test_meta_arch_step_on_the_emulated_kernels(emu, monkeypatch, fused):
    """Two whole steps against the reference's recorded vectors with every ssl.cu kernel emulated (oracle backbone,
    GEMMs restated at their bf16 rounding points)."""
    EmuOps.dll = emu
    import helpers
    import apla_b200.dinov2 as D
    from apla_b200.dinov2 import dino_head, loss
    from apla_b200.hostdino import SSLMetaArch
    monkeypatch.setattr(loss, "ops", EmuOps)
    monkeypatch.setattr(dino_head, "R", EmuOps)
    monkeypatch.setattr(dino_head, "G", HOST.FakeGemm)
    cfg, student, teacher, trainable, batch, arr = helpers.ssl_step_case()
    split = lambda sd, pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}          # noqa: E731

    def head_of(sd):
        h = D.DINOHead(cfg["embed_dim"], cfg["K"], nlayers=3, hidden_dim=cfg["head_hidden"],
                       bottleneck_dim=cfg["head_bottleneck"])
        h.load_state_dict(split(sd, "dino_head."))
        return h

    bb_train = {n[len("backbone."):] for n in trainable if n.startswith("backbone.")}
    model = SSLMetaArch(HOST.OracleDinoBackbone(split(student, "backbone."), bb_train, cfg),
                        HOST.OracleDinoBackbone(split(teacher, "backbone."), set(), cfg), head_of(student),
                        head_of(teacher), cfg["K"], n_global_crops=cfg["n_global"], n_local_crops=cfg["n_local"],
                        dino_loss_weight=cfg["dino_w"], koleo_loss_weight=cfg["koleo_w"], ibot_loss_weight=cfg["ibot_w"],
                        fused_objective=fused)
    ema = lambda s, t, m: [EmuOps.ema_update_(b.data, a.data, m) for a, b in zip(s, t)]            # noqa: E731
    helpers.run_ssl_meta_steps(model, cfg, trainable, batch, arr, ema_fn=ema)
