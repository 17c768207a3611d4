"""Per-kernel timing of the self-supervised row kernels (csrc/ssl.cu) at the C4 shapes of SURVEY.md App. A: 64 images
per GPU -> 128 global + 512 local CLS rows, K = 65 536 prototypes, ~20 % of 128 x 256 global patches masked.  CUDA
events, L2 flushed between iterations; one JSON line per kernel with the ALGORITHMIC bytes and the
achieved fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json).  Fills profiles/, never bench.py.

    python tools/bench_ssl_kernels.py [--rows-patch 6554] [--K 65536]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apla_b200.dinov2 import ops  # noqa: E402

dev = "cuda"


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6541.5, "fallback of B200_PROFILING.md"


def timeit(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--K", type=int, default=65536)
    ap.add_argument("--rows-patch", type=int, default=6554)
    ap.add_argument("--n-local", type=int, default=8)
    ap.add_argument("--D", type=int, default=1024)
    ap.add_argument("--bottleneck", type=int, default=256)
    a = ap.parse_args()
    B, K, NP, NL = a.B, a.K, a.rows_patch, a.n_local
    peak, src = peak_gbs()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = []

    def rec(name, secs, nbytes, launches=1):
        gbs = nbytes / secs / 1e9
        out.append(dict(kernel=name, us=round(secs * 1e6, 1), algorithmic_bytes=int(nbytes), gbs=round(gbs, 1),
                        frac_of_hbm_peak=round(gbs / peak, 3), launches=launches))
        print(json.dumps(out[-1]), flush=True)

    f4 = 4
    # teacher: 2B CLS rows + NP masked-patch rows
    for name, n in (("softmax_center cls", 2 * B), ("softmax_center patches", NP)):
        t = torch.randn(n, K, device=dev)
        c = torch.zeros(1, K, device=dev)
        o = torch.empty_like(t)
        rec(name, timeit(lambda: ops.softmax_center(t, c, 0.05, out=o), flush), n * K * f4 * 2 + K * f4)
        rec(name.replace("softmax_center", "colsum"), timeit(lambda: ops.colsum(t), flush), n * K * f4 + K * f4, 2)
        del t, o
    # student: DINO local (8B rows against two teacher crops), DINO global (2B rows), iBOT (NP rows, per-row weights)
    t2 = torch.softmax(torch.randn(2, B, K, device=dev), -1)
    for name, rows, t0, t1, t_rows, w in (
            ("soft_ce dino local", NL * B, t2[0], t2[1], B, None),
            ("soft_ce dino global", 2 * B, t2.flatten(0, 1), None, 2 * B, None),
            ("soft_ce ibot", NP, None, None, NP, "rows")):
        s = torch.randn(rows, K, device=dev)
        if t0 is None:
            t0 = torch.softmax(torch.randn(rows, K, device=dev), -1)
        wr = torch.rand(rows, device=dev) if w else None
        nt = 2 if t1 is not None else 1
        res = ops.soft_ce_fwd(s, t0, t1, t_rows, wr, 1.0 / rows, 10.0)
        rec(name + " fwd", timeit(lambda: ops.soft_ce_fwd(s, t0, t1, t_rows, wr, 1.0 / rows, 10.0), flush),
            rows * K * f4 * (1 + nt), 2)
        g = torch.ones((), device=dev)
        for dt, nb in ((torch.float32, 4), (torch.bfloat16, 2)):
            rec(f"{name} bwd ds={str(dt)[6:]}",
                timeit(lambda: ops.soft_ce_bwd(s, t0, t1, t_rows, wr, 1.0 / rows, 10.0, res[1], res[2], g, dt), flush),
                rows * K * (f4 * (1 + nt) + nb))
        del s, res
    # the whole objective as one native launch sequence (APLA_SSL_SPLIT_CE=1 in the environment = two-kernel CE form)
    n_s, n_t = NL * B + 2 * B + NP, 2 * B + NP
    s_all, t_all = torch.randn(n_s, K, device=dev), torch.randn(n_t, K, device=dev)
    dc, ic = torch.zeros(1, K, device=dev), torch.zeros(1, 1, K, device=dev)
    mw = torch.rand(NP, device=dev)
    form = "split CE" if os.environ.get("APLA_SSL_SPLIT_CE", "0") not in ("", "0") else "one-launch CE"
    # algorithmic bytes: teacher rows read + probs written (+ read once for the column sums), student rows read once,
    # teacher probs read once per consuming row group, bf16 ds written
    nbytes = n_t * K * f4 * 3 + n_s * K * f4 + (2 * B * 2 + NP) * K * f4 + n_s * K * 2
    rec(f"ssl_objective ({form}), bf16 ds", timeit(lambda: ops.ssl_objective(s_all, t_all, dc, ic, mw, B, NL, 0.05), flush, iters=5),
        nbytes, 12 if form == "one-launch CE" else 15)
    del s_all, t_all
    # head tail: L2 normalisation of the bottleneck rows, weight normalisation of the last layer
    rows = NL * B + 2 * B + NP
    z = torch.randn(rows, a.bottleneck, device=dev).bfloat16()
    rec("l2norm fwd bf16", timeit(lambda: ops.l2norm_fwd(z, 1e-12, torch.bfloat16), flush), rows * a.bottleneck * 4)
    rec("l2norm bwd bf16", timeit(lambda: ops.l2norm_bwd(z, z, 1e-12), flush), rows * a.bottleneck * 6)
    g_, v_ = torch.ones(K, 1, device=dev), torch.randn(K, a.bottleneck, device=dev)
    rec("weightnorm fwd", timeit(lambda: ops.weightnorm_fwd(g_, v_), flush), K * a.bottleneck * 6 + K * 4)
    rec("weightnorm bwd", timeit(lambda: ops.weightnorm_bwd(g_, v_, v_), flush), K * a.bottleneck * 12 + K * 8)
    # KoLeo on the B global CLS rows of one crop, teacher EMA over a ViT-L sized arena
    x = torch.randn(B, a.D, device=dev)
    rec("koleo fwd (l2norm + nn + sum)", timeit(lambda: ops.koleo_fwd(x, 1e-8), flush), B * a.D * 4 * 3, 3)
    st = ops.koleo_fwd(x, 1e-8)
    g = torch.ones((), device=dev)
    rec("koleo bwd", timeit(lambda: ops.koleo_bwd(x, st[1], st[2], st[3], 1e-8, g), flush), B * a.D * 4 * 3)
    n = 304_000_000
    tt, ss = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    rec("teacher EMA 304 M floats", timeit(lambda: ops.ema_update_(tt, ss, 0.994), flush, iters=5), n * 12)
    print(json.dumps(dict(peak_gbs=peak, peak_source=src, shapes=dict(B=B, K=K, rows_patch=NP, n_local=NL))))


if __name__ == "__main__":
    main()
