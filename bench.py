#!/usr/bin/env python
"""bench.py -- the APLA fine-tune step on B200 (BASELINE.json configs[1]: ViT-B/14, partial_size=8, 224 px,
555 classes, batch 64 per GPU, bf16 compute, synthetic data, seeded random-init weights of that architecture).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the reference algorithm on the host CPU cores

One step = Trainer.global_step (src/defaults/trainer.py:106-138): forward, CrossEntropy, backward (weight
gradients only for the APLA projection rows + head), data-parallel gradient mean, clip_grad_norm_(1.0), AdamW.
Prints ONE JSON line (rank 0).  `value` = whole-job images/s with the batch resident in HBM; `e2e` = the same
through the public API from pinned host buffers (H2D of the batch and D2H of the loss inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ViT-B/14 APLA fine-tune images/sec/GPU at 1/2/4/8 B200; % bf16 tensor roofline"
WORKLOAD = dict(arch="vit_base", img=224, table_img=518, patch=14, n_classes=555, partial_size=8, batch_per_gpu=64)
# --workload: the default is the configuration the metric is quoted on (BASELINE.json configs[1] = C2); the others are
# the remaining single-box shapes of SURVEY.md App. A through the SAME engine (reported in DESIGN.md, not the bench line
# the driver reads).  c5 / vitl end in the classifier head: the mmseg decoder and the DINOv2 SSL heads are out of scope.
WORKLOADS = {
    "c2": WORKLOAD,
    "c3": dict(WORKLOAD, partial_size=768),
    "c5": dict(WORKLOAD, img=518, partial_size=768, batch_per_gpu=2),
    "vitl": dict(WORKLOAD, arch="vit_large", partial_size=128),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops_burst=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm_gbs=d["hbm_gbs"], source="measured")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML, else nvidia-smi)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {}
        if nv is not None:
            for k in dir(nv):
                if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason"):
                    v = getattr(nv, k)
                    if isinstance(v, int) and v not in (0,):
                        names.setdefault(v, k.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, nm in names.items():
                        if mask & bit and bit & (bit - 1) == 0:
                            self.reasons.add(nm)
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}",
                                          "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                    f = [x.strip() for x in out.stdout.strip().split(",")]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    self.reasons.add(f[2])
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        s = sorted(self.samples)
        rs = sorted(r for r in self.reasons if r and r.lower() not in ("none", "gpuidle", "applicationsclockssetting"))
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=rs, samples=len(s))


def synthetic_batch(batch, img, n_classes, seed=1234, rank=0):
    import torch
    g = torch.Generator().manual_seed(seed + rank)
    return torch.randn(batch, 3, img, img, generator=g), torch.randint(0, n_classes, (batch,), generator=g)


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, batch, w=WORKLOAD):
    """Times oracle.fine_tune_step (fp32 restatement of the reference step) on `batch` images of the workload."""
    import torch
    from oracle import apla_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = O.VitCfg(**(O.VIT_L14 if w["arch"] == "vit_large" else O.VIT_B14), n_classes=w["n_classes"],
                   partial_size=w["partial_size"])
    sd = O.build_state(cfg, seed=0)
    images, labels = O.synthetic_batch(batch, w["img"], w["n_classes"])
    st = O.AdamWState()
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.fine_tune_step(sd, cfg, images, labels, st)
        ts.append(time.perf_counter() - t0)
    ts = ts[warmup:]
    sec = sum(ts) / len(ts)
    return dict(value=batch / sec, unit="images/s", cores=threads, kind="port",
                sample=f"oracle/apla_oracle.py fine_tune_step, {w['arch']}/14 r={w['partial_size']} {w['img']}px fp32, "
                       f"batch {batch}, "
                       f"{len(ts)} steps after {warmup} warm-up, {sec * 1e3:.0f} ms/step"), sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    w = WORKLOADS[args.workload]
    cb, sec = cpu_reference_steps(steps, warmup, batch=min(8, w["batch_per_gpu"]), w=w)
    line = dict(metric=METRIC, value=cb["value"], unit="images/s", n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=f"{w['arch']}/14 dinov2-arch APLA partial_size={w['partial_size']} fine-tune step, "
                                     f"{w['img']}px, 555 classes; CPU sample: batch {min(8, w['batch_per_gpu'])} per step",
                            **w),
                cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
# DRAM bytes (read + write) per launch of the dominant kernel from `ncu --set full` (profiles/ncu_r1d_summary.md):
# qkv 50.0 MB, fc1-dgrad 119.1 MB, qkv-dgrad 86.6 MB, weighted by 12 / 12 / 11 launches per step.
DOMINANT_TRAFFIC_BYTES = (12 * 50.0e6 + 12 * 119.1e6 + 11 * 86.6e6) / 35


def time_dominant_kernel(eng, torch, rounds=3):
    """CUDA-event timing (torch's current stream = the launching stream) of the kernel with the largest share of the
    step, gemm2_kernel<256, EPI_BIAS> (21.6 % in profiles/launches_r1d_summary.txt): the plain 2-CTA tcgen05 GEMM that
    runs qkv forward [T,768]x[768,2304], fc1 dgrad [T,3072]x[3072,768] and qkv dgrad [T,2304]x[2304,768] -- timed
    on exactly those shapes in the step's 12 / 12 / 11 proportion, rotating over 4 operand sets so they come from
    HBM / L2 like in the step.  Also times the fc1 + GELU epilogue kernel (the single most expensive launch)."""
    from apla_b200 import ops
    sh = eng.shape
    T, D, Hd, L = sh["T"], sh["D"], sh["hidden"], sh["L"]
    bf = torch.bfloat16

    def mats(n, k):
        return [(torch.randn(n, k, device="cuda") * 0.02).to(bf) for _ in range(4)]
    x_d = [torch.randn(T, D, device="cuda").to(bf) for _ in range(2)]
    x_h = [torch.randn(T, Hd, device="cuda").to(bf) for _ in range(2)]
    x_q = [torch.randn(T, 3 * D, device="cuda").to(bf) for _ in range(2)]
    w_qkv, w_fc1t, w_qkvt = mats(3 * D, D), mats(D, Hd), mats(D, 3 * D)
    o_q = [torch.empty(T, 3 * D, device="cuda", dtype=bf) for _ in range(2)]
    o_d = [torch.empty(T, D, device="cuda", dtype=bf) for _ in range(2)]
    bias = torch.zeros(3 * D, device="cuda")
    plan = []      # (callable, flops) in the step's proportion
    for i in range(L):
        plan.append((lambda i=i: ops.gemm_bias(x_d[i % 2], w_qkv[i % 4], bias, out=o_q[i % 2]), 2.0 * T * 3 * D * D))
        plan.append((lambda i=i: ops.gemm_dgrad(x_h[i % 2], w_fc1t[i % 4], out=o_d[i % 2]), 2.0 * T * D * Hd))
        if i:
            plan.append((lambda i=i: ops.gemm_dgrad(x_q[i % 2], w_qkvt[i % 4], out=o_d[i % 2]), 2.0 * T * D * 3 * D))
    for fn, _ in plan:
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rounds):
        for fn, _ in plan:
            fn()
    b.record()
    torch.cuda.synchronize()
    sec = a.elapsed_time(b) * 1e-3 / (rounds * len(plan))
    flops = sum(f for _, f in plan) / len(plan)
    dom = dict(kernel="g2::gemm2_kernel<256, EPI_BIAS> (qkv fwd / fc1 dgrad / qkv dgrad, 12:12:11)", launches=len(plan),
               us_per_launch=sec * 1e6, flops_per_launch=flops, tflops=flops / sec / 1e12)
    # the fc1 + GELU + saved-derivative epilogue kernel
    ws = mats(Hd, D)
    hs = [torch.empty(T, Hd, device="cuda", dtype=torch.float16) for _ in range(2)]
    gs = [torch.empty(T, Hd, device="cuda", dtype=bf) for _ in range(2)]
    b1 = torch.zeros(Hd, device="cuda")
    for i in range(4):
        ops.gemm_bias_gelu_dgelu(x_d[i % 2], ws[i], b1, d=hs[i % 2], g=gs[i % 2])
    torch.cuda.synchronize()
    a.record()
    for i in range(24):
        ops.gemm_bias_gelu_dgelu(x_d[i % 2], ws[i % 4], b1, d=hs[i % 2], g=gs[i % 2])
    b.record()
    torch.cuda.synchronize()
    sec1 = a.elapsed_time(b) * 1e-3 / 24
    fc1 = dict(kernel="g2::gemm2_kernel<256, EPI_BIAS_GELU_D> (Mlp.fc1 + GELU + saved GELU', 16448x3072x768)",
               us_per_launch=sec1 * 1e6, tflops=2.0 * T * D * Hd / sec1 / 1e12)
    return dom, fc1


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from apla_b200._lib import LIB
    from apla_b200.config import AplaConfig
    from apla_b200.engine import FineTuneEngine
    from apla_b200.flops import flops_per_image
    from apla_b200.hostvit import ARCHS, build_classifier

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.gpus != world:
        raise RuntimeError(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 through torch.distributed.run")
    w = WORKLOADS[args.workload]
    B = args.batch or w["batch_per_gpu"]

    # identical weights and APLA indices on every rank: same seed, same constructor order
    model = build_classifier(w["arch"], img_size=w["table_img"], patch_size=w["patch"], n_classes=w["n_classes"],
                             apla_config=AplaConfig(w["partial_size"]), seed=0)
    eng = FineTuneEngine(model, batch_size=B, img_size=w["img"], device=f"cuda:{local}")
    images, labels = synthetic_batch(B, w["img"], w["n_classes"], rank=rank)
    images_pin, labels_pin = images.pin_memory(), labels.pin_memory()
    images_dev, labels_dev = images.cuda(), labels.cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    sampler = ClockSampler(local) if rank == 0 else None
    n0 = LIB.load().apla_launch_count()
    eng.step(images_dev, labels_dev)
    launches_per_step = int(LIB.load().apla_launch_count() - n0)
    if sampler:
        sampler.start()
    # (the engine runs a buffer pair eagerly once, captures its CUDA graph on the second step and replays from the third:
    #  at least two untimed steps here, four for the two input slots of the end-to-end path)
    ms_step = timed(lambda: eng.step(images_dev, labels_dev), args.steps, max(2, args.warmup - 1))
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(lambda: eng.step_from_host(images_pin, labels_pin), args.steps, max(4, args.warmup // 2))
    loss = eng.drain()
    if loss is None:
        loss = float(eng.loss.item())

    if rank == 0:
        peaks = load_peaks()
        a = ARCHS[w["arch"]]
        f_fwd, f_bwd = flops_per_image(a.embed_dim, a.depth, w["patch"], w["img"], w["partial_size"], w["n_classes"])
        flops_step_ref = (f_fwd + f_bwd) * B                      # the reference's dense algorithm
        e_fwd, e_bwd = flops_per_image(a.embed_dim, a.depth, w["patch"], w["img"], w["partial_size"], w["n_classes"],
                                       cls_only_last_block=int(eng.cls_only_last_block))
        flops_step = (e_fwd + e_bwd) * B                          # what this engine executes (last block: CLS rows)
        achieved = flops_step / (ms_step * 1e-3) / 1e12
        dom, fc1 = time_dominant_kernel(eng, torch)
        line = dict(
            metric=METRIC, value=B * world / (ms_step * 1e-3), unit="images/s", n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="bf16", data="synthetic", per_gpu=B / (ms_step * 1e-3), loss=loss,
            config=dict(workload=f"{'ViT-L' if w['arch'] == 'vit_large' else 'ViT-B'}/14 dinov2-arch (518-px pos table, "
                                 f"LayerScale, qkv bias) APLA partial_size={w['partial_size']} supervised fine-tune step, "
                                 f"{w['img']}px, 555 classes, AdamW lr 3e-5 wd 1e-5 clip 1.0",
                        global_batch=B * world, batch_per_gpu=B, tokens_per_image=eng.shape["N"],
                        parallelism=f"dp{world}",
                        l2=f"per-step working set ~{sum(t.numel() * t.element_size() for t in eng._keep) / 1e9:.1f} GB "
                           ">> 126 MB L2 (saved activations of every block), no flush needed",
                        **{k: v for k, v in w.items() if k != "batch_per_gpu"}),
            # dominant kernel (largest share of the step), algorithmic FLOPs 2*M*N*K per launch / live CUDA-event time;
            # peak = measured burst cuBLAS bf16 (kernel timed alone); `step` = the whole step against the sustained peak
            roofline=dict(bound="tensor", kernel=dom["kernel"], achieved=dom["tflops"], peak=peaks["tflops_burst"],
                          unit="TFLOP/s", frac=dom["tflops"] / peaks["tflops_burst"],
                          peak_source=f"{peaks['source']} MEASURED_PEAKS.json bf16_tflops (burst)",
                          us_per_launch=dom["us_per_launch"], flops_per_launch=dom["flops_per_launch"],
                          launches_timed=dom["launches"],
                          traffic=DOMINANT_TRAFFIC_BYTES if (args.workload in ("c2", "c3") and B == 64) else None,
                          traffic_note="dram read+write bytes per launch, ncu --set full (profiles/ncu_r1d_summary.md)",
                          fc1_gelu_kernel=dict(**fc1, frac=fc1["tflops"] / peaks["tflops_burst"]),
                          step=dict(scope="whole step (all kernels); EXECUTED algorithmic FLOPs: SURVEY.md App. B minus the "
                                          "last block's per-token work on non-CLS tokens, which the engine proves dead "
                                          "and does not run (DESIGN.md section 5)",
                                    achieved=achieved, peak=peaks["tflops_sustained"], unit="TFLOP/s",
                                    frac=achieved / peaks["tflops_sustained"],
                                    frac_of_burst=achieved / peaks["tflops_burst"], flops_per_step=flops_step,
                                    flops_per_step_reference_dense=flops_step_ref,
                                    frac_reference_dense_flops=flops_step_ref / (ms_step * 1e-3) / 1e12
                                    / peaks["tflops_sustained"])),
            e2e=dict(value=B * world / (ms_e2e * 1e-3), unit="images/s", ms_per_step=ms_e2e,
                     h2d_bytes_per_step=images_pin.numel() * 4 + labels_pin.numel() * 8, d2h_bytes_per_step=4),
            gpu_launches=launches_per_step * args.steps, gpu_launches_per_step=launches_per_step, clocks=clocks)
        if world == 1 and not args.no_cpu:
            cb, _ = cpu_reference_steps(steps=3, warmup=1, batch=min(8, w["batch_per_gpu"]), w=w)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (default: the workload's 64)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS),
                    help="c2 = BASELINE.json configs[1] (the metric's configuration, default); c3 / c5 / vitl: other shapes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
