"""Self-supervised APLA adaptation step at BASELINE config C4 shapes on ONE GPU: ViT-L/14 student and
teacher over 2 global 224-px + 8 local 98-px crops per image, shared DINO / iBOT head with K = 65 536 prototypes, iBOT
masks on half of the global crops, KoLeo, teacher EMA.  Module-level path: `apla_b200.hostdino.SSLMetaArch` under PyTorch
autograd, `torch.optim.AdamW` over the trainable tensors (projection rows + head) as the reference's trainer does.
CUDA events over whole steps, one JSON line.  A tools/ measurement for DESIGN.md / profiles/ -- bench.py's line stays C2.
NOT YET RUN (round 1's GPU budget was spent before the SSL kernels existed).

    python tools/bench_ssl_step.py [--images 64] [--steps 5] [--warmup 2] [--arch vit_large] [--K 65536]"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apla_b200.config import AplaConfig  # noqa: E402
from apla_b200.hostdino import SSLMetaArch, build_dino_backbone  # noqa: E402
from apla_b200.hostvit import ARCHS  # noqa: E402

GFLOP_PER_IMAGE_C4 = 1534.0          # student 2 x 343.6 + 8 x 63.4, teacher fwd 2 x 162.0, DINO head ~16


def make_batch(B, n_local, global_px, local_px, patch, mask_prob=0.5, ratio=(0.1, 0.5), seed=1234):
    """Synthetic stand-in for `collate_data_and_cast` (dinov2_utils.py:21-62): every other global crop is masked at a
    ratio drawn uniformly from `ratio` (uniform random positions instead of the block-wise generator)."""
    g = torch.Generator().manual_seed(seed)
    P = (global_px // patch) ** 2
    glob = torch.randn(2 * B, 3, global_px, global_px, generator=g)
    loc = torch.randn(n_local * B, 3, local_px, local_px, generator=g)
    masks = torch.zeros(2 * B, P, dtype=torch.bool)
    for i in range(2 * B):
        if float(torch.rand(1, generator=g)) < mask_prob:
            n = int(P * (ratio[0] + (ratio[1] - ratio[0]) * float(torch.rand(1, generator=g))))
            masks[i, torch.randperm(P, generator=g)[:n]] = True
    idx = masks.flatten().nonzero().flatten()
    mw = (1 / masks.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(masks)[masks]
    return {"collated_global_crops": glob, "collated_local_crops": loc, "collated_masks": masks,
            "mask_indices_list": idx, "masks_weight": mw, "upperbound": int(idx.shape[0]),
            "n_masked_patches": torch.full((1,), idx.shape[0], dtype=torch.long)}


def build(arch, K, partial_size, global_px, patch, hidden, bottleneck, n_local, device, fused_objective=True):
    from apla_b200.dinov2 import DINOHead
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        student = build_dino_backbone(arch, img_size=global_px, patch_size=patch, apla_config=AplaConfig(partial_size))
    inds = [b.attn.inds.clone() for b in student.blocks]
    with contextlib.redirect_stdout(io.StringIO()):
        teacher = build_dino_backbone(arch, img_size=global_px, patch_size=patch, apla_config=AplaConfig(partial_size),
                                      indices=inds)
    teacher.load_state_dict(student.state_dict())                     # models.py:138
    D = ARCHS[arch].embed_dim if isinstance(arch, str) else arch.embed_dim
    sh = DINOHead(D, K, nlayers=3, hidden_dim=hidden, bottleneck_dim=bottleneck)
    th = DINOHead(D, K, nlayers=3, hidden_dim=hidden, bottleneck_dim=bottleneck)
    th.load_state_dict(sh.state_dict())
    return SSLMetaArch(student, teacher, sh, th, K, n_local_crops=n_local, fused_objective=fused_objective).to(device)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--arch", default="vit_large")
    ap.add_argument("--K", type=int, default=65536)
    ap.add_argument("--partial-size", type=int, default=128)
    ap.add_argument("--n-local", type=int, default=8)
    ap.add_argument("--global-px", type=int, default=224)
    ap.add_argument("--local-px", type=int, default=98)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--per-term-losses", action="store_true",
                    help="the reference's call structure (one loss-class call per term) instead of the fused "
                         "head + objective node")
    a = ap.parse_args()
    model = build(a.arch, a.K, a.partial_size, a.global_px, 14, 2048, 256, a.n_local, a.device,
                  fused_objective=not a.per_term_losses)
    trainable = [p for p in model.student.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(trainable, lr=3e-5, weight_decay=1e-5)
    batch = make_batch(a.images, a.n_local, a.global_px, a.local_px, 14)
    batch = {k: (v.to(a.device) if torch.is_tensor(v) else v) for k, v in batch.items()}

    def step():
        opt.zero_grad(set_to_none=True)
        loss, parts = model(batch, teacher_temp=0.04)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(trainable, 3.0)
        opt.step()
        model.update_teacher(0.994)
        return loss

    for _ in range(a.warmup):
        loss = step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        loss = step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / a.steps
    full = a.arch == "vit_large" and a.K == 65536 and a.global_px == 224 and a.local_px == 98 and a.n_local == 8
    out = dict(workload=f"C4-shape SSL step: {a.arch}/14 student+teacher, {a.images} images -> {2 * a.images} x "
                        f"{a.global_px}px + {a.n_local * a.images} x {a.local_px}px crops, K={a.K}, r={a.partial_size}",
               objective="per-term loss classes" if a.per_term_losses else "fused head + apla_ssl_objective",
               ms_per_step=round(ms, 2), images_per_s=round(a.images / ms * 1e3, 1), loss=float(loss),
               masked_patches=int(batch["mask_indices_list"].shape[0]), trainable_params=sum(p.numel() for p in trainable),
               tflops=round(GFLOP_PER_IMAGE_C4 * a.images / ms, 1) if full else None,
               peak_hbm_gb=round(torch.cuda.max_memory_allocated() / 1e9, 1), when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
