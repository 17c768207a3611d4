#!/usr/bin/env python
"""Module-level paths on one B200, ViT-B/14 224 px batch 64 (the bench.py workload), fwd + CE + bwd + clip + AdamW driven
by plain PyTorch autograd / torch.optim -- i.e. what a user gets who keeps the reference's trainer:

  eager_full   the host ViT with STOCK attention and attn.proj trainable (the reference's multi-GPU 'full' mode,
               apla_vit.py:65-75) in PyTorch eager under torch.autocast(bf16): the reference's own algorithm
               (N x N probabilities in HBM, ~10 ATen kernels around every GEMM) on the same GPU
  attn_only    build_apla(...): only the attention module replaced by the fused APLA_Attention, block glue in eager fp32/bf16
  fused_block  + fuse_apla_blocks(model): every block one autograd node
  fused_all    + fuse_patch_embed(model) + cache_pos_encoding(model): frozen stem on the library's kernels, position table
               resized once
  (the step engine = bench.py)

Not a bench.py substitute: informational numbers for DESIGN.md.  Prints one JSON line per path."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200.apla import fuse_apla_blocks            # noqa: E402
from apla_b200.config import AplaConfig                # noqa: E402
from apla_b200.hostvit import build_classifier         # noqa: E402


def groups(model):
    reg, noreg = [], []
    for n, p in model.named_parameters():
        if p.requires_grad:
            (noreg if (n.endswith(".bias") or p.dim() == 1) else reg).append(p)
    return [{"params": reg}, {"params": noreg, "weight_decay": 0.0}]


def run(name, model, B, steps, warmup, autocast):
    model.cuda().train()
    opt = torch.optim.AdamW(groups(model), lr=3e-5, weight_decay=1e-5)
    g = torch.Generator().manual_seed(1234)
    images = torch.randn(B, 3, 224, 224, generator=g).cuda()
    labels = torch.randint(0, 555, (B,), generator=g).cuda()
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = torch.nn.functional.cross_entropy(model(images).float(), labels)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print(json.dumps(dict(path=name, images_per_s=B / ms * 1e3, ms_per_step=ms, batch=B, steps=steps, loss=float(loss),
                          peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)), flush=True)
    torch.cuda.reset_peak_memory_stats()


def run_c4_backbone(images_n, steps, warmup):
    """C4's block path: the DINOv2 student's 24 ViT-L/14 blocks over packed multi-crop input (2 global 224-px crops =
    257 tokens and 8 local 98-px crops = 50 tokens per image, ISIC2019 config) forward + backward through the fused
    blocks (partial_size = dim: every projection row trainable, the SSL configs' 'full'), and the teacher's forward over
    the global crops.  Heads, losses and the patch embedding are outside SURVEY 8 (row f2): the loss gradient is a
    random tensor.  Prints crops-per-second and the algorithmic TFLOP/s of the block path."""
    m = build_classifier("vit_large", apla_config=AplaConfig(1024), img_size=518, patch_size=14, n_classes=8, seed=0)
    fuse_apla_blocks(m.cuda())
    blocks = m.backbone.blocks
    D, L = 1024, len(blocks)
    g = torch.Generator(device="cuda").manual_seed(0)
    glob = torch.randn(2 * images_n, 257, D, device="cuda", generator=g)
    loc = torch.randn(8 * images_n, 50, D, device="cuda", generator=g)
    dg, dl = torch.randn_like(glob) * 1e-3, torch.randn_like(loc) * 1e-3

    def step():
        for p in m.parameters():
            p.grad = None
        xs = [glob.clone().requires_grad_(True), loc.clone().requires_grad_(True)]     # tokens after a trainable-free embed
        for blk in blocks:
            xs = blk(xs)
        torch.autograd.backward(xs, [dg, dl])
        with torch.no_grad():                       # teacher: forward only, global crops
            t = glob
            for blk in blocks:
                t = blk(t)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    lin = 24 * D * D
    tok = lambda n, cnt: cnt * n * (lin + 4 * n * D)                                   # noqa: E731  forward FLOPs
    f_student = tok(257, 2 * images_n) + tok(50, 8 * images_n)
    b_student = sum(cnt * n * (lin + 2.5 * 4 * n * D + 2 * D * D) for n, cnt in ((257, 2 * images_n), (50, 8 * images_n)))
    flops = L * (f_student + b_student + tok(257, 2 * images_n))
    print(json.dumps(dict(path="c4_block_path_vitl", images=images_n, student_tokens=2 * images_n * 257 + 8 * images_n * 50,
                          teacher_tokens=2 * images_n * 257, ms_per_step=ms, images_per_s=images_n / ms * 1e3,
                          tflops=flops / ms / 1e9, peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--paths", default="eager_full,attn_only,fused_block,fused_all,fused_block_r768,fused_all_r768")
    a = ap.parse_args()
    kw = dict(img_size=518, patch_size=14, n_classes=555, seed=0)
    for p in a.paths.split(","):
        if p == "eager_full":
            m = build_classifier("vit_base", apla_config=AplaConfig("full"), is_multi_gpu=True, **kw)
            run(p, m, a.batch, a.steps, a.warmup, autocast=True)
        elif p == "attn_only":
            m = build_classifier("vit_base", apla_config=AplaConfig(8), **kw)
            run(p, m, a.batch, a.steps, a.warmup, autocast=True)
        elif p == "c4":
            run_c4_backbone(a.batch, a.steps, a.warmup)
            continue
        elif p in ("fused_block", "fused_block_r768"):
            m = build_classifier("vit_base", apla_config=AplaConfig(768 if p.endswith("768") else 8), **kw)
            run(p, fuse_apla_blocks(m), a.batch, a.steps, a.warmup, autocast=False)
        elif p in ("fused_all", "fused_all_r768"):
            # + the frozen stem as patchify + tcgen05 GEMM and the position-table resize computed once (round 2)
            from apla_b200.apla import cache_pos_encoding, fuse_patch_embed
            m = build_classifier("vit_base", apla_config=AplaConfig(768 if p.endswith("768") else 8), **kw)
            run(p, cache_pos_encoding(fuse_patch_embed(fuse_apla_blocks(m))), a.batch, a.steps, a.warmup, autocast=False)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
