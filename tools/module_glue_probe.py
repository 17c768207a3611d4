"""Torch-native CUDA time inside one step of the module-level path (fuse_apla_blocks + fuse_patch_embed +
cache_pos_encoding under plain autograd + torch.optim.AdamW): which aten ops, how long, how often.  Development aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200.apla import cache_pos_encoding, fuse_apla_blocks, fuse_patch_embed  # noqa: E402
from apla_b200.config import AplaConfig  # noqa: E402
from apla_b200.hostvit import build_classifier  # noqa: E402
from tools.bench_module_paths import groups  # noqa: E402


def main():
    from torch.profiler import ProfilerActivity, profile
    B = int(os.environ.get("BATCH", "64"))
    m = build_classifier("vit_base", apla_config=AplaConfig(int(os.environ.get("R", "8"))), img_size=518, patch_size=14,
                         n_classes=555, seed=0)
    model = cache_pos_encoding(fuse_patch_embed(fuse_apla_blocks(m))).cuda().train()
    opt = torch.optim.AdamW(groups(model), lr=3e-5, weight_decay=1e-5)
    g = torch.Generator().manual_seed(1234)
    images = torch.randn(B, 3, 224, 224, generator=g).cuda()
    labels = torch.randint(0, 555, (B,), generator=g).cuda()
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(images).float(), labels)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step()
    b.record()
    torch.cuda.synchronize()
    print(f"step: {a.elapsed_time(b) / 10:.3f} ms (batch {B})")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages(group_by_input_shape=True):
        t = getattr(ev, "self_device_time_total", 0) or getattr(ev, "self_cuda_time_total", 0)
        if t > 0 and ev.key.startswith("aten::"):
            rows.append((t, ev.count, ev.key, str(ev.input_shapes)[:90]))
    print(f"aten ops with device time of their own: {sum(r[0] for r in rows) / 1e3:.3f} ms")
    for t, c, name, shp in sorted(rows, key=lambda r: -r[0])[:24]:
        print(f"{t / 1e3:8.3f} ms  x{c:<4d} {name:30s} {shp}")


if __name__ == "__main__":
    main()
