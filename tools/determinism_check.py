"""Run the same forward + backward several times (no optimiser step) and report how much logits / loss / gradients
move between runs.  Forward must be bit-identical; the APLA weight gradient uses split-K fp32 atomics, so gradients
may differ in the last bits only (relative 1e-6), never at the bf16 level."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200.config import AplaConfig  # noqa: E402
from apla_b200.engine import FineTuneEngine  # noqa: E402
from apla_b200.hostvit import build_classifier  # noqa: E402

B = int(os.environ.get("B", 64))
R = int(os.environ.get("R", 8))
model = build_classifier("vit_base", img_size=518, patch_size=14, n_classes=555, apla_config=AplaConfig(R), seed=0)
eng = FineTuneEngine(model, batch_size=B, img_size=224, device="cuda:0")
g = torch.Generator().manual_seed(1234)
images = torch.randn(B, 3, 224, 224, generator=g).cuda()
labels = torch.randint(0, 555, (B,), generator=g).cuda()
runs = []
for i in range(int(os.environ.get("N", 6))):
    eng.forward(images, labels)
    eng.backward()
    torch.cuda.synchronize()
    runs.append((eng.logits.clone(), eng.loss.clone(), eng.grads.clone(),
                 [eng.xs[j].clone() for j in (1, 2, 24)]))
l0, s0, g0, x0 = runs[0]
for i, (l, s, gr, xs) in enumerate(runs[1:], 1):
    dl = float((l - l0).abs().max())
    dg = float((gr - g0).norm() / g0.norm())
    dx = [float((a - b).abs().max()) for a, b in zip(xs, x0)]
    print(f"run {i}: logits max|diff| {dl:.3e}  loss diff {float((s - s0).abs()):.3e}  grads rel diff {dg:.3e}  "
          f"xs[1,2,24] max|diff| {dx}")
