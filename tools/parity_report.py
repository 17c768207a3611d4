"""Prints the measured parity of the native step against every golden fixture (for DESIGN.md / profiles)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CASES, build_case, cosine, rel, synthetic_batch  # noqa: E402
from apla_b200.engine import FineTuneEngine  # noqa: E402

print(f"{'case':24s} {'logits rel':>11s} {'loss rel':>10s} {'grad cos':>10s} {'grad rel':>10s} {'gnorm rel':>10s}")
for name in CASES:
    model, meta, arr = build_case(name)
    m = meta["meta"]
    eng = FineTuneEngine(model, batch_size=m["batch"], img_size=m["img"])
    images, labels = synthetic_batch(m["batch"], m["img"], m["n_classes"])
    eng.forward(images.cuda(), labels.cuda())
    eng.backward()
    eng.optim_step()
    torch.cuda.synchronize()
    g = eng.named_grads()
    sub = m["sub"]
    ours = torch.cat([g[k].flatten()[::sub].cpu() for k in meta["trainable"]])
    ref = torch.cat([torch.as_tensor(arr["s0/grad/" + k]).flatten() for k in meta["trainable"]])
    lr = abs(float(eng.loss) - float(arr["s0/loss"])) / abs(float(arr["s0/loss"]))
    gn = abs(float(eng.grad_norm()) - float(arr["s0/grad_norm"])) / float(arr["s0/grad_norm"])
    print(f"{name:24s} {rel(eng.logits, arr['s0/logits']):11.2e} {lr:10.2e} {cosine(ours, ref):10.6f} "
          f"{rel(ours, ref):10.2e} {gn:10.2e}")
    del eng
