"""Per-tensor gradient error of one golden case (default: c3_vitb14_r768): where along the depth the bf16 noise sits."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_case, cosine, rel, synthetic_batch  # noqa: E402
from apla_b200.engine import FineTuneEngine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3_vitb14_r768"
model, meta, arr = build_case(name)
m = meta["meta"]
eng = FineTuneEngine(model, batch_size=m["batch"], img_size=m["img"])
images, labels = synthetic_batch(m["batch"], m["img"], m["n_classes"])
eng.forward(images.cuda(), labels.cuda())
eng.backward()
torch.cuda.synchronize()
g = eng.named_grads()
sub = m["sub"]
tot_ref = 0.0
rows = []
for k in meta["trainable"]:
    ours = g[k].flatten()[::sub].cpu()
    ref = torch.as_tensor(arr["s0/grad/" + k]).flatten()
    rows.append((k, float(ref.norm()), rel(ours, ref), cosine(ours, ref), float((ours.double() - ref.double()).norm())))
    tot_ref += float(ref.norm()) ** 2
print(f"{'tensor':48s} {'|ref|':>10s} {'rel':>9s} {'cos':>9s} {'share of err^2':>14s}")
tot_err = sum(r[4] ** 2 for r in rows)
for k, n, r, c, e in rows:
    print(f"{k:48s} {n:10.3e} {r:9.2e} {c:9.6f} {100 * e * e / tot_err:13.1f}%")
print("global rel", (tot_err / tot_ref) ** 0.5)
