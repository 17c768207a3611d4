"""A/B in one process: the step launched eagerly (~200 stream launches) vs replayed from the engine's CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200.config import AplaConfig  # noqa: E402
from apla_b200.engine import FineTuneEngine  # noqa: E402
from apla_b200.hostvit import build_classifier  # noqa: E402

B = 64
g = torch.Generator().manual_seed(1234)
images = torch.randn(B, 3, 224, 224, generator=g).cuda()
labels = torch.randint(0, 555, (B,), generator=g).cuda()
engs = {}
for use_graph in (False, True):
    model = build_classifier("vit_base", img_size=518, patch_size=14, n_classes=555, apla_config=AplaConfig(8), seed=0)
    engs[use_graph] = FineTuneEngine(model, batch_size=B, img_size=224, device="cuda:0", use_graph=use_graph)
for rnd in range(3):
    for use_graph, eng in engs.items():
        for _ in range(5):
            eng.step(images, labels)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            eng.step(images, labels)
        b.record()
        torch.cuda.synchronize()
        print(f"round {rnd} graph={use_graph}: {a.elapsed_time(b) / 20:.3f} ms/step  loss {float(eng.loss):.5f}", flush=True)
# same trajectory?
pa, pb = engs[False].params, engs[True].params
print("params rel diff eager vs graph after the same number of steps:", float((pa - pb).norm() / pa.norm()))
