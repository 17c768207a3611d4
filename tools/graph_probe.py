"""Experiment: how much of the step is launch gaps?  Times the eager step (222 stream launches) against the same
launch sequence replayed from one CUDA graph (hyper-parameters frozen, timing only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200.config import AplaConfig  # noqa: E402
from apla_b200.engine import FineTuneEngine  # noqa: E402
from apla_b200.hostvit import build_classifier  # noqa: E402

B = int(os.environ.get("B", 64))
model = build_classifier("vit_base", img_size=518, patch_size=14, n_classes=555, apla_config=AplaConfig(8), seed=0)
eng = FineTuneEngine(model, batch_size=B, img_size=224, device="cuda:0")
g = torch.Generator().manual_seed(1234)
images = torch.randn(B, 3, 224, 224, generator=g).cuda()
labels = torch.randint(0, 555, (B,), generator=g).cuda()


def timed(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t_eager = timed(lambda: eng.step(images, labels))
t_fwd = timed(lambda: eng.forward(images, labels))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    eng.step(images, labels)
torch.cuda.current_stream().wait_stream(s)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    eng.step(images, labels)
t_graph = timed(graph.replay)
print(f"eager step {t_eager:.3f} ms   forward only {t_fwd:.3f} ms   graph replay {t_graph:.3f} ms   "
      f"({B / t_graph * 1e3:.0f} img/s)")
