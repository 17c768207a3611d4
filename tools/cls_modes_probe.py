"""A/B in one process: step time of the engine with the last block evaluated on every token (0), with its per-token
tail on the CLS rows only (1), and with its attention for the CLS query only as well (2, default)."""
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
for mode in (0, 1, 2):
    model = build_classifier("vit_base", img_size=518, patch_size=14, n_classes=555, apla_config=AplaConfig(8), seed=0)
    engs[mode] = FineTuneEngine(model, batch_size=B, img_size=224, device="cuda:0", cls_only_last_block=mode)
for rnd in range(3):
    for mode, eng in engs.items():
        for _ in range(5):
            eng.step(images, labels)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            eng.step(images, labels)
        b.record()
        torch.cuda.synchronize()
        print(f"round {rnd} mode {mode}: {a.elapsed_time(b) / 20:.3f} ms/step", flush=True)
