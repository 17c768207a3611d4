"""Golden vectors for the Sinkhorn-Knopp teacher targets of the reference's DINO / iBOT losses
(src/self_supervised/dinov2/loss/dino_clstoken_loss.py:33-60, loss/ibot_patch_loss.py:53-83), UNMODIFIED reference, files
loaded by path like make_golden_ssl.py.  No shipped config selects `centering: sinkhorn_knopp`, but it is part of the two
classes' API.  The iBOT variant all-reduces its sample count unconditionally (ibot_patch_loss.py:59), so it only runs
inside a process group: a single-rank gloo group is initialised here.

    python tests/golden/make_golden_ssl_sk.py        # build container only
Output: tests/golden/ssl_sk_small.npz"""
import os

import numpy as np
import torch
import torch.distributed as dist

from make_golden_ssl import load

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    dino_m = load("loss/dino_clstoken_loss.py", "ref_dino_loss_sk")
    ibot_m = load("loss/ibot_patch_loss.py", "ref_ibot_loss_sk")
    g = torch.Generator().manual_seed(5)
    K = 128
    t_cls = torch.randn(12, K, generator=g) * 0.3
    t_patch = torch.randn(23, K, generator=g) * 0.3
    arrays = {"t_cls": t_cls.numpy(), "t_patch": t_patch.numpy()}
    arrays["dino_sk_t0.04_it3"] = dino_m.DINOLoss(K).sinkhorn_knopp_teacher(t_cls.clone(), 0.04).numpy()
    arrays["dino_sk_t0.07_it1"] = dino_m.DINOLoss(K).sinkhorn_knopp_teacher(t_cls.clone(), 0.07, n_iterations=1).numpy()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=0, world_size=1)
    arrays["ibot_sk_t0.04_it3"] = ibot_m.iBOTPatchLoss(K).sinkhorn_knopp_teacher(
        t_patch.clone(), 0.04, n_masked_patches_tensor=torch.full((1,), 23, dtype=torch.long)).numpy()
    arrays["dino_sk_t0.04_it3_in_group"] = dino_m.DINOLoss(K).sinkhorn_knopp_teacher(t_cls.clone(), 0.04).numpy()
    dist.destroy_process_group()
    np.savez_compressed(os.path.join(HERE, "ssl_sk_small.npz"), **arrays)
    print({k: v.shape for k, v in arrays.items()}, float(arrays["dino_sk_t0.04_it3"].sum(-1).mean()))


if __name__ == "__main__":
    main()
