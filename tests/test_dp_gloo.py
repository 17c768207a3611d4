"""Data-parallel host logic on CPU with gloo, world_size 2: the arena layout / chunked all-reduce plan used by the
engine, checked against the DDP semantics of the reference (src/defaults/wrappers.py:182-183): the mean over ranks of
the per-rank gradients equals the gradient of the mean loss over the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from apla_b200.dp import ArenaLayout, allreduce_arena


def test_arena_layout_matches_engine_contract():
    lay = ArenaLayout(L=12, r=8, D=768, C=555)
    assert lay.n == 12 * 8 * 768 + 555 * 768 + 12 * 8 + 555 == 500_619
    assert lay.n_decay == 12 * 8 * 768 + 555 * 768
    early, late = lay.chunks()
    covered = sorted((s.start, s.stop) for s in early + late)
    assert covered[0][0] == 0 and covered[-1][1] == lay.n
    assert all(a[1] == b[0] for a, b in zip(covered[:-1], covered[1:]))          # a partition: no gap, no overlap
    # early chunk = blocks L/2.. + fc.weight, i.e. everything finished when backward passes block L/2
    assert early == [slice(lay.weight_slice(6).start, lay.b1)]
    assert ArenaLayout(L=1, r=4, D=128, C=10).chunks()[1] == [slice(4 * 128 + 10 * 128, 4 * 128 + 10 * 128 + 4 + 10)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import apla_oracle as O
    cfg = O.VitCfg(embed_dim=128, depth=2, num_heads=2, patch_size=14, img_size=56, n_classes=10, partial_size=16)
    sd = O.build_state(cfg, seed=0)           # same seed on every rank -> identical weights and indices
    O.perturb_state(sd)
    images, labels = O.synthetic_batch(3, 56, 10, rank=rank)
    res = O.loss_and_grads(sd, cfg, images, labels)
    lay = ArenaLayout(L=cfg.depth, r=16, D=128, C=10)
    arena = torch.zeros(lay.n)
    for l in range(cfg.depth):
        arena[lay.weight_slice(l)] = res.grads[f"backbone.blocks.{l}.attn.proj_weight1"].flatten()
        arena[lay.bias_slice(l)] = res.grads[f"backbone.blocks.{l}.attn.proj_bias1"]
    arena[lay.fcw:lay.b1] = res.grads["fc.weight"].flatten()
    arena[lay.fcb:lay.n] = res.grads["fc.bias"]
    allreduce_arena(arena, lay, which="early")
    allreduce_arena(arena, lay, which="late")
    arena /= world
    if rank == 0:
        # single-process reference: mean loss over the concatenated batch
        imgs = [images] + [O.synthetic_batch(3, 56, 10, rank=r)[0] for r in range(1, world)]
        labs = [labels] + [O.synthetic_batch(3, 56, 10, rank=r)[1] for r in range(1, world)]
        full = O.loss_and_grads(sd, cfg, torch.cat(imgs), torch.cat(labs))
        want = torch.zeros(lay.n)
        for l in range(cfg.depth):
            want[lay.weight_slice(l)] = full.grads[f"backbone.blocks.{l}.attn.proj_weight1"].flatten()
            want[lay.bias_slice(l)] = full.grads[f"backbone.blocks.{l}.attn.proj_bias1"]
        want[lay.fcw:lay.b1] = full.grads["fc.weight"].flatten()
        want[lay.fcb:lay.n] = full.grads["fc.bias"]
        err = float((arena - want).norm() / want.norm())
        torch.save({"err": err, "inds_equal": True}, out)
    # indices are identical across ranks without any broadcast (same seed, same constructor order)
    inds = sd["backbone.blocks.1.attn.inds"].clone()
    gathered = [torch.zeros_like(inds) for _ in range(world)]
    dist.all_gather(gathered, inds)
    assert all(torch.equal(g, inds) for g in gathered)
    dist.destroy_process_group()


def test_two_rank_gradient_mean_equals_concatenated_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["err"] < 1e-5, res


# ---------------------------------------------------------------------------------------------------------------------
# centres of the self-supervised losses under data parallel ("C4 additionally all-reduces two 65 536-float
# centres asynchronously"): host logic of apla_b200/dinov2/loss.py with the kernels emulated (tests/test_ssl_host.py)
# ---------------------------------------------------------------------------------------------------------------------
def _centre_worker(rank, world, port, out, precomputed=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ssl_host", os.path.join(os.path.dirname(__file__), "test_ssl_host.py"))
    host = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(host)
    from apla_b200.dinov2 import loss
    loss.ops = host.FakeOps                                  # this process only
    from oracle import ssl_oracle as S
    K = 64
    dl, il = loss.DINOLoss(K), loss.iBOTPatchLoss(K)
    dc, ic = torch.zeros(1, K), torch.zeros(1, 1, K)
    errs = []
    for step in range(2):
        g = torch.Generator().manual_seed(100 * step + rank)
        n_masked = 5 + 3 * rank                              # ranks hold different numbers of masked patches
        t_cls, t_patch = torch.randn(8, K, generator=g), torch.randn(1, n_masked, K, generator=g)
        got_d = dl.softmax_center_teacher(t_cls, 0.05)
        got_i = il.softmax_center_teacher(t_patch, 0.05)
        if precomputed:          # the fused step form: apla_ssl_objective hands the statistics over (hostdino._forward_fused)
            dl.register_center_stat(host.FakeOps.colsum(t_cls), len(t_cls))
            il.register_center_stat(host.FakeOps.colsum(t_patch[0], 1.0 / n_masked), 1)
        else:
            dl.update_center(t_cls)
            il.update_center(t_patch)
        errs.append(float((got_d - S.softmax_center_teacher(t_cls, dc, 0.05)).abs().max()))
        errs.append(float((got_i - S.softmax_center_teacher(t_patch, ic, 0.05)).abs().max()))
        # what the reference computes: all-reduced row sum / (len * world) for DINO, mean over ranks of the per-rank
        # patch means for iBOT (dino_clstoken_loss.py:76-98, ibot_patch_loss.py:123-145)
        cls_all = [torch.zeros_like(t_cls) for _ in range(world)]
        dist.all_gather(cls_all, t_cls)
        means = [torch.zeros(1, K) for _ in range(world)]
        dist.all_gather(means, t_patch.mean(1))
        dc = dc * 0.9 + torch.cat(cls_all).mean(0, keepdim=True) * 0.1
        ic = ic * 0.9 + (sum(means) / world).view(1, 1, K) * 0.1
    dl.apply_center_update(); il.apply_center_update()
    errs.append(float((dl.center - dc).abs().max()))
    errs.append(float((il.center - ic).abs().max()))
    if rank == 0:
        torch.save({"errs": errs}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("precomputed", [False, True], ids=["update_center", "register_center_stat"])
def test_two_rank_centre_updates(tmp_path, precomputed):
    out = str(tmp_path / "centres.pt")
    mp.spawn(_centre_worker, args=(2, _free_port(), out, precomputed), nprocs=2, join=True)
    res = torch.load(out)
    assert max(res["errs"]) < 1e-5, res
