"""BASELINE.json's full size on the B200 (configs[1]: ViT-B/14, 224 px, 64 images per GPU, partial_size 8, 555 classes)
checked through properties that need no CPU oracle -- the oracle takes minutes at this size, the golden / live-oracle
cases of test_engine_gpu.py cover the arithmetic at small sizes, and these cover the SIZE: 16 448 tokens are 65 row
tiles of the GEMMs (the last one 64 rows deep), 768 (image, head) groups are 5.19 rounds of the persistent attention
kernels, and the tensors no longer fit L2.

  * images are independent: the logits of an image do not depend on which other images share its batch, and the loss /
    the trainable gradients of the batch are the means of those of its two halves (what DDP relies on, wrappers.py:182);
  * the step is deterministic where it claims to be (logits and residual stream bit-identical between runs);
  * nothing frozen moves, everything trainable does, and the dense bf16 projection the next forward reads is the
    refreshed one (appla_attn.py:64-79: the trainable rows sit at their index positions).
Bars: those of BASELINE.json's north_star (relative error <= 1e-2, gradient cosine >= 0.999)."""
import pytest
import torch

from helpers import cosine, rel

pytestmark = pytest.mark.gpu

B, IMG, NCLS, R = 64, 224, 555, 8


def _build(batch):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from apla_b200.config import AplaConfig
    from apla_b200.engine import FineTuneEngine
    from apla_b200.hostvit import build_classifier
    model = build_classifier("vit_base", img_size=518, patch_size=14, n_classes=NCLS, apla_config=AplaConfig(R), seed=0)
    return model, FineTuneEngine(model, batch_size=batch, img_size=IMG, device="cuda:0")


def _batch(seed=77):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, IMG, IMG, generator=g).cuda(), torch.randint(0, NCLS, (B,), generator=g).cuda()


def _flat_grads(eng):
    g = eng.named_grads()
    return torch.cat([g[k].flatten().float() for k in eng.trainable_names()])


@pytest.fixture(scope="module")
def full():
    model, eng = _build(B)
    images, labels = _batch()
    eng.forward(images, labels)
    eng.backward()
    torch.cuda.synchronize()
    out = dict(model=model, eng=eng, images=images, labels=labels, logits=eng.logits.clone(), loss=float(eng.loss),
               grads=_flat_grads(eng).clone(), xs_last=eng.xs[-1].clone())
    yield out
    del eng


def test_full_size_outputs_are_finite_and_sized(full):
    eng = full["eng"]
    assert eng.shape["B"] * eng.shape["N"] == 16448 and eng.shape["D"] == 768 and eng.shape["L"] == 12
    assert full["logits"].shape == (B, NCLS) and torch.isfinite(full["logits"]).all()
    assert torch.isfinite(full["grads"]).all() and float(full["grads"].norm()) > 0
    assert abs(full["loss"] - float(torch.nn.functional.cross_entropy(full["logits"].float(), full["labels"]))) < 1e-3


def test_full_size_step_is_deterministic(full):
    eng = full["eng"]
    eng.forward(full["images"], full["labels"])
    eng.backward()
    torch.cuda.synchronize()
    assert torch.equal(eng.logits, full["logits"]) and torch.equal(eng.xs[-1], full["xs_last"])
    assert rel(_flat_grads(eng), full["grads"]) < 1e-5         # the split-K weight gradient adds with fp32 atomics


def test_images_are_independent_and_halves_average(full):
    """Two engines of 32 images on the two halves of the batch (other row-tile counts, other attention rounds)."""
    _, half = _build(B // 2)
    logits, losses, grads = [], [], []
    for i in range(2):
        sl = slice(i * (B // 2), (i + 1) * (B // 2))
        half.forward(full["images"][sl].contiguous(), full["labels"][sl].contiguous())
        half.backward()
        torch.cuda.synchronize()
        logits.append(half.logits.clone())
        losses.append(float(half.loss))
        grads.append(_flat_grads(half).clone())
    logits = torch.cat(logits)
    assert rel(logits, full["logits"]) < 2e-3, rel(logits, full["logits"])        # same arithmetic, other tiling
    assert abs(0.5 * (losses[0] + losses[1]) - full["loss"]) < 1e-3 * abs(full["loss"])
    mean_g = 0.5 * (grads[0] + grads[1])
    assert cosine(mean_g, full["grads"]) >= 0.999, cosine(mean_g, full["grads"])
    assert rel(mean_g, full["grads"]) <= 1e-2, rel(mean_g, full["grads"])


def test_batch_permutation_permutes_the_logits(full):
    eng = full["eng"]
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3)).cuda()
    eng.forward(full["images"][perm].contiguous(), full["labels"][perm].contiguous())
    eng.backward()
    torch.cuda.synchronize()
    assert rel(eng.logits, full["logits"][perm]) < 2e-3
    assert abs(float(eng.loss) - full["loss"]) < 1e-3 * abs(full["loss"])
    g = _flat_grads(eng)
    assert cosine(g, full["grads"]) >= 0.999 and rel(g, full["grads"]) <= 1e-2


def test_only_the_trainable_tensors_move(full):
    """Five optimiser steps at full size: frozen tensors bit-identical, every trainable tensor changed, and the dense bf16
    projection copies hold bf16(proj_weight1) at the index rows and the frozen rows everywhere else."""
    model, eng = full["model"], full["eng"]
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    trainable = set(eng.trainable_names())
    p0 = eng.params.clone()
    for _ in range(5):
        eng.step(full["images"], full["labels"])
    torch.cuda.synchronize()
    assert float((eng.params - p0).abs().max()) > 0
    eng.sync_to_model()
    after = model.state_dict()
    for k, v in before.items():
        if k in trainable:
            assert not torch.equal(after[k].cpu(), v.cpu()), f"{k} did not move"
        else:
            assert torch.equal(after[k].cpu(), v.cpu()), f"frozen tensor {k} changed"
    # the engine's working copy of block 5's projection against the module's parameters
    blk = model.backbone.blocks[5].attn
    inds = torch.as_tensor(blk.indices).long()
    dense = torch.zeros(768, 768)
    dense[inds[:R]] = blk.proj_weight1.detach().cpu().float()
    dense[inds[R:]] = blk.proj_weight2.detach().cpu().float()
    got = eng._wproj_all[5].float().cpu()
    assert torch.equal(got, dense.bfloat16().float())
    assert torch.equal(eng._wprojT_all[5].float().cpu(), dense.bfloat16().float().t())
