"""Data-parallel correctness of the step engine on REAL NCCL ("N-rank grads == 1-rank grads on the
concatenated batch, since DDP averages"; reference: DDP wrap src/defaults/wrappers.py:182-183).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py

Every rank builds the same tiny model, runs K engine steps on its own shard through the engine's three-graph + side-stream
all-reduce choreography, and checks
  1. after step 1: the all-reduced gradient arena / world == the gradient a 1-rank engine computes on the CONCATENATED
     batch (up to the fp32 atomics of the split-K weight gradient and bf16 noise of the different batch shape);
  2. after K steps: parameters and Adam moments are bit-identical on every rank;
  3. the K-step parameters track the 1-rank run on the concatenated batch.
Prints one JSON line on rank 0; exit status 1 on failure."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apla_b200.config import AplaConfig  # noqa: E402
from apla_b200.engine import FineTuneEngine  # noqa: E402
from apla_b200.hostvit import VitArch, build_classifier  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solo = [dist.new_group([r]) for r in range(world)]           # one-rank groups: the "1-rank" engine of the comparison
    arch, img, patch, C, B, K = VitArch(128, 4, 2), 56, 14, 10, 4, 6
    results = {}
    ok = True
    for r_apla, use_graph in ((16, True), (128, True), (16, False)):
        def make(batch, pg):
            m = build_classifier(arch, img_size=img, patch_size=patch, n_classes=C, apla_config=AplaConfig(r_apla), seed=0)
            return FineTuneEngine(m, batch_size=batch, img_size=img, device=f"cuda:{local}", process_group=pg, lr=1e-3,
                                  use_graph=use_graph)
        g = torch.Generator().manual_seed(99)
        images = torch.randn(world * B, 3, img, img, generator=g)
        labels = torch.randint(0, C, (world * B,), generator=g)
        mine = slice(rank * B, (rank + 1) * B)
        eng = make(B, None)                                       # default group: all ranks
        ref = make(world * B, solo[rank])                         # this rank alone on the concatenated batch
        xi, yi = images[mine].cuda(), labels[mine].cuda()
        xa, ya = images.cuda(), labels.cuda()
        # --- step 1, eager pieces: compare the reduced gradient with the concatenated-batch gradient
        eng.forward(xi, yi); eng.backward()
        ref.forward(xa, ya); ref.backward()
        torch.cuda.synchronize()
        g_dp = eng.grads / world
        gg = [torch.empty_like(eng.grads) for _ in range(world)]
        dist.all_gather(gg, eng.grads)
        o = eng.offsets
        regions = dict(w1=(o["w1"], o["fcw"]), fcw=(o["fcw"], o["b1"]), b1=(o["b1"], o["fcb"]), fcb=(o["fcb"], o["n"]))
        grad_diff = {k: float((gg[0][a:b] - gg[-1][a:b]).abs().max()) for k, (a, b) in regions.items()}
        e_grad = rel(g_dp, ref.grads)
        cos = float(torch.nn.functional.cosine_similarity(g_dp.double(), ref.grads.double(), dim=0))
        eng.optim_step(); ref.optim_step()
        # --- K more steps through step() (graphs captured on the way)
        for _ in range(K):
            eng.step(xi, yi)
            ref.step(xa, ya)
        torch.cuda.synchronize()
        e_param = rel(eng.params, ref.params)
        sig = torch.cat([eng.params, eng.exp_avg, eng.exp_avg_sq])
        gathered = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(gathered, sig)
        identical = all(torch.equal(gathered[0], t) for t in gathered)
        n = eng.n_arena
        param_diff = {k: float((gathered[0][a:b] - gathered[-1][a:b]).abs().max()) for k, (a, b) in regions.items()}
        half = eng.shape["L"] // 2 * eng.shape["r"] * eng.shape["D"]
        param_diff["w1_lower_half"] = float((gathered[0][:half] - gathered[-1][:half]).abs().max())
        param_diff["w1_upper_half"] = float((gathered[0][half:o["fcw"]] - gathered[-1][half:o["fcw"]]).abs().max())
        moved = rel(eng.params, make(B, solo[rank]).params)       # how far training moved the parameters (sanity)
        results[f"r{r_apla}_{'graph' if use_graph else 'eager'}"] = dict(
            grad_rel=e_grad, grad_cosine=cos, grad_max_diff_between_ranks_step1=grad_diff,
            param_max_diff_between_ranks=param_diff, params_rel_after_steps=e_param, identical_across_ranks=identical,
            params_moved_rel=moved)
        # (e_param: K Adam steps at lr 1e-3 amplify the 1e-7 gradient differences of the two batch shapes through
        #  m / sqrt(v) where gradients are near zero; 2e-4 .. 1e-3 measured at 2 and 8 ranks, against parameters that moved 0.25)
        ok &= identical and e_grad < 1e-2 and cos > 0.9999 and e_param < 5e-3 and moved > 1e-4
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps(dict(world=world, ok=bool(flag.item() == 1.0), cases=results)))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
