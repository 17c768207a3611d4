"""The native gradient all-reduce (`apla_grad_arena_allreduce`, csrc/dp_allreduce.cu) against NCCL on the gradient-arena
sizes of BASELINE configs C2 (2.0 MB), C3 (30 MB) and C4 (195 MB), isolated, CUDA events, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/bench_allreduce.py

Bus bandwidth = 2 (W-1)/W x bytes / time (the all-reduce convention); NVLink reference: 770 GB/s per direction per GPU
measured peer copy, 725 GB/s 8-rank NCCL bus bandwidth at 1 GiB (B200_PROFILING.md)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apla_b200.dp import PeerArena  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sizes = {"c2_2MB": 500_619, "c3_30MB": 7_513_899, "c4_195MB": 48_853_248}
    n_max = max(sizes.values())
    arena = PeerArena(n_max, f"cuda:{local}")
    nccl_buf = torch.zeros(n_max, device="cuda")
    out = []

    def timed(fn, iters):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) * 1e3

    for name, n in sizes.items():
        n4 = (n + 3) // 4 * 4
        # correctness first: every rank contributes rank + 1 -> sum = W (W + 1) / 2 everywhere
        arena.buf[:n4].fill_(float(rank + 1))
        torch.cuda.synchronize(); dist.barrier()
        arena.all_reduce(0, n4, 3, 32)
        torch.cuda.synchronize(); dist.barrier()
        ok = bool((arena.buf[:n4] == world * (world + 1) / 2).all())
        arena.buf.zero_()
        row = dict(slice=name, floats=n, bytes=n * 4, correct=ok)
        for ctas in (32, 64, 128):
            us = timed(lambda: arena.all_reduce(0, n4, 3, ctas), 20)
            row[f"native_{ctas}ctas_us"] = round(us, 1)
            row[f"native_{ctas}ctas_busbw_gbs"] = round(2 * (world - 1) / world * n * 4 / us / 1e3, 1)
        us = timed(lambda: dist.all_reduce(nccl_buf[:n]), 20)
        row["nccl_us"] = round(us, 1)
        row["nccl_busbw_gbs"] = round(2 * (world - 1) / world * n * 4 / us / 1e3, 1)
        out.append(row)
    if rank == 0:
        print(json.dumps(dict(world=world, multimem=bool(arena.multicast_ptr), results=out)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
