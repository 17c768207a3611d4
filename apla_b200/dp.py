"""Data-parallel host logic (device-agnostic, so it is testable with gloo on CPU): the layout of the trainable
gradient arena and the chunk plan of its all-reduce.

The reference wraps the model in DDP (src/defaults/wrappers.py:182-183): bucketed all-reduce(avg) of every trainable
gradient, plus a per-forward broadcast of the int64 `inds` buffers and a barrier per iteration.
Here all trainable gradients live in ONE contiguous fp32 arena

    [ proj_weight1 x L | fc.weight | proj_bias1 x L | fc.bias ]      (first `n_decay` elements are weight-decayed)

so the exchange is three slices of one tensor, issued on a side stream: the upper half of the blocks (+ fc.weight, which
is contiguous with them) as soon as backward has passed block L/2, the rest when backward ends.  The sum is turned into
DDP's mean by the 1/world factor applied inside the fused clip+AdamW kernel; indices are never re-broadcast (they are
immutable after construction and identical on all ranks by seed / inds_path), and there is no per-step barrier.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple


@dataclass(frozen=True)
class ArenaLayout:
    L: int
    r: int
    D: int
    C: int

    @property
    def w1(self) -> int:
        return 0

    @property
    def fcw(self) -> int:
        return self.L * self.r * self.D

    @property
    def b1(self) -> int:
        return self.fcw + self.C * self.D

    @property
    def fcb(self) -> int:
        return self.b1 + self.L * self.r

    @property
    def n(self) -> int:
        return self.fcb + self.C

    @property
    def n_decay(self) -> int:
        return self.b1

    def weight_slice(self, block: int) -> slice:
        return slice(self.w1 + block * self.r * self.D, self.w1 + (block + 1) * self.r * self.D)

    def bias_slice(self, block: int) -> slice:
        return slice(self.b1 + block * self.r, self.b1 + (block + 1) * self.r)

    def split_block(self) -> int:
        """Backward runs blocks L-1 .. split first; their gradients are reduced while blocks split-1 .. 0 run."""
        return self.L // 2

    def chunks(self) -> Tuple[List[slice], List[slice]]:
        """-> (early, late): slices reduced after the upper blocks' backward, and after the whole backward."""
        h = self.split_block()
        early = [slice(self.w1 + h * self.r * self.D, self.b1)]
        late = []
        if h > 0:
            late.append(slice(self.w1, self.w1 + h * self.r * self.D))
        late.append(slice(self.b1, self.n))
        return early, late


def allreduce_arena(grads, layout: ArenaLayout, group=None, which: str = "all") -> None:
    """Sum-all-reduce the arena in the engine's chunk order (`which` in {'early', 'late', 'all'})."""
    import torch.distributed as dist
    early, late = layout.chunks()
    todo = early + late if which == "all" else (early if which == "early" else late)
    for s in todo:
        dist.all_reduce(grads[s], group=group)


class PeerArena:
    """The gradient arena in SYMMETRIC memory + the native all-reduce over it (`apla_grad_arena_allreduce`,
    csrc/dp_allreduce.cu): every rank's arena and flag words are mapped into every process of the node
    (torch.distributed._symmetric_memory does the allocation and the rendezvous -- plumbing), and ONE kernel per slice
    reads / writes the peers' arenas through NVLink.  Unlike an NCCL call the kernel can be captured in the step's CUDA
    graph, so the data-parallel step is one graph with the early slice's reduction forked beside the lower backward.

    Raises if symmetric memory cannot be set up (the caller then keeps the NCCL path)."""

    def __init__(self, n_floats: int, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import LIB
        group = group if group is not None else dist.group.WORLD
        dll = LIB.load()
        self.n = int(n_floats)
        n_pad = (self.n + 1023) // 1024 * 1024
        self.buf = symm_mem.empty(n_pad, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(dll.apla_grad_arena_allreduce_flag_words(), dtype=torch.int32, device=device)
        self.buf.zero_()
        self.flags.zero_()
        hb = symm_mem.rendezvous(self.buf, group)
        hf = symm_mem.rendezvous(self.flags, group)
        self.rank, self.world = int(hb.rank), int(hb.world_size)
        if self.world > 8:
            raise RuntimeError("the native all-reduce covers one node (<= 8 GPUs)")
        self._buf_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hb.buffer_ptrs])
        self._flag_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hf.buffer_ptrs])
        self._handles = (hb, hf)                                     # keep the mappings alive
        # NVLS multicast mapping of the arena (0 when the fabric has none): the switch then does the adding
        # (measured at 2 GPUs the plain peer loop is faster -- 73 vs 102 us for 30 MB -- so NVLS is used from 4 ranks up;
        #  APLA_DP_MULTIMEM=0 / 1 forces it off / on)
        mc = int(getattr(hb, "multicast_ptr", 0) or 0)
        want = os.environ.get("APLA_DP_MULTIMEM", "auto")
        if want == "0" or (want == "auto" and self.world < 4):
            mc = 0
        self.multicast_ptr = mc
        self.epochs = torch.zeros(dll.apla_grad_arena_allreduce_epoch_words(), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                          # every rank's flags are zero before the first launch

    def grads(self):
        return self.buf[:self.n]

    def all_reduce(self, offset: int, count: int, channel: int, ctas: int = 32, offset_b: int = 0, count_b: int = 0) -> None:
        """Sum arena[offset, offset + count) (and a second slice [offset_b, offset_b + count_b)) over the ranks, in
        place, as one launch on the current CUDA stream (capturable)."""
        from ._lib import LIB, ptr, stream
        count = (count + 3) // 4 * 4                                 # the padding past n is zero on every rank
        count_b = (count_b + 3) // 4 * 4
        for o, c in ((offset, count), (offset_b, count_b)):
            if o % 4 or o + c > self.buf.numel():
                raise RuntimeError(f"slice [{o}, {o + c}) is not 16-byte aligned inside the arena")
        LIB.call("apla_grad_arena_allreduce", self._buf_ptrs, self._flag_ptrs, self.multicast_ptr or None, ptr(self.epochs),
                 self.rank, self.world, offset, count, offset_b, count_b, channel, ctas, stream())

    def all_reduce_chunks(self, layout: ArenaLayout, which: str, ctas: int = 16) -> None:
        """The engine's chunk plan (`ArenaLayout.chunks`): 'early' on channel 0; the two 'late' slices share ONE launch on
        channel 1 (one pair of cross-GPU barriers)."""
        early, late = layout.chunks()
        if which == "early":
            for s in early:
                self.all_reduce(s.start, s.stop - s.start, 0, ctas)
        elif len(late) == 2:
            self.all_reduce(late[0].start, late[0].stop - late[0].start, 1, ctas, late[1].start, late[1].stop - late[1].start)
        else:
            for ch, s in enumerate(late):
                self.all_reduce(s.start, s.stop - s.start, 1 + ch, ctas)


def make_peer_arena(n_floats: int, device, group=None) -> Optional[PeerArena]:
    """PeerArena, or None when APLA_DP_ALLREDUCE=nccl or symmetric memory is not available on this node."""
    if os.environ.get("APLA_DP_ALLREDUCE", "native").lower() == "nccl":
        return None
    try:
        return PeerArena(n_floats, device, group)
    except Exception as e:                                           # noqa: BLE001 -- fall back to the NCCL reducer
        import warnings
        warnings.warn(f"apla_b200: native all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
        return None
