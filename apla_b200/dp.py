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

from dataclasses import dataclass
from typing import List, Tuple


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
