"""`APLA_MemEffAttention`: drop-in for src/apla/appla_attn_mem_eff.py:22-67.

The reference routes attention through xformers' `memory_efficient_attention` and accepts a `BlockDiagonalMask`
so that crops of different length, concatenated into one [1, sum(N), C] sequence by dinov2's NestedTensorBlock
(src/self_supervised/dinov2/layers/block.py:191-217), attend only within themselves.  Here the same fused kernel
as `APLA_Attention` is used, with the mask expressed as packed sequence offsets (cu_seqlens).  Returns the tensor
only, like the reference (:67).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .appla_attn import APLA_Attention


class BlockDiagonalMask:
    """Minimal stand-in for xformers.ops.fmha.attn_bias.BlockDiagonalMask: independent attention per packed
    sequence.  `from_seqlens([257, 257, 50, ...])` mirrors the xformers constructor used at block.py:202."""

    def __init__(self, seqlens: Sequence[int]):
        self.seqlens: List[int] = [int(s) for s in seqlens]
        if not self.seqlens or min(self.seqlens) <= 0:
            raise ValueError("BlockDiagonalMask needs at least one positive sequence length")
        self._cu = {}

    @classmethod
    def from_seqlens(cls, seqlens: Sequence[int]) -> "BlockDiagonalMask":
        return cls(seqlens)

    def cu_seqlens(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._cu:
            off = [0]
            for s in self.seqlens:
                off.append(off[-1] + s)
            self._cu[key] = torch.tensor(off, dtype=torch.int32, device=device)
        return self._cu[key]


def _seqlens_of(attn_bias) -> List[int]:
    """Accept our BlockDiagonalMask or an xformers one (duck-typed through q_seqinfo.seqstart_py)."""
    if isinstance(attn_bias, BlockDiagonalMask):
        return attn_bias.seqlens
    info = getattr(attn_bias, "q_seqinfo", None)
    starts = getattr(info, "seqstart_py", None)
    if starts is not None:
        return [int(b - a) for a, b in zip(starts[:-1], starts[1:])]
    raise AssertionError("attn_bias must be a block-diagonal mask (packed variable-length sequences)")


class APLA_MemEffAttention(APLA_Attention):
    def forward(self, x: torch.Tensor, attn_bias=None) -> torch.Tensor:
        if attn_bias is None:
            return self._run(x)
        seqlens = _seqlens_of(attn_bias)
        if x.dim() != 3 or x.shape[0] != 1 or x.shape[1] != sum(seqlens):
            raise AssertionError(f"packed input must be [1, {sum(seqlens)}, C], got {tuple(x.shape)}")
        if isinstance(attn_bias, BlockDiagonalMask):
            cu = attn_bias.cu_seqlens(x.device)
        else:
            cu = BlockDiagonalMask(seqlens).cu_seqlens(x.device)
        return self._run(x, cu_seqlens=cu, seqlens=seqlens)
