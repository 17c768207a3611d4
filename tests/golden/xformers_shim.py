"""A minimal stand-in for the parts of xformers 0.0.18 the reference's self-supervised path imports
(pinned in docker/conda-dinov2.yaml:17, absent from this image and from /root/reference).

USED ONLY BY THE GOLDEN-VECTOR GENERATORS in this directory, so that the UNMODIFIED reference code
(self_supervised/dinov2/{models,dinov2_vits}.py, layers/*, apla/appla_attn_mem_eff.py) can be executed on the CPU.
It restates the published semantics of three entry points in plain fp32 torch:

  xformers.ops.unbind(x, dim)                                   == torch.unbind
  xformers.ops.memory_efficient_attention(q, k, v, attn_bias)   q, k, v [B, M, H, K]; softmax(q k^T / sqrt(K)) v per
                                                                (batch, head); with a BlockDiagonalMask (B == 1) every
                                                                sequence of the packed batch attends to itself only
  xformers.ops.fmha.BlockDiagonalMask                           from_seqlens / from_tensor_list / split

Fixtures made through it pin everything of the reference EXCEPT the xformers kernels themselves; that residue is
stated wherever those fixtures are used ("pinned modulo the xformers shim").  `cross_entropy`, `SwiGLU` are left out
on purpose: the reference falls back to its own torch code when they are missing (ibot_patch_loss.py:11-27,
swiglu_ffn.py:37-49).  `scaled_index_add` / `index_select_cat` exist only so that layers/block.py's import succeeds;
they are reached with stochastic depth > 0 only, which no shipped config uses, and raise if called.
"""
import sys
import types

import torch


class BlockDiagonalMask:
    def __init__(self, seqlens):
        self.seqlens = [int(n) for n in seqlens]
        self._batch_sizes = None

    @classmethod
    def from_seqlens(cls, q_seqlen, kv_seqlen=None):
        assert kv_seqlen is None or list(kv_seqlen) == list(q_seqlen)
        return cls(q_seqlen)

    @classmethod
    def from_tensor_list(cls, tensors):
        """tensors [b_i, n_i, ...] -> (mask, [1, sum b_i n_i, ...])"""
        seqlens, sizes = [], []
        for t in tensors:
            sizes.append(t.shape[0])
            seqlens += [t.shape[1]] * t.shape[0]
        m = cls(seqlens)
        m._batch_sizes = sizes
        return m, torch.cat([t.reshape(1, -1, *t.shape[2:]) for t in tensors], dim=1)

    def split(self, x, batch_sizes=None):
        """[1, sum n, ...] -> list of [b_i, n_i, ...] (batch_sizes) or of [1, n_j, ...] per sequence."""
        assert x.shape[0] == 1 and x.shape[1] == sum(self.seqlens)
        sizes = batch_sizes or self._batch_sizes
        if sizes is None:
            return list(x.split(self.seqlens, dim=1))
        out, o, s = [], 0, 0
        for b in sizes:
            n = self.seqlens[s]
            assert all(v == n for v in self.seqlens[s:s + b])
            out.append(x[:, o:o + b * n].reshape(b, n, *x.shape[2:]))
            o += b * n
            s += b
        return out


def memory_efficient_attention(q, k, v, attn_bias=None, p=0.0, scale=None):
    assert p == 0.0
    scale = scale if scale is not None else q.shape[-1] ** -0.5

    def dense(q_, k_, v_):                                  # [B, M, H, K]
        a = torch.einsum("bmhk,bnhk->bhmn", q_, k_) * scale
        return torch.einsum("bhmn,bnhk->bmhk", a.softmax(dim=-1), v_)

    if attn_bias is None:
        return dense(q, k, v)
    assert isinstance(attn_bias, BlockDiagonalMask) and q.shape[0] == 1
    outs, o = [], 0
    for n in attn_bias.seqlens:
        outs.append(dense(q[:, o:o + n], k[:, o:o + n], v[:, o:o + n]))
        o += n
    assert o == q.shape[1]
    return torch.cat(outs, dim=1)


def _unsupported(*a, **k):
    raise NotImplementedError("xformers shim: stochastic-depth helpers are not restated (no shipped config uses them)")


def install():
    """Register the shim as `xformers` / `xformers.ops` / `xformers.ops.fmha` (no-op if a real xformers is importable)."""
    try:
        import xformers.ops  # noqa: F401
        return False
    except ImportError:
        pass
    x = types.ModuleType("xformers")
    ops = types.ModuleType("xformers.ops")
    fmha = types.ModuleType("xformers.ops.fmha")
    fmha.BlockDiagonalMask = BlockDiagonalMask
    ops.fmha = fmha
    ops.unbind = torch.unbind
    ops.memory_efficient_attention = memory_efficient_attention
    ops.scaled_index_add = _unsupported
    ops.index_select_cat = _unsupported
    x.ops = ops
    sys.modules["xformers"] = x
    sys.modules["xformers.ops"] = ops
    sys.modules["xformers.ops.fmha"] = fmha
    return True
