"""`FineTuneEngine`: host side of the native step engine (csrc/engine.cu).

Takes a classifier built the reference way (`Classifier`: backbone ViT with APLA attention + `fc` head,
src/defaults/models.py:24-92; here usually `hostvit.HostClassifier`), lays its tensors out for the B200 and runs
`Trainer.global_step` (src/defaults/trainer.py:106-138) -- forward, CrossEntropy, backward restricted to the APLA
rows and the head, data-parallel gradient mean, clip_grad_norm_(1.0), AdamW -- as native kernel launches.

PyTorch's role here is plumbing only: device allocation (torch tensors own every buffer), the one-time conversion
of frozen fp32 weights into bf16 [out,in] / [in,out] copies, stream handles and torch.distributed/NCCL for the
gradient all-reduce.  All step arithmetic is in libapla_b200.so; nothing falls back to torch ops.

Memory layout in HBM (T = B*N tokens, D embed, L blocks), sized for B=64 ViT-B/14 (~5 GB of 180 GB):
  xs        fp32 [2L+1, T, D]   residual-stream checkpoints (block input / after attention / ... / final)
  qkv[l]    bf16 [T, 3D]  ao[l] bf16 [T, D]  lse[l] fp32 [T, H]  hpre[l] fp16 [T, 4D] (gelu' of the fc1 pre-activation)   saved for backward
  ln_out, gelu_out, dx, dxb, dO, dqkv, dsub, delta                                         transients, reused
  params / grads / exp_avg / exp_avg_sq   fp32 arenas [W1 x L | fc.weight | b1 x L | fc.bias]  (one all-reduce)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import os

import torch
import torch.nn as nn

from ._lib import LIB, ptr, require_device, stream

BF16, F32 = torch.bfloat16, torch.float32


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def interpolate_pos_table(pos_embed: torch.Tensor, npatch: int) -> torch.Tensor:
    """Bicubic resize of a [1, 1+n0, D] position table to an npatch grid (vit.py:421-437); done ONCE at engine
    construction because pos_embed is frozen (the reference recomputes it every forward)."""
    n0 = pos_embed.shape[1] - 1
    if npatch == n0:
        return pos_embed[0]
    D = pos_embed.shape[-1]
    s = int(math.sqrt(n0))
    grid = pos_embed[:, 1:].reshape(1, s, s, D).permute(0, 3, 1, 2)
    grid = nn.functional.interpolate(grid, scale_factor=math.sqrt(npatch / n0), mode="bicubic", align_corners=False,
                                     recompute_scale_factor=False)
    return torch.cat((pos_embed[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, D)), dim=1)[0]


class FineTuneEngine:
    def __init__(self, model: nn.Module, batch_size: int, img_size: int, device="cuda", lr: float = 3e-5,
                 weight_decay: float = 1e-5, clip: float = 1.0, betas=(0.9, 0.999), adam_eps: float = 1e-8,
                 process_group=None, cls_only_last_block: bool = True, use_graph: bool = True):
        require_device()
        # Only norm(x)[:, 0] of the last block reaches the head (vit.py:417-419): its projection, LayerNorm 2, MLP and
        # their input gradients run on the CLS rows only.  False = every token (identical results, for A/B checks).
        self.cls_only_last_block = 2 if cls_only_last_block is True else int(cls_only_last_block)   # 1: tail only
        self.device = torch.device(device)
        self.model = model
        self.lr, self.wd, self.clip, self.betas, self.adam_eps = lr, weight_decay, clip, betas, adam_eps
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if self._dist_on() else 1
        self.step_count = 0
        # One CUDA graph per (images, labels) buffer pair replays the step's ~200 launches (single GPU only: with a
        # process group the all-reduce runs on a side stream between two backward ranges).  lr and Adam's bias
        # corrections reach the captured AdamW kernel through a 3-float device buffer refreshed before every replay.
        # With a process group the step is three graphs (forward + upper backward | lower backward | optimiser) with the
        # two gradient all-reduces launched eagerly between them on the side stream: capturing the NCCL collectives
        # themselves hung on this stack (torch 2.11 / NCCL 2.28), and they are two launches anyway.
        self.use_graph = bool(use_graph)
        self._graphs: Dict[tuple, object] = {}
        self._graph_seen: Dict[tuple, int] = {}
        self._keep: List[torch.Tensor] = []          # every device buffer the engine points at
        self._handle = None
        self._build(model, batch_size, img_size)

    # ------------------------------------------------------------------------------------------------------------
    def _dist_on(self) -> bool:
        return torch.distributed.is_available() and torch.distributed.is_initialized()

    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        out = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self._keep.append(out)
        return out

    def _new(self, *shape, dtype=BF16, zero=False) -> torch.Tensor:
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.device)
        self._keep.append(t)
        return t

    def _set(self, name: str, t: Optional[torch.Tensor], block: int = -1):
        LIB.call("apla_engine_set_ptr", self._handle, name.encode(), block, ptr(t))

    def _build(self, model, B, img):
        bb = model.backbone
        conv = bb.patch_embed.proj
        patch = conv.kernel_size[0]
        D = bb.cls_token.shape[-1]
        L = len(bb.blocks)
        P = (img // patch) ** 2
        N = P + 1
        T = B * N
        blk0 = bb.blocks[0]
        H = blk0.attn.num_heads
        hidden = blk0.mlp.fc1.out_features
        C = model.fc.out_features
        kpad = (3 * patch * patch + 63) // 64 * 64
        self.stock_proj = hasattr(blk0.attn, "proj")            # multi-GPU 'full': stock attention, proj trainable
        r = D if self.stock_proj else int(blk0.attn.partial_size)
        full_rows = 1 if r > 128 else 0
        r_pad = _pad64(r)
        self.shape = dict(B=B, N=N, D=D, H=H, L=L, hidden=hidden, C=C, P=P, patch=patch, img=img, kpad=kpad, r=r,
                          r_pad=r_pad, full_rows=full_rows, T=T)
        eps = float(blk0.norm1.eps)
        scale = float(blk0.attn.scale)
        self._handle = LIB.load().apla_engine_create(B, N, D, H, L, hidden, C, patch, img, kpad, r, r_pad, full_rows,
                                                     eps, scale)
        if not self._handle:
            raise RuntimeError("apla_engine_create failed: " + LIB.last_error())
        LIB.call("apla_engine_set_option", self._handle, b"cls_only_last_block", int(self.cls_only_last_block))
        # measured: no gain on C2 (7.94 vs 7.98 ms), +0.4 % on C3 -- the step runs into the power cap, not into idle SMs;
        # kept as an option (APLA_SIDE_WGRAD=1), off by default
        self.side_wgrad = os.environ.get("APLA_SIDE_WGRAD", "0") != "0"

        # ---- embedding ----
        wpe = torch.zeros(D, kpad, dtype=F32)
        wpe[:, :3 * patch * patch] = conv.weight.detach().float().reshape(D, -1).cpu()
        self._set("wpe", self._dev(wpe, BF16))
        self._set("bpe", self._dev(conv.bias if conv.bias is not None else torch.zeros(D), F32))
        self._set("cls", self._dev(bb.cls_token.reshape(D), F32))
        self._set("pos", self._dev(interpolate_pos_table(bb.pos_embed.detach().float().cpu(), P), F32))
        self._set("patches", self._new(B * P, kpad))
        self._set("pe_out", self._new(B * P, D))

        # ---- trainable arena ----
        n = LIB.load().apla_engine_arena_size(self._handle)
        self.n_arena = int(n)
        self.params = self._new(n, dtype=F32, zero=True)
        # data parallel: the gradient arena lives in symmetric (peer-mapped) memory and is reduced by the library's own
        # kernel (apla_grad_arena_allreduce), which -- unlike an NCCL call -- is captured in the step's CUDA graph
        self._peer = None
        if self.world > 1:
            from .dp import make_peer_arena
            self._peer = make_peer_arena(n, self.device, self.pg)
            # every rank must take the same path: if symmetric memory failed anywhere, all ranks use NCCL
            ok = torch.tensor([1 if self._peer is not None else 0], device=self.device)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=self.pg)
            if int(ok.item()) == 0:
                self._peer = None
        if self._peer is not None:
            self.grads = self._peer.grads()
            self._keep.append(self._peer.buf)
        else:
            self.grads = self._new(n, dtype=F32, zero=True)
        self.exp_avg = self._new(n, dtype=F32, zero=True)
        self.exp_avg_sq = self._new(n, dtype=F32, zero=True)
        self.sumsq = self._new(600, dtype=F32, zero=True)      # APLA_SUMSQ_FLOATS: [0] = result, rest scratch
        o_w1, o_fcw = 0, L * r * D
        o_b1 = o_fcw + C * D
        o_fcb = o_b1 + L * r
        self.offsets = dict(w1=o_w1, fcw=o_fcw, b1=o_b1, fcb=o_fcb, n=o_fcb + C)
        assert self.offsets["n"] == n
        self.slices: Dict[str, slice] = {}     # parameter name -> slice of the arena
        self.shapes: Dict[str, tuple] = {}

        idx_all = torch.zeros(L, r, dtype=torch.int32)
        rowmap_all = torch.full((L, D), -1, dtype=torch.int32)
        # dense projection copies of all blocks back to back: refreshed by ONE kernel after every optimiser step
        self._wproj_all = self._new(L, D, D)
        self._wprojT_all = self._new(L, D, D)
        self._bproj_all = self._new(L, D, dtype=F32)
        # ---- blocks ----
        for l, blk in enumerate(bb.blocks):
            at = blk.attn
            pre = f"backbone.blocks.{l}."
            if self.stock_proj:
                inds = torch.arange(D)
                w_full = at.proj.weight.detach().float().cpu()
                b_full = at.proj.bias.detach().float().cpu()
                w1, b1 = w_full, b_full
                names = (pre + "attn.proj.weight", pre + "attn.proj.bias")
            else:
                inds = torch.as_tensor(at.indices).long().cpu()
                w1 = at.proj_weight1.detach().float().cpu()
                b1 = at.proj_bias1.detach().float().cpu()
                w_full = torch.zeros(D, D)
                b_full = torch.zeros(D)
                w_full[inds[:r]] = w1
                b_full[inds[:r]] = b1
                if r < D:
                    w_full[inds[r:]] = at.proj_weight2.detach().float().cpu()
                    b_full[inds[r:]] = at.proj_bias2.detach().float().cpu()
                names = (pre + "attn.proj_weight1", pre + "attn.proj_bias1")
            idx_all[l] = inds[:r].to(torch.int32)
            rowmap_all[l, inds[:r]] = torch.arange(r, dtype=torch.int32)
            self.slices[names[0]] = slice(o_w1 + l * r * D, o_w1 + (l + 1) * r * D)
            self.shapes[names[0]] = (r, D)
            self.slices[names[1]] = slice(o_b1 + l * r, o_b1 + (l + 1) * r)
            self.shapes[names[1]] = (r,)
            self.params[self.slices[names[0]]] = w1.reshape(-1).to(self.device)
            self.params[self.slices[names[1]]] = b1.to(self.device)

            def lin(mod):
                w = mod.weight.detach().float()
                b = mod.bias.detach().float() if mod.bias is not None else torch.zeros(w.shape[0])
                return self._dev(w, BF16), self._dev(w.t(), BF16), self._dev(b, F32)

            self._wproj_all[l].copy_(w_full.to(self.device))
            self._wprojT_all[l].copy_(w_full.t().to(self.device))
            self._bproj_all[l].copy_(b_full.to(self.device))
            wqkv, wqkvT, bqkv = lin(at.qkv)
            wfc1, wfc1T, bfc1 = lin(blk.mlp.fc1)
            wfc2, wfc2T, bfc2 = lin(blk.mlp.fc2)
            for k, v in dict(wqkv=wqkv, wqkvT=wqkvT, bqkv=bqkv, wfc1=wfc1, wfc1T=wfc1T, bfc1=bfc1, wfc2=wfc2,
                             wfc2T=wfc2T, bfc2=bfc2, wproj=self._wproj_all[l], wprojT=self._wprojT_all[l],
                             bproj=self._bproj_all[l], ln1w=self._dev(blk.norm1.weight, F32),
                             ln1b=self._dev(blk.norm1.bias, F32), ln2w=self._dev(blk.norm2.weight, F32),
                             ln2b=self._dev(blk.norm2.bias, F32)).items():
                self._set(k, v, l)
            g1 = getattr(blk.ls1, "gamma", None)
            g2 = getattr(blk.ls2, "gamma", None)
            self._set("g1", self._dev(g1, F32) if g1 is not None else None, l)
            self._set("g2", self._dev(g2, F32) if g2 is not None else None, l)
            self._set("qkv", self._new(T, 3 * D), l)
            self._set("ao", self._new(T, D, zero=True), l)   # last block: only its CLS rows are ever written
            self._set("hpre", self._new(T, hidden, dtype=torch.float16), l)
            self._set("lse", self._new(T, H, dtype=F32), l)

        # ---- head ----
        self.slices["fc.weight"] = slice(o_fcw, o_fcw + C * D)
        self.shapes["fc.weight"] = (C, D)
        self.slices["fc.bias"] = slice(o_fcb, o_fcb + C)
        self.shapes["fc.bias"] = (C,)
        self.params[self.slices["fc.weight"]] = model.fc.weight.detach().float().reshape(-1).to(self.device)
        self.params[self.slices["fc.bias"]] = model.fc.bias.detach().float().to(self.device)
        self._set("lnfw", self._dev(bb.norm.weight, F32))
        self._set("lnfb", self._dev(bb.norm.bias, F32))

        # ---- globals ----
        self.xs = self._new(2 * L + 1, T, D, dtype=F32)
        self.logits = self._new(B, C, dtype=F32)
        self.loss = self._new(1, dtype=F32, zero=True)
        self.idx = self._dev(idx_all, torch.int32)
        for name, t in dict(xs=self.xs, ln_out=self._new(T, D), gelu_out=self._new(T, hidden),
                            dx=self._new(T, D, dtype=F32), dxb=self._new(T, D), dO=self._new(T, D),
                            dqkv=self._new(T, 3 * D), delta=self._new(T, H, dtype=F32),
                            cls_ln=self._new(B, D), logits=self.logits, dlogits=self._new(B, C, dtype=F32),
                            loss=self.loss, dcls=self._new(B, D), params=self.params, grads=self.grads,
                            exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, idx=self.idx, sumsq=self.sumsq).items():
            self._set(name, t)
        if full_rows:
            self._set("rowmap", self._dev(rowmap_all, torch.int32))
        else:
            self._set("dsub", self._new(2 if self.side_wgrad else 1, T, r_pad))
        # weight / bias gradients of a block on the engine's side stream beside the projection dgrad, the attention
        # backward and the qkv dgrad (csrc/engine.cu, Engine::side_wgrad) when APLA_SIDE_WGRAD=1
        LIB.call("apla_engine_set_option", self._handle, b"side_wgrad", (1 if full_rows else 2) if self.side_wgrad else 0)
        self._hyper_dev = self._new(3, dtype=F32, zero=True)
        if self.use_graph:
            self._set("hyper", self._hyper_dev)
        self.images_dev = self._new(B, 3, img, img, dtype=F32)
        self.labels_dev = self._new(B, dtype=torch.int64)
        self._ar_stream = torch.cuda.Stream(device=self.device) if self.world > 1 else None
        torch.cuda.synchronize(self.device)

    def __del__(self):
        try:
            if self._handle:
                LIB.load().apla_engine_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------
    def trainable_names(self) -> List[str]:
        """Same order as the reference's named_parameters() filter (I5/I7 of SURVEY.md 4.2)."""
        L = self.shape["L"]
        w, b = ("attn.proj.weight", "attn.proj.bias") if self.stock_proj else ("attn.proj_weight1", "attn.proj_bias1")
        out = []
        for l in range(L):
            out += [f"backbone.blocks.{l}.{w}", f"backbone.blocks.{l}.{b}"]
        return out + ["fc.weight", "fc.bias"]

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return {k: self.grads[self.slices[k]].view(self.shapes[k]) for k in self.trainable_names()}

    def named_params(self) -> Dict[str, torch.Tensor]:
        return {k: self.params[self.slices[k]].view(self.shapes[k]) for k in self.trainable_names()}

    def sync_to_model(self):
        """Write the arena's fp32 parameters back into the nn.Module (state_dict / checkpoint compatibility)."""
        sd = dict(self.model.named_parameters())
        with torch.no_grad():
            for k, v in self.named_params().items():
                sd[k].copy_(v.to(sd[k].device))

    # ---- checkpoint interchange with the reference (apla_b200/checkpoint.py) ---------------
    def _named_shapes(self):
        return [(k, self.shapes[k]) for k in self.trainable_names()]

    def state_dict(self):
        """The Classifier's CPU state dict with the engine's current trainable tensors written back
        (what bases.py:455-458 saves under 'state_dict')."""
        from .checkpoint import model_to_cpu_state
        torch.cuda.current_stream().synchronize()
        self.sync_to_model()
        return model_to_cpu_state(self.model)

    def optimizer_state_dict(self) -> dict:
        """`torch.optim.AdamW.state_dict()` of the reference's optimiser (two groups of wrappers.py:205-221) holding
        the engine's moments and step count."""
        from .checkpoint import optimizer_state_dict
        torch.cuda.current_stream().synchronize()
        m = {k: self.exp_avg[self.slices[k]] for k in self.trainable_names()}
        v = {k: self.exp_avg_sq[self.slices[k]] for k in self.trainable_names()}
        return optimizer_state_dict(self._named_shapes(), m, v, self.step_count, lr=self.lr, betas=self.betas,
                                    eps=self.adam_eps, weight_decay=self.wd)

    def load_state_dict(self, sd):
        """Take the TRAINABLE tensors of a reference state dict into the arena and refresh the dense projection copies.
        Frozen tensors are laid out (bf16, both orientations) when the engine is built, so they are only checked: a
        state dict whose APLA indices differ from the ones this engine was built with is refused."""
        bb = self.model.backbone
        for l, blk in enumerate(bb.blocks):
            k = f"backbone.blocks.{l}.attn.inds"
            if k in sd and hasattr(blk.attn, "inds") and not torch.equal(sd[k].cpu().long(), blk.attn.inds.cpu().long()):
                raise RuntimeError(f"{k} differs from the indices this engine was built with; load the checkpoint into "
                                   "the model (load_from_pretrained) before constructing the engine")
        missing = [k for k in self.trainable_names() if k not in sd]
        if missing:
            raise KeyError(f"state dict lacks trainable tensors: {missing[:4]}{' ...' if len(missing) > 4 else ''}")
        for k in self.trainable_names():
            t = sd[k]
            if tuple(t.shape) != tuple(self.shapes[k]):
                raise RuntimeError(f"{k}: checkpoint shape {tuple(t.shape)}, engine {tuple(self.shapes[k])}")
            self.params[self.slices[k]] = t.detach().to(self.device, F32).reshape(-1)
        self._refresh_proj()
        self.sync_to_model()

    def _refresh_proj(self):
        s = self.shape
        o = self.offsets
        LIB.call("apla_proj_refresh", self.params[o["w1"]:].data_ptr(), self.params[o["b1"]:].data_ptr(), ptr(self.idx),
                 ptr(self._wproj_all), ptr(self._wprojT_all), ptr(self._bproj_all), s["L"], s["r"], s["D"],
                 s["r"] * s["D"], s["r"], stream())

    def load_optimizer_state_dict(self, opt_sd: dict):
        """Adam moments, step count and hyper-parameters from a reference optimiser state dict (bases.py:423-428)."""
        from .checkpoint import split_optimizer_state
        m, v, step, hyper = split_optimizer_state(opt_sd, self._named_shapes())
        for k in self.trainable_names():
            self.exp_avg[self.slices[k]] = m[k].to(self.device).reshape(-1)
            self.exp_avg_sq[self.slices[k]] = v[k].to(self.device).reshape(-1)
        self.step_count = step
        changed = (hyper["betas"] != tuple(self.betas) or hyper["eps"] != self.adam_eps
                   or hyper["weight_decay"] != self.wd)
        self.lr, self.betas, self.adam_eps, self.wd = hyper["lr"], hyper["betas"], hyper["eps"], hyper["weight_decay"]
        if changed:
            self.reset_graphs()         # captured graphs bake everything except lr and the step count

    def save_session(self, path: str, *, iters: Optional[int] = None, epoch: int = 0, parameters=None,
                     best_val_target=None, original_state=None) -> str:
        """Write `<path>` in the reference's session format (bases.py:448-468).  Data parallel: rank 0 writes (every
        rank holds the same parameters and moments), then all ranks meet, as the reference does (`if self.is_rank0`
        ... `synchronize()`, bases.py:453,467)."""
        from .checkpoint import save_session
        dist_on = self._dist_on() and self.world > 1
        if not dist_on or torch.distributed.get_rank(self.pg) == 0:
            save_session(path, state_dict=self.state_dict(), optimizer=self.optimizer_state_dict(),
                         iters=self.step_count if iters is None else iters, epoch=epoch, parameters=parameters,
                         best_val_target=best_val_target, original_state=original_state)
        if dist_on:
            torch.distributed.barrier(self.pg)
        return path

    def load_session(self, path: str, restore_only_model: bool = False) -> dict:
        """bases.py:405-433: model tensors, then (unless restore_only_model) iters / epoch / optimiser state.
        Returns the checkpoint dict (iters, epoch, parameters, ... for the caller's bookkeeping)."""
        from .checkpoint import load_session_file
        ckpt = load_session_file(path)
        self.load_state_dict(ckpt["state_dict"])
        if not restore_only_model:
            self.load_optimizer_state_dict(ckpt["optimizer"])
        torch.cuda.current_stream().synchronize()
        return ckpt

    # ------------------------------------------------------------------------------------------------------------
    def _check_inputs(self, images, labels):
        B = self.shape["B"]
        if images.shape != self.images_dev.shape or images.dtype != F32 or not images.is_cuda:
            raise RuntimeError(f"images must be a CUDA fp32 tensor of shape {tuple(self.images_dev.shape)}")
        if labels is not None and (labels.dtype != torch.int64 or labels.shape != (B,) or not labels.is_cuda):
            raise RuntimeError("labels must be a CUDA int64 tensor of shape [B]")

    def forward(self, images: torch.Tensor, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """images fp32 [B,3,S,S] on the engine's device -> logits fp32 [B,C] (and loss/dlogits when labels given)."""
        B = self.shape["B"]
        self._check_inputs(images, labels)
        images = images.contiguous()
        # CrossEntropyLoss(mean) over the local batch (wrappers.py:314); DDP later averages gradients over ranks
        LIB.call("apla_engine_forward", self._handle, ptr(images), ptr(labels), 1.0 / B, 1.0 / B, stream())
        return self.logits

    def backward(self):
        L = self.shape["L"]
        if self.world == 1:
            LIB.call("apla_engine_backward", self._handle, L - 1, 0, stream())
            return
        # data parallel: reduce the upper blocks' gradients (plus fc.weight, contiguous with them in the arena) on a
        # side stream while the lower blocks are still in backward (chunk plan: apla_b200/dp.py)
        from .dp import ArenaLayout, allreduce_arena
        lay = ArenaLayout(L=L, r=self.shape["r"], D=self.shape["D"], C=self.shape["C"])
        half = lay.split_block()
        cur = torch.cuda.current_stream()

        def reduce(which):
            if self._peer is not None:
                # few CTAs while the lower backward shares the GPU, more for the exposed tail
                self._peer.all_reduce_chunks(lay, which, ctas=32 if which == "early" else 96)
            else:
                allreduce_arena(self.grads, lay, group=self.pg, which=which)

        LIB.call("apla_engine_backward", self._handle, L - 1, half, stream())
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(self._ar_stream):
            self._ar_stream.wait_event(ev)
            reduce("early")
        if half > 0:
            LIB.call("apla_engine_backward", self._handle, half - 1, 0, stream())
        ev2 = torch.cuda.Event()
        ev2.record(cur)
        with torch.cuda.stream(self._ar_stream):
            self._ar_stream.wait_event(ev2)
            reduce("late")
        cur.wait_stream(self._ar_stream)

    def reset_graphs(self):
        """Drop the captured CUDA graphs (they bake weight_decay, clip, betas and adam_eps -- everything except lr and the
        step count, which travel through the device-side hyper buffer); the next steps re-capture."""
        self._graphs.clear()
        self._graph_seen.clear()
        self._graph_keep = []

    def _push_hyper(self):
        """lr and the bias corrections of the step about to run -> device (stream-ordered, before the AdamW kernel)."""
        t = self.step_count
        # a fresh PAGEABLE tensor: the runtime stages its bytes at call time, so the host may prepare the next step's
        # values while this copy is still queued (a reused pinned buffer would be read when the copy executes)
        self._hyper_dev.copy_(torch.tensor([self.lr, 1.0 - self.betas[0] ** t, (1.0 - self.betas[1] ** t) ** 0.5],
                                           dtype=F32), non_blocking=True)

    def optim_step(self):
        self.step_count += 1
        if self.use_graph:
            self._push_hyper()
        LIB.call("apla_engine_optim", self._handle, 1.0 / self.world, float(self.clip or 0.0), self.lr, self.wd,
                 self.betas[0], self.betas[1], self.adam_eps, self.step_count, stream())

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """One Trainer.global_step on device-resident inputs; returns the (device) loss scalar."""
        if not self.use_graph:
            self.forward(images, labels)
            self.backward()
            self.optim_step()
            return self.loss
        # graph path: a buffer pair is run eagerly the first time it is seen, captured on its second step (same launch
        # sequence, nothing executes during capture) and replayed from then on
        if not images.is_contiguous():
            images = images.contiguous()      # (a fresh buffer every call: stays on the eager path)
        key = (images.data_ptr(), labels.data_ptr(), tuple(images.shape))
        g = self._graphs.get(key)
        if g is None:
            seen = self._graph_seen.get(key, 0)
            self._graph_seen[key] = seen + 1
            if seen == 0 or len(self._graphs) >= 4:
                self.forward(images, labels)
                self.backward()
                self.optim_step()
                return self.loss
            self._check_inputs(images, labels)
            if not images.is_contiguous():
                raise RuntimeError("images must be contiguous")
            torch.cuda.current_stream().synchronize()
            L = self.shape["L"]
            inv_b = 1.0 / self.shape["B"]

            def capture(fn):
                gr = torch.cuda.CUDAGraph()
                # captured on a high-priority stream: the kernel nodes inherit it, so the main chain's CTAs are placed
                # before those of the lowest-priority side branch (weight gradients) whenever both are pending
                if getattr(self, "_cap_stream", None) is None:
                    self._cap_stream = torch.cuda.Stream(device=self.device, priority=-1)
                with torch.cuda.graph(gr, stream=self._cap_stream, capture_error_mode="thread_local"):
                    fn()
                return gr

            def optim():
                LIB.call("apla_engine_optim", self._handle, 1.0 / self.world, float(self.clip or 0.0), self.lr, self.wd,
                         self.betas[0], self.betas[1], self.adam_eps, 1, stream())

            if self.world == 1:
                def whole():
                    LIB.call("apla_engine_forward", self._handle, ptr(images), ptr(labels), inv_b, inv_b, stream())
                    LIB.call("apla_engine_backward", self._handle, L - 1, 0, stream())
                    optim()
                g = (capture(whole),)
            elif self._peer is not None:
                # data parallel with the native all-reduce: ONE graph; the early slice's reduction is forked onto the
                # side stream inside the capture and joins before the optimiser
                def whole_dp():
                    LIB.call("apla_engine_forward", self._handle, ptr(images), ptr(labels), inv_b, inv_b, stream())
                    self.backward()
                    optim()
                g = (capture(whole_dp),)
            else:
                from .dp import ArenaLayout
                half = ArenaLayout(L=L, r=self.shape["r"], D=self.shape["D"], C=self.shape["C"]).split_block()

                def upper():
                    LIB.call("apla_engine_forward", self._handle, ptr(images), ptr(labels), inv_b, inv_b, stream())
                    LIB.call("apla_engine_backward", self._handle, L - 1, half, stream())

                def lower():
                    LIB.call("apla_engine_backward", self._handle, half - 1, 0, stream())
                g = (capture(upper), capture(lower) if half > 0 else None, capture(optim))
            self._graphs[key] = g
            self._graph_keep = getattr(self, "_graph_keep", []) + [(images, labels)]   # keep the buffers alive
        self.step_count += 1
        self._push_hyper()
        if self.world == 1 or self._peer is not None:
            g[0].replay()
            return self.loss
        from .dp import ArenaLayout, allreduce_arena
        lay = ArenaLayout(L=self.shape["L"], r=self.shape["r"], D=self.shape["D"], C=self.shape["C"])
        cur = torch.cuda.current_stream()
        g[0].replay()
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(self._ar_stream):
            self._ar_stream.wait_event(ev)
            allreduce_arena(self.grads, lay, group=self.pg, which="early")
        if g[1] is not None:
            g[1].replay()
        ev2 = torch.cuda.Event()
        ev2.record(cur)
        with torch.cuda.stream(self._ar_stream):
            self._ar_stream.wait_event(ev2)
            allreduce_arena(self.grads, lay, group=self.pg, which="late")
        cur.wait_stream(self._ar_stream)
        g[2].replay()
        return self.loss

    def step_from_host(self, images_pinned: torch.Tensor, labels_pinned: torch.Tensor, lag: bool = True):
        """The end-to-end call: pinned host batch -> H2D -> step -> loss read back to the host.

        The batch is copied on a side stream into one of two device slots (so the copy of step k overlaps the compute
        of step k-1) and the loss comes back through a pinned buffer.  With `lag=True` (default) the call returns the
        loss of the PREVIOUS step (None on the first call) so that the host never waits for the step it just enqueued
        -- the same one-step-late logging the reference does every `log_every` (src/defaults/trainer.py:145-151);
        `lag=False` synchronises and returns this step's loss.  `drain()` returns the last pending loss."""
        if not hasattr(self, "_e2e"):
            B, img = self.shape["B"], self.shape["img"]
            self._e2e = dict(
                copy_stream=torch.cuda.Stream(device=self.device),
                img=[self.images_dev, self._new(B, 3, img, img, dtype=F32)],
                lab=[self.labels_dev, self._new(B, dtype=torch.int64)],
                copied=[torch.cuda.Event(), torch.cuda.Event()],
                done=[torch.cuda.Event(), torch.cuda.Event()],
                loss_host=[torch.zeros(1, dtype=F32).pin_memory(), torch.zeros(1, dtype=F32).pin_memory()],
                pending=[False, False], slot=0)
        st = self._e2e
        slot = st["slot"]
        st["slot"] = slot ^ 1
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(st["copy_stream"]):
            if st["pending"][slot]:
                st["copy_stream"].wait_event(st["done"][slot])      # the step that last read this slot has finished
            st["img"][slot].copy_(images_pinned, non_blocking=True)
            st["lab"][slot].copy_(labels_pinned, non_blocking=True)
            st["copied"][slot].record()
        cur.wait_event(st["copied"][slot])
        self.step(st["img"][slot], st["lab"][slot])
        st["loss_host"][slot].copy_(self.loss, non_blocking=True)
        st["done"][slot].record(cur)
        st["pending"][slot] = True
        if lag:
            # this step is enqueued; only now wait for the previous one, so the GPU always has a step queued
            prev = slot ^ 1
            prev_loss = None
            if st["pending"][prev]:
                st["done"][prev].synchronize()
                prev_loss = float(st["loss_host"][prev][0])
                st["pending"][prev] = False
            return prev_loss
        st["done"][slot].synchronize()
        st["pending"][slot] = False
        return float(st["loss_host"][slot][0])

    def drain(self):
        """Wait for the last enqueued step_from_host() and return its loss (None if nothing is pending)."""
        st = getattr(self, "_e2e", None)
        if st is None:
            return None
        last = st["slot"] ^ 1
        if not st["pending"][last]:
            return None
        st["done"][last].synchronize()
        st["pending"][last] = False
        return float(st["loss_host"][last][0])

    def grad_norm(self) -> torch.Tensor:
        """sqrt of the sum of squares the last optim_step clipped with (after the 1/world scaling)."""
        return self.sumsq[:1].sqrt()
