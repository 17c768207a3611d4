"""Helpers that turn a host ViT into an APLA model: drop-ins for src/apla/apla_vit.py:11-101.

`replace_attn_with_apla(model, config, attn_module)` swaps every `block.attn` for an APLA module, cloning the
pretrained qkv weights and splitting the projection rows by index; `build_apla(config, model, attn_class,
is_multi_gpu)` applies the freeze policy and dispatches on the attention class.  Behaviour kept from the reference:
  * `config` must answer both `hasattr(config, 'inds_path')` and `'inds_path' in config` (:12, :77);
  * with `inds_path`, the index order is the json's trainable list followed by the ascending complement (:20-24);
  * multi-GPU + partial_size == 'full' leaves the stock attention in place and trains `attn.proj.*` (:65-75);
  * multi-GPU + partial requires `inds_path` (AssertionError, :77); unknown attn_class -> NotImplementedError (:84-89).
"""
from __future__ import annotations

import json
import os

import torch

from .appla_attn import APLA_Attention, print_ddp
from .appla_attn_mem_eff import APLA_MemEffAttention


def _load_json(path):
    with open(os.path.abspath(path), "r") as f:
        return json.load(f)


def replace_attn_with_apla(model, config, attn_module):
    use_file = hasattr(config, "inds_path")
    table = _load_json(config.inds_path) if use_file else None
    if use_file:
        print_ddp(f"Registering inds based on path: {config.inds_path}")
    for i, block in enumerate(model.blocks):
        old = block.attn
        indices = None
        if use_file:
            chosen = list(table[f"block_{i}"])
            taken = set(chosen)
            indices = torch.tensor(chosen + [j for j in range(old.dim) if j not in taken])
        new = attn_module(config=config, dim=old.dim, indices=indices, num_heads=old.num_heads,
                          qkv_bias=old.qkv.bias is not None, qk_scale=old.scale, attn_drop=old.attn_drop.p,
                          proj_drop=old.proj_drop.p)
        with torch.no_grad():
            new.qkv.weight.data = old.qkv.weight.data.clone()
            if old.qkv.bias is not None:
                new.qkv.bias.data = old.qkv.bias.data.clone()
            w = old.proj.weight.data.clone()
            new.proj_weight1.data = w[new.trainable_inds, :]
            new.proj_weight2.data = w[new.freezed_inds, :]
            if old.proj.bias is not None:
                b = old.proj.bias.data.clone()
                new.proj_bias1.data = b[new.trainable_inds]
                new.proj_bias2.data = b[new.freezed_inds]
        block.attn = new
        print_ddp(f"Replaced {old.__class__.__name__} in block {i} with {new.__class__.__name__}")


def build_apla(config, model, attn_class, is_multi_gpu=False):
    if is_multi_gpu:
        if config.partial_size == "full":
            for name, p in model.named_parameters():
                p.requires_grad = "attn.proj" in name
                print_ddp(f"Building apla -- Set requires_grad to {p.requires_grad} for: {name}")
            return model
        assert "inds_path" in config, '"inds_path" should be present with multi-gpu training with random sampling'

    for p in model.parameters():
        p.requires_grad = False

    if attn_class == "apla_attn":
        attn_module = APLA_Attention
    elif attn_class == "apla_attn_mem_eff":
        attn_module = APLA_MemEffAttention
    else:
        raise NotImplementedError

    replace_attn_with_apla(model=model, config=config, attn_module=attn_module)
    for name, p in model.named_parameters():
        print_ddp(f"Building apla -- {name} requires_grad: {p.requires_grad}")
    print_ddp("Successfully built APLA-enabled model")
    return model
