"""Forward-only attention check (development aid): small cases first, prints the first failing one."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops
def ref(qkv, seqlens, H):
    outs, o, D = [], 0, H * 64
    for n in seqlens:
        t = qkv[o:o + n].float().reshape(n, 3, H, 64).permute(1, 2, 0, 3)
        a = ((t[0] @ t[1].transpose(-2, -1)) * 0.125).softmax(-1)
        outs.append((a @ t[2]).transpose(0, 1).reshape(n, D)); o += n
    return torch.cat(outs, 0)
import json
CASES = json.loads(os.environ.get("CASES", "[[[50],1],[[96],1],[[128],2],[[197],2],[[257],1],[[257,257,257,257],12]]"))
for seqlens, H in CASES:
    T, D = sum(seqlens), H * 64
    qkv = torch.randn(T, 3 * D, device="cuda").bfloat16()
    r = ref(qkv, seqlens, H)
    out, lse = ops.attn_fwd(qkv, H, 0.125, len(seqlens), max(seqlens))
    torch.cuda.synchronize()
    e = float((out.float() - r).norm() / r.norm())
    print(seqlens[:2], len(seqlens), H, "rel err", e, flush=True)
