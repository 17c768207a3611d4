"""Builds tests/emu/build/libssl_emu.so: apla_b200/csrc/ssl.cu compiled for the CPU over tests/emu/cuda_emu.h.

The kernel source is used AS IS; three mechanical rewrites make it C++:
  * `kernel<<<grid, block, smem, stream>>>(args);`  ->  `emu_launch(grid, block, smem, [&] { kernel(args); });`
  * `extern __shared__ float sm[];`                 ->  `float* sm = emu_dyn_smem;`
  * `#include "common.cuh"` / `"kernels.cuh"`       ->  `#include "cuda_emu.h"`
and the C-ABI wrappers of capi.cu for these entry points are regenerated from include/apla_b200.h's prototypes.
TEST INFRASTRUCTURE ONLY (tests/test_ssl_emu.py)."""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "apla_b200", "csrc", "ssl.cu")
HDR = os.path.join(ROOT, "include", "apla_b200.h")
OUT = os.path.join(HERE, "build")
ENTRY = {  # C-ABI name -> launcher in namespace apla
    "apla_softmax_center": "ssl_softmax_center", "apla_colsum_f32": "ssl_colsum_f32", "apla_center_ema": "ssl_center_ema",
    "apla_soft_ce_fwd": "ssl_soft_ce_fwd", "apla_soft_ce_bwd": "ssl_soft_ce_bwd", "apla_sum_f32": "ssl_sum_f32",
    "apla_l2norm_fwd": "ssl_l2norm_fwd", "apla_l2norm_bwd": "ssl_l2norm_bwd", "apla_weightnorm_fwd": "ssl_weightnorm_fwd",
    "apla_weightnorm_bwd": "ssl_weightnorm_bwd", "apla_koleo_fwd": "ssl_koleo_fwd", "apla_koleo_bwd": "ssl_koleo_bwd",
    "apla_ema_update": "ssl_ema", "apla_ssl_objective": "ssl_objective", "apla_soft_ce_fwd_bwd": "ssl_soft_ce_fwd_bwd",
    "apla_sk_exp": "ssl_sk_exp", "apla_sk_normalize": "ssl_sk_normalize",
}


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite(src: str) -> str:
    n_launch = 0

    def launch(m):
        nonlocal n_launch
        n_launch += 1
        cfg = _split_top(m.group(2))
        assert len(cfg) == 4, cfg
        return f"emu_launch({cfg[0]}, {cfg[1]}, {cfg[2]}, [&] {{ {m.group(1)}({m.group(3)}); }});"

    out = re.sub(r"([A-Za-z_]\w*(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", launch, src, flags=re.S)
    assert "<<<" not in out and n_launch >= 16, n_launch
    out, n = re.subn(r"extern\s+__shared__\s+float\s+sm\[\];", "float* sm = emu_dyn_smem;", out)
    assert n == 2, n
    out = out.replace('#include "common.cuh"', '#include "cuda_emu.h"').replace('#include "kernels.cuh"', "")
    assert "cuda_emu.h" in out
    return out


def c_abi_wrappers() -> str:
    hdr = re.sub(r"/\*.*?\*/", " ", open(HDR).read(), flags=re.S)
    lines = ['extern "C" {', "const char* emu_last_error(void) { return apla::emu_error; }"]
    for name, impl in ENTRY.items():
        m = re.search(r"int\s+" + name + r"\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
        assert m, name
        params = [" ".join(p.split()) for p in m.group(1).split(",")]
        args = [re.split(r"[\s\*]+", p)[-1] for p in params]
        assert args[-1] == "stream"
        sig = ", ".join(p.replace("apla_stream_t", "void*") for p in params)
        lines.append(f"int {name}({sig}) {{ return apla::{impl}({', '.join(args)}); }}")
    lines.append("}")
    return "\n".join(lines) + "\n"


def build(force=False) -> str:
    os.makedirs(OUT, exist_ok=True)
    cpp = rewrite(open(SRC).read()) + "\n" + c_abi_wrappers()
    emu_h = open(os.path.join(HERE, "cuda_emu.h")).read()
    digest = hashlib.sha256((cpp + emu_h).encode()).hexdigest()
    lib, stamp, gen = (os.path.join(OUT, f) for f in ("libssl_emu.so", "libssl_emu.sha", "ssl_emu.cpp"))
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == digest:
        return lib
    with open(gen, "w") as f:
        f.write(cpp)
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-I", HERE, gen, "-o", lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the emulated ssl.cu:\n" + r.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(digest)
    return lib


if __name__ == "__main__":
    print(build(force=True))
