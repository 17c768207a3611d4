"""ctypes loader for libapla_b200.so (the C ABI of include/apla_b200.h).

The product path has no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
Function prototypes are parsed from the header so that the Python side cannot drift from the ABI.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libapla_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "apla_b200.h")

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32,
    "apla_stream_t": ctypes.c_void_p, "apla_engine_t": ctypes.c_void_p, "double": ctypes.c_double,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """-> {name: (restype, [argtypes])} for every prototype in the header."""
    with open(path) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int|void|apla_engine_t)\s+(apla_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        if ret.startswith("const"):
            restype = ctypes.c_char_p
        elif ret == "void":
            restype = None
        else:
            restype = _CTYPES[ret]
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    toks = [t for t in a.replace("const", " ").split() if t]
                    argtypes.append(_CTYPES[toks[0]])
        protos[name] = (restype, argtypes)
    return protos


class BlockWeights(ctypes.Structure):
    """`apla_block_weights` of include/apla_b200.h (field for field; the library reports its sizeof for a check)."""
    _fields_ = ([(n, ctypes.c_void_p) for n in
                 ("wqkv", "wqkvT", "wproj", "wprojT", "wfc1", "wfc1T", "wfc2", "wfc2T",
                  "bqkv", "bproj", "bfc1", "bfc2", "ln1w", "ln1b", "ln2w", "ln2b", "g1", "g2", "idx", "rowmap")]
                + [(n, ctypes.c_int32) for n in ("D", "H", "hidden", "r", "r_pad")]
                + [(n, ctypes.c_float) for n in ("eps1", "eps2", "scale")])


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m apla_b200.build` (nvcc, sm_100a). "
                "apla_b200 has no CPU or PyTorch fallback.")
        dll = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(dll, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if dll.apla_block_weights_size() != ctypes.sizeof(BlockWeights):
            raise RuntimeError("apla_block_weights: the ctypes mirror in apla_b200/_lib.py and the header disagree "
                               f"({ctypes.sizeof(BlockWeights)} vs {dll.apla_block_weights_size()} bytes)")
        self._dll = dll
        return dll

    def last_error(self) -> str:
        return self.load().apla_last_error().decode()

    def call(self, name: str, *args):
        rc = getattr(self.load(), name)(*args)
        if rc != 0:
            raise RuntimeError(f"{name} failed (rc={rc}): {self.last_error()}")


LIB = _Lib()


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


_checked = False


def require_device():
    """Raise unless the CUDA library is present and the current device is an sm_100 part."""
    global _checked
    if _checked:
        return
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("apla_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    LIB.call("apla_device_check")
    _checked = True
