"""ctypes binding of libmgn_b200.so (the C ABI declared in include/mgn_b200.h).

This is the whole "torch extension": torch owns device memory and streams, this module
unwraps tensors to raw pointers + the current stream and calls the C entry points.  There
is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import re
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

ROOT = Path(__file__).resolve().parent
LIB_PATH = ROOT / "lib" / "libmgn_b200.so"
HEADER = ROOT.parent / "include" / "mgn_b200.h"

MGN_F32, MGN_BF16 = 0, 1
MGN_OK, MGN_EINVAL, MGN_EUNSUPPORTED, MGN_EALIGN, MGN_EWORKSPACE = 0, -1, -2, -3, -4

ACT_IDS = {
    None: 0, "identity": 0, "none": 0,
    "relu": 1, "silu": 2, "tanh": 3, "sigmoid": 4, "gelu": 5, "leaky_relu": 6, "elu": 7,
}

_CTYPE = {
    "int": c_int, "int64_t": c_int64, "size_t": c_size_t, "float": c_float,
    "mgn_stream_t": c_void_p, "void": None,
}


def _parse_header(text: str):
    """Return {name: (restype, [argtypes])} for every prototype in the header, so the Python
    binding cannot drift from include/mgn_b200.h."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\n\s*(const\s+char\s*\*|size_t|int)\s+(mgn_\w+)\s*\(([^;{]*?)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = c_char_p if "char" in ret else _CTYPE[ret.strip()]
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(c_void_p)
                else:
                    ty = a.replace("const", "").split()[0]
                    argtypes.append(_CTYPE[ty])
        protos[name] = (restype, argtypes)
    return protos


PROTOTYPES = _parse_header(HEADER.read_text())

_lib = None


class MGNError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise MGNError(
                f"{LIB_PATH} not found: build it with `python -m modulus_b200.build` "
                "(there is no CPU / PyTorch fallback for the MeshGraphNet kernels)"
            )
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError => header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    """Map a C-ABI status to the exception types the reference raises."""
    if rc == MGN_OK:
        return
    msg = load().mgn_error_string(rc).decode()
    text = f"libmgn_b200: {what}: {msg} (code {rc})"
    if rc == MGN_EINVAL:
        raise ValueError(text)
    if rc in (MGN_EUNSUPPORTED, MGN_EALIGN, MGN_EWORKSPACE):
        raise MGNError(text)
    raise MGNError(text)  # positive: cudaError_t


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
