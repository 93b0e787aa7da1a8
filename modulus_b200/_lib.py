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
    for m in re.finditer(r"\n\s*(const\s+char\s*\*|size_t|int64_t|int)\s+(mgn_\w+)\s*\(([^;{]*?)\)\s*;", text):
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
# profiling hooks: bound only when the library was built with -DMGN_DEBUG_HOOKS (tools/prof_kernels.py)
DEBUG_HEADER = HEADER.parent / "mgn_b200_debug.h"
DEBUG_PROTOTYPES = _parse_header(DEBUG_HEADER.read_text()) if DEBUG_HEADER.exists() else {}

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
        for name, (restype, argtypes) in DEBUG_PROTOTYPES.items():
            fn = getattr(lib, name, None)
            if fn is not None:
                fn.restype = restype
                fn.argtypes = argtypes
        _verify_digest(lib)
        _lib = lib
    return _lib


def _verify_digest(lib) -> None:
    """The library must have been built from the sources next to it (same header the prototypes above were parsed
    from).  A mismatch is a stale build: calling it would be undefined behaviour, so it is an error, not a warning."""
    import os

    from . import build

    if os.environ.get("MGN_ALLOW_STALE_LIB") == "1" or not build.CSRC.exists():
        return
    have, want = lib.mgn_build_digest().decode(), build._digest()
    if have != want:
        raise MGNError(
            f"{LIB_PATH} was built from different sources or flags (library {have[:12]}, sources {want[:12]}): "
            "rebuild it with `python -m modulus_b200.build`")


def has_debug_hooks() -> bool:
    return all(hasattr(load(), n) for n in DEBUG_PROTOTYPES)


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
    if PROFILE.active and (PROFILE.only is None or PROFILE.only == name):
        PROFILE.record(name, args)
        return
    check(getattr(load(), name)(*args), name)


# ----------------------------------------------------------------------------------------
# measurement hooks (bench.py): CUDA-event bracketing of C-ABI calls on the calling stream
# ----------------------------------------------------------------------------------------
_DT_BYTES = {MGN_F32: 4, MGN_BF16: 2}


def algorithmic_work(name: str, a) -> "tuple[str, float] | None":
    """(bound, amount) of ONE call from its arguments: ("hbm", algorithmic bytes) for the
    gather / scatter / elementwise entry points, ("tensor", useful flops) for the dense ones.
    Formulas: DESIGN.md section 'Kernels and their rooflines' (SURVEY 8d)."""
    if name == "mgn_edge_block_bwd_tc":
        return "tensor", 20.0 * 128 * 128 * a[6]  # useful flops of the reference formulation (dgrad + wgrad = 2 x forward)
    if name == "mgn_mlp3_bwd_tc":
        small_in, g1, g2, M = a[3], a[5], a[9], a[17]
        if g2:
            return "tensor", 20.0 * 128 * 128 * M  # edge block: dgrad + wgrad = 2 x forward
        if g1:
            return "tensor", 16.0 * 128 * 128 * M  # node block
        k1 = small_in if small_in > 0 else 128
        return "tensor", 4.0 * M * (k1 * 128 + 2 * 128 * 128)
    if name == "mgn_mlp3_fwd2_tc":
        small_in, g1, g2, M = a[3], a[5], a[9], a[15]
        if g2:
            return "tensor", 10.0 * 128 * 128 * M
        if g1:
            return "tensor", 8.0 * 128 * 128 * M
        k1 = small_in if small_in > 0 else 128
        return "tensor", 2.0 * M * (k1 * 128 + 2 * 128 * 128)
    if name in ("mgn_edge_block_fwd_tc", "mgn_edge_block_fwd_part_tc"):
        return "tensor", 10.0 * 128 * 128 * a[9]
    if name == "mgn_node_block_fwd_tc":
        return "tensor", 8.0 * 128 * 128 * a[5]
    if name == "mgn_node_gemm_tc":
        kb, M, nb, res = a[2], a[3], a[6], a[7]
        return "hbm", 2.0 * M * 128 * (kb + nb + (1 if res else 0))  # one pass over x, (residual,) out
    if name == "mgn_gemm_bf16_tc":
        return "tensor", 2.0 * a[2] * a[3] * a[6]
    if name == "mgn_linear128_tc":
        return "tensor", 2.0 * a[2] * 128 * 128
    if name == "mgn_wgrad_tc":
        jb, M = a[2], a[5]
        return "tensor", 2.0 * M * 128 * 128 * jb
    if name == "mgn_linear_fwd":
        M, K, N = a[3], a[4], a[7]
        return "tensor", 2.0 * M * K * N
    if name == "mgn_linear_bwd_data":
        M, N, K = a[2], a[3], a[5]
        return "tensor", 2.0 * M * K * N
    if name == "mgn_linear_bwd_weight":
        M, N, K = a[4], a[5], a[6]
        return "tensor", 2.0 * M * K * N
    if name in ("mgn_segment_sum", "mgn_segment_sum_balanced"):
        b, D, n_seg = _DT_BYTES[a[0]], a[4], a[7]
        return "hbm", None if PROFILE.n_edges is None else (PROFILE.n_edges + n_seg) * D * b + 4.0 * n_seg
    if name == "mgn_gather_rows":
        b, D, rows = _DT_BYTES[a[0]], a[4], a[6]
        return "hbm", 2.0 * rows * D * b
    if name == "mgn_concat_efeat_fwd":
        b, De, Ds, Dd, E = _DT_BYTES[a[0]], a[2], a[4], a[6], a[9]
        return "hbm", E * (De + De + Ds + Dd) * b + 8.0 * E  # + one pass over the node tables (not known here)
    if name == "mgn_layernorm_fwd":
        b, M, D = _DT_BYTES[a[0]], a[2], a[3]
        return "hbm", (3.0 if a[7] else 2.0) * M * D * b
    if name == "mgn_layernorm_bwd":
        b, M, D = _DT_BYTES[a[0]], a[6], a[7]
        return "hbm", 3.0 * M * D * b
    if name in ("mgn_act_bwd", "mgn_add"):
        b, n = _DT_BYTES[a[0]], (a[5] if name == "mgn_act_bwd" else a[4])
        return "hbm", 3.0 * n * b
    return None


class _Profile:
    def __init__(self):
        self.active = False
        self.only = None
        self.n_edges = None
        self._ev = []

    def start(self, all_symbols: bool = True, only: "str | None" = None, n_edges: "int | None" = None) -> None:
        self.only = None if all_symbols else only
        if not all_symbols and only is None:
            return
        self.n_edges = n_edges
        self._ev = []
        self.active = True

    def record(self, name, args) -> None:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(load(), name)(*args)
        e1.record()
        check(rc, name)
        self._ev.append((name, e0, e1, algorithmic_work(name, args)))

    def stop(self) -> dict:
        import torch

        if not self.active:
            return {}
        self.active = False
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, work in self._ev:
            d = out.setdefault(name, {"ms": 0.0, "calls": 0, "bound": None, "work": 0.0, "per_call": []})
            ms = e0.elapsed_time(e1)
            d["ms"] += ms
            d["calls"] += 1
            if work is not None and work[1] is not None:
                d["bound"] = work[0]
                d["work"] += work[1]
                d["per_call"].append((ms, work[1]))
        self._ev = []
        return out


PROFILE = _Profile()
