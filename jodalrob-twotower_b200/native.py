"""ctypes binding of the C ABI declared in include/twotower_b200.h.

The prototypes are parsed from the header itself, so the binding cannot drift from the ABI.
There is NO fallback: if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_PKG_DIR)
HEADER = os.path.join(REPO_ROOT, "include", "twotower_b200.h")
LIB_PATH = os.path.join(_PKG_DIR, "csrc", "libtwotower_b200.so")

TT_OK = 0
TT_F32, TT_BF16 = 0, 1
TT_LAYOUT_SAMPLE_MAJOR, TT_LAYOUT_KEY_MAJOR = 0, 1
TT_ADAM_DENSE_EXACT, TT_ADAM_SPARSE, TT_ADAM_DENSE_UNTOUCHED = 0, 1, 2
TT_ADAM_LAZY_EXACT = 3  # engine-side name: tt_ebc_backward_adam_lazy / tt_ebc_forward_mixed_lazy (include/twotower_b200.h, tt_lazy_ctx)
HYPER_BYTES = 64

_SCALARS = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint32_t": ctypes.c_uint32,
    "uint64_t": ctypes.c_uint64, "float": ctypes.c_float, "double": ctypes.c_double, "tt_stream_t": ctypes.c_void_p,
}


class TwoTowerNativeError(RuntimeError):
    pass


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """{name: (restype, [argtypes])} for every `tt_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(const\s+char\s*\*|int64_t|int)\s+(tt_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else _SCALARS[ret.strip()]
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const ", "").split()[0]
                    argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.isfile(LIB_PATH):
            raise TwoTowerNativeError(
                f"CUDA extension not built: {LIB_PATH} is missing. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {os.path.dirname(LIB_PATH)}`. There is no CPU fallback.")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError as e:
                raise TwoTowerNativeError(f"{LIB_PATH} does not export {name} declared in {HEADER}") from e
            fn.restype = restype
            fn.argtypes = argtypes
        if self.cdll.tt_version() != 1:
            raise TwoTowerNativeError("ABI version mismatch between header and library")

    def last_error(self) -> str:
        return (self.cdll.tt_last_error() or b"").decode()

    def call(self, name: str, *args):
        """Call an int-status entry point; raise on a negative status."""
        rc = getattr(self.cdll, name)(*args)
        if rc != TT_OK:
            raise TwoTowerNativeError(f"{name} failed ({rc}): {self.last_error()}")

    def query(self, name: str, *args) -> int:
        return int(getattr(self.cdll, name)(*args))


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()
