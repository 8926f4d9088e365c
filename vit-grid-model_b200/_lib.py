"""ctypes binding of libvitgrid.so.

The prototypes are parsed from ``include/vitgrid.h`` so the binding cannot drift from the C ABI.  Loading fails
loudly: there is no CPU or PyTorch fallback for any op in this package.
"""
from __future__ import annotations

import ctypes
import os
import re

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "vitgrid.h")

_CTYPES = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float}


class VitGridError(RuntimeError):
    pass


def parse_header(path: str = HEADER):
    """-> {name: (restype, [(argname, ctype)])} for every VG_API declaration."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"VG_API\s+([\w\s\*]+?)\s*\b(vg_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "*" in ret else _CTYPES[ret.replace("const", "").strip()]
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                am = re.match(r"^(.*?)(\w+)$", a)
                ctype_s, aname = am.group(1).strip(), am.group(2)
                if "*" in ctype_s:
                    argl.append((aname, ctypes.c_void_p))
                else:
                    argl.append((aname, _CTYPES[ctype_s.replace("const", "").strip()]))
        protos[name] = (restype, argl)
    return protos


_lib = None
_protos = None


def load(build_if_missing: bool = True):
    """Load (building first when the in-tree .so is missing or stale and nvcc is present)."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if build_if_missing and _build.is_stale():
        # a stale binary is never loaded against the freshly parsed header: rebuild (under a file lock: the ranks of a
        # multi-process job all arrive here) or fail.  A box without nvcc runs the prebuilt .so that travelled with the tree.
        if _build.have_nvcc():
            try:
                _build.build_library()
            except Exception as e:
                raise VitGridError(f"libvitgrid.so is missing or older than its sources and the rebuild failed: {e}") from e
        elif os.path.exists(path):
            import warnings
            warnings.warn("libvitgrid.so is older than its sources and nvcc is not available to rebuild it")
    if not os.path.exists(path):
        raise VitGridError(f"{path} not found: run `python __graft_entry__.py build` (no CPU fallback exists)")
    lib = ctypes.CDLL(path)
    _protos = parse_header()
    for name, (restype, args) in _protos.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise VitGridError(f"libvitgrid.so does not export {name} (declared in include/vitgrid.h)")
        fn.restype = restype
        fn.argtypes = [t for _, t in args]
    _lib = lib
    return lib


def prototypes():
    load()
    return _protos


# optional per-call CUDA-event trace: set to a list to record (entry point, tag, start_event, end_event)
TRACE = None
TRACE_TAG = ""


def call(name: str, *args):
    """Call an int-returning entry point; raise VitGridError with vg_last_error() on failure."""
    lib = load()
    if TRACE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        TRACE.append((name, TRACE_TAG, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise VitGridError(f"{name} failed: {lib.vg_last_error().decode(errors='replace')}")


def raise_device_errors():
    """raise what kernels of EARLIER calls flagged (vg_device_error; meaningful for work that has completed)"""
    if _lib is None:
        return
    code = _lib.vg_device_error(0)
    if code:
        _lib.vg_device_error(1)
        if code & 1:
            raise IndexError("a timestamp's month / day / hour was outside the embedding tables 13 / 32 / 25 "
                             "(metnet3.py:392); the affected predictions are NaN")
        raise VitGridError(f"device-side error word {code:#x}")


def launch_count() -> int:
    return int(load().vg_launch_count())


_device_ok = False


def require_device():
    """The product path runs on an sm_100 GPU only."""
    global _device_ok
    if _device_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise VitGridError("vit_grid_model_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    call("vg_device_check")
    _device_ok = True
