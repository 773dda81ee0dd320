"""ctypes binding of libadamml_b200.so, generated from include/adamml_b200.h.

The header is the single source of truth for the C-ABI: every prototype in it is parsed
here into a ctypes signature, so a symbol that is declared but not exported (or vice versa)
fails at import time.  There is NO fallback: if the shared library is missing the import
raises, and every op raises if the CUDA call reports an error.
"""
import ctypes
import os
import re

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
HEADER = os.path.join(_ROOT, "include", "adamml_b200.h")
LIB_PATH = os.path.join(_PKG, "lib", "libadamml_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
ERR_UNSUPPORTED = 3

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "cudaStream_t": ctypes.c_void_p,
    "const char*": ctypes.c_char_p,
    "void": None,
}


def _ctype(decl):
    decl = decl.strip()
    if decl.endswith("*") and decl != "const char*":
        return ctypes.c_void_p
    return _CTYPES[decl]


def parse_header(path=HEADER):
    """-> {name: (restype, [(argtype, argname), ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\s*\b(adamml_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        ret = ret.replace(" *", "*")
        arglist = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                am = re.match(r"(.*?)(\w+)$", a)
                t = am.group(1).strip().replace(" *", "*")
                arglist.append((t, am.group(2)))
        protos[name] = (ret, arglist)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(adamml_b200 has no CPU / PyTorch fallback)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the symbol is not exported
            fn.restype = _ctype(ret)
            fn.argtypes = [_ctype(t) for t, _ in args]

    def last_error(self):
        return self.cdll.adamml_last_error().decode()


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def _conv(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a


# bench.py sets this to a list to time every call with CUDA events on the launching stream:
# entries are (name, start_event, end_event, scalar_args)
PROFILE = None


def call(name, *args, allow_unsupported=False):
    """Call `adamml_<name>` with torch tensors / scalars; the current CUDA stream is appended."""
    L = lib()
    fn = getattr(L.cdll, "adamml_" + name)
    stream = torch.cuda.current_stream().cuda_stream
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*[_conv(a) for a in args], stream)
        e1.record()
        PROFILE.append((name, e0, e1, tuple("T" if isinstance(a, torch.Tensor) else a for a in args)))
    else:
        rc = fn(*[_conv(a) for a in args], stream)
    if rc != 0:
        if allow_unsupported and rc == ERR_UNSUPPORTED:
            return rc
        raise RuntimeError(f"adamml_{name} failed (rc={rc}): {L.last_error()}")
    return 0


def launch_count():
    return int(lib().cdll.adamml_launch_count())


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported activation dtype {dt}")
