"""ctypes binding of libhgb.so (the C-ABI declared in include/hgb.h).

The prototypes are read from the header itself, so the Python side can never drift from the ABI.
There is NO fallback: if the shared library is missing, ``lib()`` raises; every op in
``hydragnn_b200.ops`` goes through ``call()`` which raises ``RuntimeError`` with the library's own
message on a non-zero return code.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "hgb.h")
LIB_PATH = os.path.join(_HERE, "csrc", "libhgb.so")

_CT = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
       "double": ctypes.c_double, "hgb_stream_t": ctypes.c_void_p}

_lib = None
_protos = None


def prototypes():
    """{name: (restype_str, [(ctype_str, argname), ...])} parsed from include/hgb.h."""
    global _protos
    if _protos is None:
        src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        _protos = {}
        for m in re.finditer(r"(const char\*|int64_t|int32_t|int)\s+(hgb_\w+)\s*\(([^)]*)\)\s*;", src):
            ret, name, args = m.group(1), m.group(2), m.group(3).strip()
            parsed = []
            if args and args != "void":
                for a in args.split(","):
                    a = " ".join(a.split())
                    mm = re.match(r"(.*?)(\w+)$", a)
                    parsed.append((mm.group(1).strip(), mm.group(2)))
            _protos[name] = (ret, parsed)
    return _protos


def _ctype(t):
    if "*" in t:
        return ctypes.c_void_p
    return _CT[t.replace("const ", "").strip()]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "hydragnn_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `python hydragnn_b200/build.py`).  There is no CPU / PyTorch fallback for the hot path." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (ret, args) in prototypes().items():
            fn = getattr(L, name)
            fn.restype = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "const char*": ctypes.c_char_p}[ret]
            fn.argtypes = [_ctype(t) for t, _ in args]
        _lib = L
    return _lib


_trace = None


def call(name, *args):
    """Invoke an int-returning entry point; raise on failure."""
    L = lib()
    if _trace is not None:
        before = int(L.hgb_launch_count())
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise RuntimeError("libhgb %s failed (%d): %s" % (name, rc, L.hgb_last_error().decode()))
    if _trace is not None:
        names = [n for _, n in prototypes()[name][1]]
        _trace.append((name, dict(zip(names, args)), int(L.hgb_launch_count()) - before))


def trace_begin():
    """Start recording (entry point, arguments by header name, kernels launched) of every call -- used by bench.py to attribute
    profiled kernel durations to C-ABI calls and to evaluate their algorithmic-byte formulas."""
    global _trace
    _trace = []


def trace_end():
    global _trace
    out, _trace = _trace, None
    return out


def query(name, *args):
    """Invoke a size-query entry point (returns its value)."""
    return getattr(lib(), name)(*args)


def launch_count():
    return int(lib().hgb_launch_count())
