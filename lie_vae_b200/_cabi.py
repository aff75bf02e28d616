"""ctypes binding of liblievae_sm100a.so -- the reference-side stub of INTEGRATION.md.

Prototypes are read from ``include/lievae.h`` (single source of truth).  There is
no fallback: if the library is missing or a call fails, a ``RuntimeError`` is
raised.  Tensors handed to ``call`` must be CUDA, contiguous and of the dtype the
entry point names; the autograd wrappers in ``_ops.py`` guarantee that.
"""
import ctypes
import os
import re
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIEVAE_LIB") or os.path.join(HERE, "liblievae_sm100a.so")   # override: tools/exp variants only
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "lievae.h")

_lock = threading.Lock()
_lib = None

_DECL = re.compile(r"^\s*(int64_t|int|const char\*)\s+(lv_\w+)\s*\(([^;]*?)\)\s*;", re.M | re.S)


def _ctype(param):
    p = re.sub(r"/\*.*?\*/", "", param).strip()
    if p in ("void", ""):
        return None
    if "*" in p:
        return ctypes.c_void_p
    if p.startswith("int64_t"):
        return ctypes.c_int64
    if p.startswith("int"):
        return ctypes.c_int
    raise ValueError("unhandled parameter %r" % param)


def header_prototypes(path=HEADER_PATH):
    """{name: (restype, [argtypes])} for every ``lv_*`` declaration in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for ret, name, params in _DECL.findall(text):
        args = [a for a in (_ctype(p) for p in params.split(",")) if a is not None]
        res = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "const char*": ctypes.c_char_p}[ret]
        protos[name] = (res, args)
    return protos


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "lie_vae_b200: %s is missing -- build it with `python -m lie_vae_b200._build` "
                        "(there is no CPU or PyTorch fallback)" % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in header_prototypes().items():
                    fn = getattr(handle, name)          # AttributeError if the header and library disagree
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def last_error():
    msg = lib().lv_last_error()
    return msg.decode() if msg else ""


_fn_cache = {}


def call(name, *args):
    """Invoke an int-returning entry point; raise RuntimeError on a non-zero status."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(lib(), name)
    rc = fn(*args)
    if rc != 0:
        kind = "argument error" if rc < 0 else "CUDA error"
        raise RuntimeError("%s failed (%s %d): %s" % (name, kind, rc, last_error()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL) as the plain integer ctypes converts to ``void*``."""
    return None if t is None else t.data_ptr()
