"""ctypes loader for libkpf_b200.so.  Prototypes are parsed from include/kpf_b200.h so the header IS the ABI.

There is deliberately no fallback: if the library is missing or a symbol is absent this raises."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "kpf_b200.h")
LIB_PATH = os.path.join(HERE, "libkpf_b200.so")

_SCALARS = {"int": ctypes.c_int, "uint32_t": ctypes.c_uint32, "float": ctypes.c_float, "long long": ctypes.c_longlong,
            "cudaStream_t": ctypes.c_void_p}
ERRORS = {-1: "KPF_ERR_BAD_ARGUMENT", -2: "KPF_ERR_UNSUPPORTED"}


def parse_header(path=HEADER):
    """-> {name: [(ctype, argname), ...]} for every `int kpf_*(...)` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(kpf_\w+)\s*\((.*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                typ, arg = mm.group(1).strip(), mm.group(2)
                if typ.endswith("*"):
                    params.append((ctypes.c_void_p, arg))
                else:
                    params.append((_SCALARS[typ.replace("const ", "")], arg))
        protos[name] = params
    return protos


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m keypointfusion_b200.build` "
                               "(there is no CPU / PyTorch fallback for the fusion hot path)")
        L = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, params in _protos.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = ctypes.c_int
            fn.argtypes = [t for t, _ in params]
        _lib = L
    return _lib


def check(rc, name):
    if rc != 0:
        if rc < 0:
            raise RuntimeError(f"{name}: {ERRORS.get(rc, rc)}")
        try:   # positive codes are cudaError_t values
            rt = ctypes.CDLL("libcudart.so")
            rt.cudaGetErrorString.restype = ctypes.c_char_p
            msg = rt.cudaGetErrorString(ctypes.c_int(rc)).decode()
        except OSError:
            import torch
            msg = torch.cuda.cudart().cudaGetErrorString(torch.cuda.cudart().cudaError(rc)) if hasattr(torch.cuda.cudart(), "cudaError") else ""
        raise RuntimeError(f"{name}: CUDA error {rc}" + (f" ({msg})" if msg else ""))
