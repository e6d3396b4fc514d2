"""ctypes binding of libsdf_b200.so (C ABI: include/sdf_b200.h).

The prototypes are parsed from the header, so the header is the single source
of truth for the boundary.  There is no fallback: if the library is missing or
a call fails, a RuntimeError is raised (reference behaviour: TORCH_CHECK /
std::runtime_error -> RuntimeError, gridencoder/src/gridencoder.cu:392,468).
"""
import ctypes as C
import os
import re

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))       # stable-dreamfusion_b200/
_REPO_ROOT = os.path.dirname(_PKG_ROOT)
LIB_PATH = os.path.join(_PKG_ROOT, "lib", "libsdf_b200.so")
HEADER_PATH = os.path.join(_REPO_ROOT, "include", "sdf_b200.h")

_SCALARS = {
    "int": C.c_int, "uint32_t": C.c_uint32, "int32_t": C.c_int32, "float": C.c_float, "uint64_t": C.c_uint64,
    "int64_t": C.c_int64, "long long": C.c_longlong, "size_t": C.c_size_t, "double": C.c_double,
}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [(ctype, argname), ...])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|long\s+long|int)\s+(sdf_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = C.c_char_p if "char" in ret else (C.c_longlong if "long" in ret else C.c_int)
        argl = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                ty, an = mm.group(1).strip(), mm.group(2)
                if "*" in ty:
                    argl.append((C.c_void_p, an))
                else:
                    ty = ty.replace("const ", "").strip()
                    argl.append((_SCALARS[ty], an))
        protos[name] = (restype, argl)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"sdf_b200: native library not found at {LIB_PATH}. Build it with `python __graft_entry__.py build` "
                f"(or `make -C {os.path.join(_PKG_ROOT, 'csrc')}`); there is no CPU/PyTorch fallback.")
        self.cdll = C.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, args) in self.protos.items():
            fn = getattr(self.cdll, name)       # AttributeError -> symbol declared but not exported
            fn.restype = restype
            fn.argtypes = [t for t, _ in args]

    def last_error(self):
        s = self.cdll.sdf_last_error()
        return s.decode() if s else ""

    def call(self, name, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            raise RuntimeError(f"{name} failed (rc={rc}): {self.last_error()}")

    def query(self, name, *args):
        """value-returning entry points (sizes): no error protocol"""
        return getattr(self.cdll, name)(*args)


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def call(name, *args):
    lib().call(name, *args)


def query(name, *args):
    return lib().query(name, *args)


def feat_stash_budget_bytes():
    """largest feature stash (fused field forward -> backward) a caller may allocate; beyond it the backward re-gathers"""
    return int(float(os.environ.get("SDF_FEAT_STASH_GB", "12")) * (1 << 30))


def ptr(t):
    """device (or pinned-host) pointer of a tensor, or NULL for None."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sdf_b200: expected a CUDA tensor (there is no CPU path)")


class PinnedRing:
    """Small host -> device uploads that never block the host: a ring of pinned staging slots, each rewritten only after the copy
    that read it has completed (event per slot).  A plain `cpu_tensor.to(device)` from pageable memory synchronises the stream —
    one such call per step stops the host from running ahead of the GPU."""

    def __init__(self, numel, device, slots=4, dtype=torch.float32):
        self.device = device
        self.slots = [torch.zeros(numel, dtype=dtype).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.next = 0

    def upload(self, src):
        """src: CPU tensor or numpy array with at most `numel` elements -> device tensor of the same shape"""
        src = torch.as_tensor(src)
        k, n = self.next, src.numel()
        slot = self.slots[k]
        if n > slot.numel():
            raise ValueError(f"PinnedRing: {n} elements exceed the slot capacity {slot.numel()}")
        self.next = (k + 1) % len(self.slots)
        if self.events[k] is not None:
            self.events[k].synchronize()
        slot[:n].copy_(src.reshape(-1).to(slot.dtype))
        out = slot[:n].to(self.device, non_blocking=True)
        self.events[k] = torch.cuda.Event()
        self.events[k].record()
        return out.view(src.shape)
