"""Build the reference's own CUDA extensions (UNMODIFIED sources, where they lie
under /root/reference) for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The resulting pybind11 modules are the GPU-side
ground truth the `-m gpu` parity tests compare our kernels with, and the
`ref_cuda` arm of bench.py.  Nothing in the product path may import them.

Sources compiled (never copied into this repo):
  /root/reference/raymarching/src/{raymarching.cu,bindings.cpp}   -> _ref/_raymarching.so
  /root/reference/gridencoder/src/{gridencoder.cu,bindings.cpp}   -> _ref/_gridencoder.so
  /root/reference/shencoder/src/{shencoder.cu,bindings.cpp}       -> _ref/_shencoder.so
  /root/reference/freqencoder/src/{freqencoder.cu,bindings.cpp}   -> _ref/_freqencoder.so

Flags follow the reference's own backend.py (raymarching/backend.py:6-12) with
one forced change: -std=c++14 -> -std=c++17 (torch 2.11 headers do not compile
as c++14).  Arch: compute_100a/sm_100a.

The reference's own Python for the path (nerf/renderer.py, nerf/network_grid.py, nerf/utils.py Trainer, nerf/provider.py,
encoding.py, activation.py, optimizer.py, main.py and the four operator wrapper packages) is byte-compiled — again from the
unmodified sources where they lie — into sourceless .pyc trees:
  oracle/_ref/refpy/      nerf/, encoding, activation, optimizer, main       (the host code that must run UNCHANGED on the drop-in ops)
  oracle/_ref/refpy_ops/  raymarching/, gridencoder/, freqencoder/, shencoder/ wrappers (bind to oracle/_ref/_*.so: the all-reference arm)
No reference source text enters the repository; the .pyc files are build outputs like the .so files.

Usage:  python oracle/build_ref.py [name ...]      (default: all four + the .pyc trees)
The outputs are git-ignored but travel to the GPU box with gpurun.
"""
import os
import shutil
import sys

REF = os.environ.get("SDF_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "_raymarching": ("raymarching", ["raymarching.cu", "bindings.cpp"]),
    "_gridencoder": ("gridencoder", ["gridencoder.cu", "bindings.cpp"]),
    "_shencoder": ("shencoder", ["shencoder.cu", "bindings.cpp"]),
    "_freqencoder": ("freqencoder", ["freqencoder.cu", "bindings.cpp"]),
}

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
    "-U__CUDA_NO_HALF2_OPERATORS__",
    "-gencode", "arch=compute_100a,code=sm_100a",
]
C_FLAGS = ["-O3", "-std=c++17"]


def build_one(name):
    from torch.utils.cpp_extension import load
    pkg, files = EXTS[name]
    srcs = [os.path.join(REF, pkg, "src", f) for f in files]
    for s in srcs:
        if not os.path.exists(s):
            raise FileNotFoundError(s)
    bdir = os.path.join(OUT, "build", name)
    os.makedirs(bdir, exist_ok=True)
    # torch appends its own -gencode from TORCH_CUDA_ARCH_LIST; keep it to 10.0a.
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    load(name=name, sources=srcs, extra_cflags=C_FLAGS,
         extra_cuda_cflags=NVCC_FLAGS, build_directory=bdir, verbose=True,
         is_python_module=False)
    so = os.path.join(bdir, name + ".so")
    shutil.copy2(so, os.path.join(OUT, name + ".so"))
    print("built", os.path.join(OUT, name + ".so"), flush=True)


PY_CORE = ["activation.py", "encoding.py", "optimizer.py", "main.py", "nerf/utils.py", "nerf/renderer.py", "nerf/network_grid.py",
           "nerf/provider.py"]
PY_OPS = ["raymarching/__init__.py", "raymarching/raymarching.py", "gridencoder/__init__.py", "gridencoder/grid.py",
          "freqencoder/__init__.py", "freqencoder/freq.py", "shencoder/__init__.py", "shencoder/sphere_harmonics.py"]


def build_refpy():
    """byte-compile the reference's host Python (unmodified, in place) into sourceless .pyc trees under oracle/_ref/"""
    import py_compile
    n = 0
    for sub, files in (("refpy", PY_CORE), ("refpy_ops", PY_OPS)):
        for rel in files:
            src = os.path.join(REF, rel)
            dst = os.path.join(OUT, sub, rel + "c")
            if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src) and not os.environ.get("SDF_REF_REBUILD"):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            n += 1
    print("reference python: %d file(s) byte-compiled into %s/{refpy,refpy_ops}" % (n, OUT), flush=True)


def main(argv):
    if not os.path.isdir(REF):
        print("reference tree not present (%s): using prebuilt oracle/_ref if any" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    names = argv or list(EXTS)
    if not argv:
        build_refpy()
    for n in names:
        if os.path.exists(os.path.join(OUT, n + ".so")) and not os.environ.get("SDF_REF_REBUILD"):
            print("up to date:", n)
            continue
        build_one(n)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
