"""CPU: the C-ABI library loads without a GPU and exports every symbol include/sdf_b200.h declares."""
import ctypes
import os
import subprocess

from sdf_b200 import _lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python __graft_entry__.py build"
    protos = _lib.parse_header()
    assert len(protos) >= 20
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, f"declared in include/sdf_b200.h but not exported: {missing}"
    # and nothing is exported that the header does not declare
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T sdf_" in l}
    extra = exported - set(protos)
    assert not extra, f"exported but undeclared: {extra}"


def test_no_torch_types_and_no_foreign_deps():
    out = subprocess.check_output(["ldd", _lib.LIB_PATH], text=True)
    assert "torch" not in out and "c10" not in out and "python" not in out, out


def test_argument_errors_are_reported_without_a_gpu():
    L = _lib.lib()
    assert L.cdll.sdf_abi_version() >= 1
    # null pointers are rejected before any launch
    rc = L.cdll.sdf_near_far_from_aabb(None, None, None, 4, 0.2, None, None, None)
    assert rc == -1 and "null" in L.last_error()
    rc = L.cdll.sdf_sh_encode_forward(1, 1, 4, 2, 4, None, None)
    assert rc == -1 and "input dim" in L.last_error()
    # N == 0 is a no-op success (empty inputs)
    assert L.cdll.sdf_near_far_from_aabb(None, None, None, 0, 0.2, None, None, None) == 0


def test_argument_errors_of_the_sd_kernels_without_a_gpu():
    """contract checks of the newer entry points are made before anything touches CUDA"""
    L = _lib.lib()
    fake = ctypes.c_void_p(0x10000)          # non-null, 16-byte aligned, never dereferenced: every call below must fail earlier
    # flash attention: unsupported head dim -> SDF_ERR_UNSUPPORTED (-3); misaligned strides -> SDF_ERR_ARG (-1)
    assert L.cdll.sdf_flash_attention(fake, fake, fake, fake, 1, 1, 8, 8, 48, 48, 48, 48, ctypes.c_float(1.0), None) == -3
    assert "head dim" in L.last_error()
    assert L.cdll.sdf_flash_attention(fake, fake, fake, fake, 1, 1, 8, 8, 40, 41, 40, 40, ctypes.c_float(1.0), None) == -1
    assert L.cdll.sdf_flash_attention(None, None, None, None, 0, 1, 0, 8, 40, 40, 40, 40, ctypes.c_float(1.0), None) == 0      # empty batch: no-op
    # GEMM plans: K extent not a multiple of 64, tile widths the kernel is not instantiated for, CTA pairs on batched products
    LL = ctypes.c_longlong
    def plan(Cin=64, taps=1, block_n=128, pair=0, w_sy=0, act=0, N=128, splitk=1):
        return L.cdll.sdf_gemm_plan_create(fake, LL(64), LL(64 * 8), LL(64 * 64), 64, fake, LL(64), LL(w_sy), LL(0), 64, N, 1, 8, 8, Cin, taps, N,
                                           fake, LL(N), LL(N * 8), LL(N * 64), None, None, 0, None, LL(0), LL(0), LL(0), act, ctypes.c_float(1.0),
                                           splitk, None, block_n, pair)
    assert plan(Cin=65) == -1 and "multiple of 64" in L.last_error()
    assert plan(taps=4) == -1
    assert plan(block_n=96) == -1 and "block_n" in L.last_error()
    assert plan(block_n=64, pair=1) == -1
    assert plan(block_n=256, pair=1, w_sy=64) == -1 and "cta_pair" in L.last_error()
    assert plan(act=3, N=104) == -1 and "GEGLU" in L.last_error()
    # strided 3x3 convolution plans: stride in {1, 2} (2 only with taps = 9), pad_lo in {0, 1}
    def splan(taps=9, stride=2, pad_lo=1):
        return L.cdll.sdf_gemm_plan_create_strided(fake, LL(64), LL(64 * 16), LL(64 * 256), 64, fake, LL(9 * 64), LL(0), LL(0), 9 * 64, 128, 1, 8, 8, 64, taps, 128,
                                                   fake, LL(128), LL(128 * 8), LL(128 * 64), None, None, 0, None, LL(0), LL(0), LL(0), 0, ctypes.c_float(1.0),
                                                   1, None, 128, 0, stride, pad_lo)
    assert splan(stride=3) == -1 and "stride" in L.last_error()
    assert splan(taps=1, stride=2) == -1
    assert splan(pad_lo=2) == -1 and "pad_lo" in L.last_error()
    assert plan(splitk=2) == -1 and "workspace" in L.last_error()
    # GroupNorm / LayerNorm shape contracts
    assert L.cdll.sdf_groupnorm_forward(fake, 12, fake, 12, 1, 4, 12, 4, fake, fake, ctypes.c_float(1e-5), 1, fake, None) == -1
    assert L.cdll.sdf_layernorm_forward(fake, 4096, fake, 4096, 1, 4096, fake, fake, ctypes.c_float(1e-5), None) == -1
    assert "2048" in L.last_error()
