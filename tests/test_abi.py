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
